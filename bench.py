#!/usr/bin/env python
"""Headline benchmark: GAN training voxels/s on the upscale8 configuration (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--config upscale8|upscale16] [--scaling weak|strong] [--dtype bf16|tf32|fp32]

Workload (``config.workload``): BASELINE.json configs[1] — the shipped
``upscale8_pix4_no_adv_no_slicing`` configuration (scale 8, 128 features, 16 RRDBs, 5x5x5 HR convs, batch 8 per
GPU, adversarial weight 0 / d_g_train_ratio 0, so every iteration is a G step that — exactly like the
reference, SURVEY §0-7 — still runs D forward twice and D's data-gradient), synthetic HARMONIE-SIMRA-shaped
volumes, BF16 tensor-core arithmetic with fp32 accumulation.  One "step" = one
``wind_field_GAN_3D.optimize_parameters`` call: G forward, D forward x2, fused wind loss, full backward
(dgrad + wgrad of every G conv, dgrad through D), Adam update.

Metric: HR voxels / s = B_global * 128*128*10 / step time.  ``value`` is measured with the batch
resident in HBM; ``e2e`` times the same call with the batch in pinned HOST memory, the H2D copies of LR/HR/Z and
the D2H read of the loss inside the timed region.

N > 1: launched under torchrun, one process per GPU, batch-sharded (weak scaling: 8 samples per GPU), bucketed
NCCL gradient all-reduce overlapped with backward; timing = max over ranks.

``--impl reference``: the reference's own CPU path for the same step (oracle port of the reference's modules —
the reference itself is Python and cannot travel to the GPU box) on the host cores, on a
bounded sample (batch 1) of the same workload.
"""
from __future__ import annotations

import argparse
import gc
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HR_XY, NZ, BATCH_PER_GPU = 128, 10, 8
VOX_PER_SAMPLE = HR_XY * HR_XY * NZ
CONFIGS = {  # BASELINE.json configs[1] (headline) and configs[3] (the DDP configuration)
    "upscale8": (os.path.join(ROOT, "configs", "upscale8_pix4_no_adv_no_slicing.ini"), 8),
    "upscale16": (os.path.join(ROOT, "configs", "upscale16_pix4_no_adv_no_slicing.ini"), 16),
}
INI, SCALE = CONFIGS["upscale8"]


def workload(name, scale, batch, d_skipped):
    d = ("D not run: adversarial_loss_weight = 0 makes its term an exact zero" if d_skipped
         else "D fwd x2 + dgrad")
    return (f"{name}_pix4_no_adv_no_slicing G-step (G fwd+dgrad+wgrad, {d}, wind loss, Adam), batch {batch}/GPU, "
            f"LR {HR_XY // scale}x{HR_XY // scale}x{NZ} -> HR {HR_XY}x{HR_XY}x{NZ}")


def conv_flops(cfg, scale, with_d: bool):
    """Dense-MAC FLOPs PER SAMPLE of one G step (SURVEY §8-a convention: 2*V*Cout*Cin*taps per pass, padding taps
    counted): forward + data-gradient + weight-gradient of every generator conv (no data-gradient for the two layers
    that read the inputs), plus D forward x2 + D data-gradient when the discriminator runs."""
    g = cfg.generator
    F, gc, nd, T = g.num_features, g.RDB_growth_chan, g.num_RDB_convs - 1, g.terrain_number_of_features
    in_ch = g.in_num_ch + int(bool(cfg.gan_config.include_pressure)) + int(bool(cfg.gan_config.include_z_channel)) \
        + int(bool(cfg.gan_config.include_above_ground_channel))
    lr_v, hr_v = (HR_XY // scale) ** 2 * NZ, VOX_PER_SAMPLE
    layers = [(lr_v, in_ch, F, 27, False)]                                        # G1 feature conv (no dgrad)
    for _ in range(g.num_RRDB * 3):
        layers += [(lr_v, F + i * gc, gc, 27, True) for i in range(nd)]           # G2 dense convs
        layers.append((lr_v, F + nd * gc, F, g.lff_kern_size ** 3, True))         # G3 LFF
    layers.append((lr_v, F, F, 27, True))                                         # G4 lr_conv
    v = lr_v
    for _ in range(int(math.log2(scale))):
        v *= 4
        layers.append((v, F, F, 27, True))                                        # G5 UpConvs
    k3 = g.hr_kern_size ** 3
    layers += [(hr_v, 1, T, 27, False), (hr_v, T, T, 27, True),                   # G6 terrain convs
               (hr_v, F + T, F + T, k3, True), (hr_v, F + T, g.out_num_ch, k3, True)]  # G7, G8
    fwd = sum(2.0 * v * ci * co * t for v, ci, co, t, _ in layers)
    total = sum(2.0 * v * ci * co * t * (3 if dg else 2) for v, ci, co, t, dg in layers)
    d_fwd = 0.0
    if with_d:
        f = cfg.discriminator.num_features
        xy, chans = HR_XY, [(cfg.discriminator.in_num_ch, f), (f, 2 * f), (2 * f, 4 * f), (4 * f, 8 * f), (8 * f, 8 * f)]
        for i, (ci, co) in enumerate(chans):
            zo = NZ // 2 if i == len(chans) - 1 else NZ
            d_fwd += 2.0 * xy * xy * NZ * ci * co * 27 + 2.0 * (xy // 2) ** 2 * zo * co * co * 48
            xy //= 2
        total += 3.0 * d_fwd
    return {"g_fwd": fwd, "step": total, "d_fwd": d_fwd}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"  # B200_PROFILING.md: sustained bf16 / copy bandwidth fallback


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled in-process through
    NVML every 100 ms from a daemon thread (a polling nvidia-smi subprocess measurably slowed the launch path of
    this launch-heavy step); falls back to one nvidia-smi query if NVML is unavailable."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        import threading
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.gpu_index = gpu_index
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[gpu_index]) if visible and visible.split(",")[gpu_index].isdigit() \
                else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.1)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if self.nv is None or not self.sm:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits", "-i", str(self.gpu_index)],
                                     capture_output=True, text=True, timeout=20).stdout.strip().split(", ")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1,
                        "reasons": ["sampled once after the timed region (NVML unavailable)"]}
            except Exception:  # noqa: BLE001
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "samples": len(self.sm),
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------------------
def _dist_stats(ms_list):
    s = sorted(ms_list)
    return {"min": s[0], "median": statistics.median(s), "max": s[-1], "n": len(s)}


def run_native(args):
    import torch
    import torch.distributed as dist

    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.config.config import Config
    from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
    from gan_sr_wind_field_b200.synthetic import make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.set_precision(args.dtype)
    ini, scale = CONFIGS[args.config]

    cfg = Config(ini)
    cfg.is_train, cfg.gpu_id, cfg.device = True, local, dev
    torch.manual_seed(cfg.env.fixed_seed)
    gan = wind_field_GAN_3D(cfg)  # under torchrun: broadcasts rank 0's weights, then reseeds per rank
    if args.scaling == "strong":
        if GLOBAL_BATCH_STRONG % world:
            raise SystemExit(f"strong scaling: global batch {GLOBAL_BATCH_STRONG} not divisible by {world} GPUs")
        B = GLOBAL_BATCH_STRONG // world
    else:
        B = BATCH_PER_GPU
    LR, HR, Z, x, y = make_batch(B, HR_XY, NZ, scale, seed=cfg.env.fixed_seed + rank, device=dev)
    t = cfg.training
    gan.feed_xy_niter(x, y, torch.tensor(t.niter, device=dev), t.d_g_train_ratio, t.d_g_train_period)
    host = [v.cpu().pin_memory() for v in (LR, HR, Z)]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host)
    d_skipped = gan._skip_D_in_G_step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """(total ms, per-step ms list): CUDA events on the compute stream around the whole leg and after every step;
        barrier + synchronize on both sides; max over ranks.  Python's cyclic GC stays out of the timed region."""
        gc.collect()
        gc.disable()
        try:
            barrier()
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
            marks[0].record()
            for i in range(steps):
                fn(i)
                marks[i + 1].record()
            barrier()
        finally:
            gc.enable()
        per = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        ms = torch.tensor([marks[0].elapsed_time(marks[-1])], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), per

    it_counter = [0]

    def next_it():
        it_counter[0] += 1
        return it_counter[0]

    def step_resident(i):
        gan.optimize_parameters(LR, HR, Z, next_it())

    # e2e: every step copies its batch from pinned host memory and reads its loss back to the host.  The read is
    # the usual logging pattern of a training loop: the loss of step i is copied to pinned memory right after the
    # step is enqueued and consumed (event wait + float) while step i+1 runs; the last one is drained inside the
    # timed region by the closing synchronize.
    LAG = 2
    loss_host = torch.zeros(LAG + 1, dtype=torch.float32).pin_memory()
    loss_ready = [torch.cuda.Event() for _ in range(LAG + 1)]
    loss_log = []
    # Input pipeline of the e2e leg: the batch of step i+1 is copied from pinned host memory into the other of two
    # device staging buffers on a copy stream while step i computes (the usual prefetch-to-device pattern); every
    # step's H2D copy is issued inside the timed region.
    copy_stream = torch.cuda.Stream()
    stage = [[torch.empty_like(v, device=dev) for v in host] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(k):
        buf = k % 2
        copy_stream.wait_event(freed[buf])  # the step that last read this buffer has finished
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(stage[buf], host):
                dst.copy_(src, non_blocking=True)
            ready[buf].record()

    e2e_i = [0]

    def step_e2e(_):
        i = e2e_i[0]
        e2e_i[0] += 1
        if i == 0:
            issue_copy(0)
        buf = i % 2
        torch.cuda.current_stream().wait_event(ready[buf])
        lr, hr, z = stage[buf]
        gan.optimize_parameters(lr, hr, z, next_it())
        freed[buf].record()
        issue_copy(i + 1)  # prefetch the next batch behind this step's kernels
        slot = i % (LAG + 1)
        loss_host[slot:slot + 1].copy_(gan.get_G_train_loss_dict_ref()["total"].detach().reshape(1), non_blocking=True)
        loss_ready[slot].record()
        if i >= LAG:
            old = (i - LAG) % (LAG + 1)
            loss_ready[old].synchronize()
            loss_log.append(float(loss_host[old]))

    for i in range(args.warmup):
        step_resident(i)
    # settle: after the eager warm-up calls the step is captured into a CUDA graph (GAN_models/graph_step.py) and the
    # caching allocator / tensor-map cache keep changing for a few more steps; keep warming (untimed) until two
    # consecutive steps agree within 3 % (at most 12 extra steps)
    prev = None
    for i in range(12):
        cur, _ = timed(step_resident, 1)
        if prev is not None and abs(cur - prev) <= 0.03 * prev:
            break
        prev = cur
    for i in range(6):  # pre-roll of back-to-back steps (the settle loop synchronises after every step)
        step_e2e(i)
    ms_e2e, per_e2e = timed(step_e2e, args.steps)
    if not all(math.isfinite(v) for v in loss_log):
        raise SystemExit(f"bench.py: non-finite generator loss in the e2e leg: {loss_log}")

    launches0 = ops.launch_count()
    clocks = ClockSampler(local)
    ms, per_res = timed(step_resident, args.steps)
    clock_info = clocks.stop()
    launches = ops.launch_count() - launches0
    graphs = {k[0]: g for k, g in gan._graphs.items() if g}
    graph_on = "G" in graphs

    # roofline leg: the dominant kernel (hr_convs.0 forward: 5x5x5, 144->144 @128x128x10, 69.6 % of G's FLOPs) timed
    # in-stream with CUDA events around its launch.  Events cannot sit inside a captured graph, so this leg runs the
    # same step eagerly (same kernels, same stream) for the same number of steps.
    is_g7 = lambda kind, s: kind == "fwd" and s.kx == 5 and s.cin == s.cout and s.x == HR_XY
    os.environ["WINDSR_CUDA_GRAPH"] = "0"
    ops.set_kernel_timer(is_g7)
    ms_eager, per_eager = timed(step_resident, args.steps)
    events = ops.kernel_timer_events()
    k_ms = [a.elapsed_time(b) for a, b in events]
    ops.set_kernel_timer(None)
    eager_plain = None
    with_d = None
    if not args.quick:
        eager_plain, _ = timed(step_resident, args.steps)  # eager, no timer events: what graph capture removes
        if d_skipped:
            # the same step with the reference's order of operations (D forward x2 + D data-gradient, times 0.0)
            os.environ["WINDSR_SKIP_D_WHEN_ZERO"] = "0"
            for i in range(3):
                step_resident(i)
            with_d, _ = timed(step_resident, args.steps)
            os.environ.pop("WINDSR_SKIP_D_WHEN_ZERO")
    os.environ.pop("WINDSR_CUDA_GRAPH")

    # Secondary leg (BASELINE config #3): the full GAN schedule — adversarial weight 0.0005, G and D steps
    # alternating (d_g_train_ratio 1) — on the same shapes; G-step and D-step times are reported separately and as the
    # 50/50 blend.  Not the headline (the shipped ini trains G only), so it runs after the timed regions.
    full_gan = None
    if not args.no_full_gan and not args.quick:
        try:
            cfg2 = Config(ini)
            cfg2.is_train, cfg2.gpu_id, cfg2.device = True, local, dev
            cfg2.training.adversarial_loss_weight = 0.0005
            cfg2.training.use_instance_noise = True
            cfg2.training.d_g_train_ratio, cfg2.training.d_g_train_period = 1, 1
            torch.manual_seed(cfg2.env.fixed_seed)
            gan2 = wind_field_GAN_3D(cfg2)
            gan2.feed_xy_niter(x, y, torch.tensor(t.niter, device=dev), 1, 1)
            for i in range(12):  # 6 G + 6 D calls: eager warm-up, capture, replays
                gan2.optimize_parameters(LR, HR, Z, i)
            n_each = max(3, args.steps // 2)
            pair_ms, per_pair = timed(lambda i: gan2.optimize_parameters(LR, HR, Z, 12 + i), 2 * n_each)
            tg = statistics.mean(per_pair[0::2]) if gan2.is_G_iteration(12) else statistics.mean(per_pair[1::2])
            td = statistics.mean(per_pair[1::2]) if gan2.is_G_iteration(12) else statistics.mean(per_pair[0::2])
            full_gan = {"g_step_ms": tg, "d_step_ms": td, "steps_per_s_g": 1e3 / tg, "steps_per_s_d": 1e3 / td,
                        "steps_per_s_blend": 2 * n_each / (pair_ms * 1e-3),
                        "voxels_per_s_blend": world * B * VOX_PER_SAMPLE * 2 * n_each / (pair_ms * 1e-3),
                        "note": "adversarial_loss_weight 0.0005, instance noise on, d_g_train_ratio 1, period 1; "
                                "alternating G / D steps timed back to back (one sync at each end)"}
            del gan2
        except Exception as exc:  # noqa: BLE001 - the secondary leg must never cost the headline line
            if world > 1:
                raise  # ranks would desynchronise: fail loudly under torchrun
            full_gan = {"error": f"{type(exc).__name__}: {exc}"}

    # BASELINE config #1: generator inference, batch 1, eval mode (the reference's CPU-runnable case) on the GPU
    infer = None
    if world == 1 and not args.quick:
        gan.G.eval()
        with torch.no_grad():
            for _ in range(3):
                gan.G(LR[:1], Z[:1])
            ms_inf, _ = timed(lambda i: gan.G(LR[:1], Z[:1]), 10)
        infer = {"ms": ms_inf / 10, "voxels_per_s": VOX_PER_SAMPLE / (ms_inf / 10 * 1e-3), "batch": 1}

    def teardown():
        """Captured CUDA graphs hold NCCL kernels: they have to go before the communicator does (destroying the
        process group first left both ranks hanging in NCCL's teardown until its watchdog fired 8 minutes later).  A
        timer guarantees the process ends even if a teardown path blocks."""
        if world == 1:
            return
        import threading
        sys.stdout.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        gan._graphs.clear()
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)

    if rank != 0:
        teardown()
        return
    vox = world * B * VOX_PER_SAMPLE
    tflops_peak, hbm_peak, peak_src = _peaks()
    fl = conv_flops(cfg, scale, with_d=not d_skipped)
    cin = cout = cfg.generator.num_features + cfg.generator.terrain_number_of_features
    flops = 2.0 * B * VOX_PER_SAMPLE * cin * cout * cfg.generator.hr_kern_size ** 3  # SURVEY §8-a G7, dense-MAC
    k_avg = sum(k_ms) / max(1, len(k_ms))
    achieved = flops / (k_avg * 1e-3) / 1e12 if k_avg > 0 else 0.0
    step_tflops = fl["step"] * B / (ms / args.steps * 1e-3) / 1e12
    peak_scale = {"bf16": 1.0, "tf32": 0.5, "fp32": None}[args.dtype]
    peak = tflops_peak * peak_scale if peak_scale else None
    out = {
        "metric": "GAN train voxels/sec", "value": vox / (ms / args.steps * 1e-3), "unit": "HR voxels/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.dtype], "data": "synthetic",
        "config": {"workload": workload(args.config, scale, B, d_skipped), "global_batch": world * B,
                   "ini": os.path.basename(ini),
                   "l2": "per-step working set (> 5 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"dp{world}", "steps_per_s": args.steps / (ms * 1e-3),
                   "cuda_graph": graph_on, "discriminator_in_g_step": "skipped" if d_skipped else "run",
                   "per_step_ms": {"resident": _dist_stats(per_res), "e2e": _dist_stats(per_e2e),
                                   "eager_with_kernel_timer": _dist_stats(per_eager)},
                   "eager_ms_per_step": eager_plain / args.steps if eager_plain else None,
                   "with_discriminator_path_ms_per_step": with_d / args.steps if with_d else None,
                   "full_gan": full_gan, "inference_b1": infer},
        "e2e": {"value": vox / (ms_e2e / args.steps * 1e-3), "unit": "HR voxels/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps,
                "loss_read": "async D2H into pinned memory every step, consumed on the host two steps later"},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": {"bound": "tensor", "kernel": "conv3d_tc2_kernel<true> (hr_convs.0 fwd, 5x5x5 144->144 @128x128x10)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None,
                     "traffic": _measured_traffic(), "traffic_unit": "bytes/launch (B = 8 capture)",
                     "peak_source": f"{peak_src} bf16_tflops_sustained" + (" / 2 (tf32)" if args.dtype == "tf32" else ""),
                     "kernel_ms": k_avg, "kernel_launches_timed": len(k_ms), "flops_per_launch": flops,
                     "timed_in": "eager leg of the same step (CUDA events cannot be recorded inside a graph replay)",
                     "kernel_share_of_eager_step": k_avg * len(k_ms) / ms_eager if ms_eager else None},
        "roofline_step": {"bound": "tensor", "what": "all Conv3d passes of the step (dense-MAC FLOPs) / step time",
                          "achieved": step_tflops, "peak": peak, "unit": "TFLOP/s",
                          "frac": step_tflops / peak if peak else None, "flops_per_step": fl["step"] * B,
                          "north_star_target_frac": 0.60},
    }
    if world == 1 and not args.no_cpu_baseline and not args.quick:
        out["cpu_baseline"] = cpu_reference_step(steps=1, warmup=0, batch=1)
    print(json.dumps(out), flush=True)
    teardown()


def _measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one hr_convs.0 forward launch, from the committed ncu capture
    (profiles/*_traffic.json, written by scripts/ncu_summary.py from the .ncu-rep of the same command)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if name.endswith("_traffic.json"):
            try:
                best = json.load(open(os.path.join(pdir, name))).get("conv3d_tc2_pair_g7_fwd", best)
            except Exception:  # noqa: BLE001
                pass
    return best


# ------------------------------------------------------------------------------------------------------------
def _stock_reference_available():
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "CNN_models"))


def cpu_reference_step(steps: int, warmup: int, batch: int = 1):
    """One G step of the reference on the host cores.  With ``baseline/_ref`` present (baseline/install_ref.py copies
    the unmodified reference there) it is the STOCK code: ``Config(ini)`` -> ``wind_field_GAN_3D(cfg)`` ->
    ``optimize_parameters`` exactly as train.py:61-148 drives it (kind "reference"); otherwise the oracle port of the
    same step (kind "port").  All host threads, fp32, synthetic batch of ``batch`` samples of the same workload."""
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import wind_oracle as wo
    LR, HR, Z, x, y = wo.synthetic_batch(batch, HR_XY, NZ, SCALE, seed=2001)
    if _stock_reference_available():
        from oracle import refshim
        refshim.REFERENCE_ROOT = os.path.join(ROOT, "baseline", "_ref")
        refshim.activate()
        import config.config as ref_config
        from GAN_models.wind_field_GAN_3D import wind_field_GAN_3D as RefGAN
        cfg = ref_config.Config(os.path.join(refshim.REFERENCE_ROOT, "pretrained_models",
                                             "upscale8_pix4_no_adv_no_slicing", "config.ini"))
        cfg.is_train, cfg.gpu_id, cfg.device = True, None, torch.device("cpu")
        torch.manual_seed(cfg.env.fixed_seed)
        gan = RefGAN(cfg)
        t = cfg.training
        gan.feed_xy_niter(x, y, torch.tensor(t.niter), t.d_g_train_ratio, t.d_g_train_period)
        it = [0]

        def step():
            it[0] += 1
            gan.optimize_parameters(LR, HR, Z, it[0])
            return float(gan.get_G_train_loss_dict_ref()["total"])

        kind, how = "reference", "the unmodified reference (baseline/_ref): wind_field_GAN_3D.optimize_parameters"
    else:
        step = _port_step(LR, HR, Z, x, y)
        kind, how = "port", "oracle/wind_oracle.py (torch CPU ops)"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch * VOX_PER_SAMPLE / dt, "unit": "HR voxels/s", "cores": threads, "kind": kind,
            "sample": f"{steps} G step(s) at batch {batch} of the same upscale8 workload, fp32, {how}, "
                      f"{dt:.2f} s/step, loss {loss:.4f}",
            "s_per_step": dt}


def _port_step(LR, HR, Z, x, y):
    import torch

    from gan_sr_wind_field_b200.CNN_models.Discriminator_3D import Discriminator_3D
    from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    from gan_sr_wind_field_b200.config.config import Config
    from gan_sr_wind_field_b200.tools import initialization
    from oracle import wind_oracle as wo
    cfg = Config(INI)
    g, d, t = cfg.generator, cfg.discriminator, cfg.training
    torch.manual_seed(cfg.env.fixed_seed)
    G = Generator_3D(g.in_num_ch + 1, g.out_num_ch, g.num_features, g.num_RRDB, upscale=cfg.scale,
                     hr_kern_size=g.hr_kern_size, number_of_RDB_convs=g.num_RDB_convs, RDB_gc=g.RDB_growth_chan,
                     lff_kern_size=g.lff_kern_size, terrain_number_of_features=g.terrain_number_of_features,
                     dropout_probability=g.dropout_probability)
    initialization.init_weights(G, g.weight_init_scale)
    D = Discriminator_3D(d.in_num_ch, d.num_features, feat_kern_size=d.feat_kern_size)
    initialization.init_weights(D, d.weight_init_scale)
    pG = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in G.state_dict().items()}
    pD = {k: v.detach().clone() for k, v in D.state_dict().items()}
    opt = torch.optim.Adam([v for v in pG.values() if v.requires_grad], lr=t.learning_rate_g,
                           betas=(t.adam_beta1_g, 0.999))
    w = dict(pixel=t.pixel_loss_weight, xy=t.gradient_xy_loss_weight, z=t.gradient_z_loss_weight,
             div=t.divergence_loss_weight, dxy=t.xy_divergence_loss_weight, adv=t.adversarial_loss_weight)
    n = HR.shape[0]
    real, fake = torch.full((n,), 0.9), torch.zeros(n)

    def step():
        opt.zero_grad(set_to_none=True)
        SR = wo.generator_forward(pG, LR, Z)
        with torch.no_grad():
            y_pred = wo.discriminator_forward(pD, HR, False).reshape(-1)
        y_fake = wo.discriminator_forward(pD, SR, False).reshape(-1)
        adv = wo.adversarial_G(y_pred, y_fake, real, fake, t.gan_type)
        total, _ = wo.generator_loss(HR, SR, Z, x, y, w, adv=adv)
        total.backward()
        opt.step()
        return float(total)

    return step


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the same step, SAME configuration (upscale8 ini,
    batch 8), all host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2))
    warm = 1 if args.warmup > 0 else 0
    batch = args.ref_batch
    base = cpu_reference_step(steps=steps, warmup=warm, batch=batch)
    out = {
        "impl": "reference", "metric": "GAN train voxels/sec", "value": base["value"], "unit": "HR voxels/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warm,
        "ms_per_step": base["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload("upscale8", SCALE, batch, False), "global_batch": batch,
                   "ini": os.path.basename(INI),
                   "note": "reference CPU path on the host cores; one process whatever --gpus says (the reference "
                           "has no distributed code)"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "HR voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


GLOBAL_BATCH_STRONG = 32  # config/wind_field_GAN_3D_config_cluster.ini:47


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="upscale8", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 8 samples per GPU; strong: global batch 32 split over the GPUs")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--ref-batch", type=int, default=BATCH_PER_GPU,
                    help="--impl reference: samples per step (default: the workload's batch of 8)")
    ap.add_argument("--quick", action="store_true", help="headline legs only (no secondary legs, no CPU baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~30 s CPU baseline leg")
    ap.add_argument("--no-full-gan", action="store_true", help="skip the G-step / D-step leg of the full GAN schedule")
    args = ap.parse_args()
    if args.impl == "native":
        args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
