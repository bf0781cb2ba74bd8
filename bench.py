#!/usr/bin/env python
"""Headline benchmark: GAN training voxels/s on the upscale8 configuration (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

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

HR_XY, NZ, SCALE, BATCH_PER_GPU = 128, 10, 8, 8
VOX_PER_SAMPLE = HR_XY * HR_XY * NZ
INI = os.path.join(ROOT, "configs", "upscale8_pix4_no_adv_no_slicing.ini")
WORKLOAD = ("upscale8_pix4_no_adv_no_slicing G-step (G fwd+dgrad+wgrad, D fwd x2 + dgrad, wind loss, Adam), "
            "batch 8/GPU, LR 16x16x10 -> HR 128x128x10")


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
    ops.set_precision("bf16")

    cfg = Config(INI)
    cfg.is_train, cfg.gpu_id, cfg.device = True, local, dev
    torch.manual_seed(cfg.env.fixed_seed)  # identical replicas on every rank
    gan = wind_field_GAN_3D(cfg)
    B = BATCH_PER_GPU
    LR, HR, Z, x, y = make_batch(B, HR_XY, NZ, SCALE, seed=cfg.env.fixed_seed + rank, device=dev)
    t = cfg.training
    gan.feed_xy_niter(x, y, torch.tensor(t.niter, device=dev), t.d_g_train_ratio, t.d_g_train_period)
    host = [v.cpu().pin_memory() for v in (LR, HR, Z)]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        # Python's cyclic GC is kept out of the timed region (collected between legs instead): a generation-2 pass
        # over the autograd graphs of a 2000-launch step takes ~100 ms on one rank and, under data parallelism,
        # stalls every rank at the next all-reduce.
        gc.collect()
        gc.disable()
        try:
            return _timed_inner(fn, steps)
        finally:
            gc.enable()

    def _timed_inner(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks, host = [], []
        e0.record()
        for i in range(steps):
            h0 = time.perf_counter()
            fn(i)
            if os.environ.get("BENCH_DEBUG_LEGS"):
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record()
                host.append(round(1e3 * (time.perf_counter() - h0), 1))
        e1.record()
        barrier()
        if marks:
            prev, per = e0, []
            for m in marks:
                per.append(round(prev.elapsed_time(m), 1))
                prev = m
            print(f"[rank {rank}] per-step ms: {per} host enqueue ms: {host}", file=sys.stderr, flush=True)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident(i):
        gan.optimize_parameters(LR, HR, Z, 1 + i)

    # e2e: every step copies its batch from pinned host memory and reads its loss back to the host.  The read is
    # the usual logging pattern of a training loop: the loss of step i is copied to pinned memory right after the
    # step is enqueued and consumed (event wait + float) while step i+1 runs; the last one is drained inside the
    # timed region by the closing synchronize.
    LAG = 2  # the host consumes the loss of step i - LAG while step i is enqueued (bounded run-ahead, as in the step)
    loss_host = torch.zeros(LAG + 1, dtype=torch.float32).pin_memory()
    loss_ready = [torch.cuda.Event() for _ in range(LAG + 1)]
    loss_log = []
    # Input pipeline of the e2e leg: the batch of step i+1 is copied from pinned host memory into the other of two
    # device staging buffers on a copy stream while step i computes (the usual prefetch-to-device pattern); every
    # step's H2D copy is issued inside the timed region.  A copy on the compute stream itself exposed the step to
    # PCIe hiccups of this shared host (sporadic 40-120 ms stalls in half of the runs).
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

    def step_e2e(i):
        if i == 0:
            issue_copy(0)
        buf = i % 2
        torch.cuda.current_stream().wait_event(ready[buf])
        lr, hr, z = stage[buf]
        gan.optimize_parameters(lr, hr, z, 1 + i)
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
    # settle: the caching allocator and the tensor-map cache keep changing for a few more steps after a cold
    # start; keep warming (untimed) until two consecutive steps agree within 3 % (at most 12 extra steps)
    prev = None
    for i in range(12):
        cur = timed(step_resident, 1)
        if prev is not None and abs(cur - prev) <= 0.03 * prev:
            break
        prev = cur
    # pre-roll: ~0.6 s of back-to-back steps before the first timed leg.  The settle loop above synchronises after
    # every step, so the first CONTINUOUS run of steps used to start inside the timed region — and 40-130 ms stalls
    # showed up 3-4 steps (~0.2 s) into it in half of the runs (never in the later legs).
    for i in range(10):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    if not all(math.isfinite(v) for v in loss_log):
        raise SystemExit(f"bench.py: non-finite generator loss in the e2e leg: {loss_log}")
    # roofline leg: time the dominant kernel (hr_convs.0 forward: 5x5x5, 144->144 @128x128x10, 69.6 % of G's
    # FLOPs) in-stream during the timed region
    is_g7 = lambda kind, s: kind == "fwd" and s.kx == 5 and s.cin == s.cout and s.x == HR_XY
    if os.environ.get("BENCH_DEBUG_LEGS"):  # diagnostic: the resident leg without timer / sampler, per rank
        ms_plain = timed(step_resident, args.steps)
        print(f"[rank {rank}] resident leg without kernel timer / clock sampler: {ms_plain / args.steps:.2f} ms/step",
              file=sys.stderr, flush=True)
    if os.environ.get("BENCH_DEBUG_LEGS") != "notimer":
        ops.set_kernel_timer(is_g7)
    launches0 = ops.launch_count()
    clocks = ClockSampler(local if os.environ.get("BENCH_DEBUG_LEGS") != "nosampler" else -1)
    ms = timed(step_resident, args.steps)
    clock_info = clocks.stop()
    launches = ops.launch_count() - launches0
    events = ops.kernel_timer_events()
    k_ms = [a.elapsed_time(b) for a, b in events]
    ops.set_kernel_timer(None)

    # Secondary leg (BASELINE config #3): the full GAN schedule — adversarial weight 0.0005, G and D steps
    # alternating (d_g_train_ratio 1) — on the same shapes; G-step and D-step times are reported separately and as the
    # 50/50 blend.  Not the headline (the shipped upscale8 ini trains G only), so it runs after the timed regions.
    full_gan = None
    if not args.no_full_gan:
        try:
            cfg2 = Config(INI)
            cfg2.is_train, cfg2.gpu_id, cfg2.device = True, local, dev
            cfg2.training.adversarial_loss_weight = 0.0005
            cfg2.training.d_g_train_ratio, cfg2.training.d_g_train_period = 1, 1
            torch.manual_seed(cfg2.env.fixed_seed)
            gan2 = wind_field_GAN_3D(cfg2)
            gan2.feed_xy_niter(x, y, torch.tensor(t.niter, device=dev), 1, 1)
            for i in range(6):
                gan2.optimize_parameters(LR, HR, Z, i)
            tg, td, n_each = 0.0, 0.0, max(3, args.steps // 2)
            for i in range(6, 6 + 2 * n_each):
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                gan2.optimize_parameters(LR, HR, Z, i)
                b.record()
                barrier()
                ms_i = torch.tensor([a.elapsed_time(b)], device=dev)
                if world > 1:
                    dist.all_reduce(ms_i, op=dist.ReduceOp.MAX)
                if gan2.is_G_iteration(i):
                    tg += float(ms_i)
                else:
                    td += float(ms_i)
            tg, td = tg / n_each, td / n_each
            full_gan = {"g_step_ms": tg, "d_step_ms": td, "steps_per_s_g": 1e3 / tg, "steps_per_s_d": 1e3 / td,
                        "steps_per_s_blend": 2e3 / (tg + td),
                        "voxels_per_s_blend": world * B * VOX_PER_SAMPLE * 2e3 / (tg + td),
                        "note": "adversarial_loss_weight 0.0005, d_g_train_ratio 1, period 1; each step timed alone "
                                "(sync on both sides), so these are upper bounds on the pipelined step time"}
            del gan2

        except Exception as exc:  # noqa: BLE001 - the secondary leg must never cost the headline line
            if world > 1:
                raise  # ranks would desynchronise: fail loudly under torchrun
            full_gan = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    vox = world * B * VOX_PER_SAMPLE
    tflops_peak, hbm_peak, peak_src = _peaks()
    cin = cout = cfg.generator.num_features + cfg.generator.terrain_number_of_features
    flops = 2.0 * B * VOX_PER_SAMPLE * cin * cout * 125  # dense-MAC convention
    k_avg = sum(k_ms) / max(1, len(k_ms))
    achieved = flops / (k_avg * 1e-3) / 1e12 if k_avg > 0 else 0.0
    out = {
        "metric": "GAN train voxels/sec", "value": vox / (ms / args.steps * 1e-3), "unit": "HR voxels/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B, "ini": os.path.basename(INI),
                   "l2": "per-step working set (> 5 GB of activations) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"dp{world}", "steps_per_s": args.steps / (ms * 1e-3), "full_gan": full_gan},
        "e2e": {"value": vox / (ms_e2e / args.steps * 1e-3), "unit": "HR voxels/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps,
                "loss_read": "async D2H into pinned memory every step, consumed on the host two steps later"},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": {"bound": "tensor", "kernel": "conv3d_tc2_kernel<true> (hr_convs.0 fwd, 5x5x5 144->144 @128x128x10, B=8)",
                     "achieved": achieved, "peak": tflops_peak, "unit": "TFLOP/s",
                     "frac": achieved / tflops_peak if tflops_peak else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture
                     # profiles/r01_ncu_conv3d_tc2_pair_g7_fwd_v2.md (algorithmic: 760.1 MB)
                     "traffic": 720.0e6, "traffic_unit": "bytes/launch",
                     "peak_source": f"{peak_src} bf16_tflops_sustained", "kernel_ms": k_avg,
                     "kernel_launches_timed": len(k_ms), "flops_per_launch": flops},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference_step(steps=1, warmup=0)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------
def cpu_reference_step(steps: int, warmup: int):
    """The reference's CPU path for one G step (oracle port, the only place bench.py touches ``oracle/``):
    bounded sample = batch 1 of the same upscale8 workload, all host threads."""
    import torch

    from gan_sr_wind_field_b200.CNN_models.Discriminator_3D import Discriminator_3D
    from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    from gan_sr_wind_field_b200.config.config import Config
    from gan_sr_wind_field_b200.tools import initialization
    from oracle import wind_oracle as wo

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
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
    LR, HR, Z, x, y = wo.synthetic_batch(1, HR_XY, NZ, SCALE, seed=cfg.env.fixed_seed)
    w = dict(pixel=t.pixel_loss_weight, xy=t.gradient_xy_loss_weight, z=t.gradient_z_loss_weight,
             div=t.divergence_loss_weight, dxy=t.xy_divergence_loss_weight, adv=t.adversarial_loss_weight)
    real, fake = torch.full((1,), 0.9), torch.zeros(1)

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

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": VOX_PER_SAMPLE / dt, "unit": "HR voxels/s", "cores": threads, "kind": "port",
            "sample": f"{steps} G step(s) at batch 1 of the same upscale8 workload, fp32, torch CPU ops "
                      f"(oneDNN) via oracle/wind_oracle.py, {dt:.2f} s/step",
            "s_per_step": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    base = cpu_reference_step(steps=steps, warmup=warm)
    out = {
        "impl": "reference", "metric": "GAN train voxels/sec", "value": base["value"], "unit": "HR voxels/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warm,
        "ms_per_step": base["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": 1, "ini": os.path.basename(INI),
                   "note": "reference CPU path, bounded sample: batch 1 per step"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "HR voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~30 s CPU baseline leg")
    ap.add_argument("--no-full-gan", action="store_true", help="skip the G-step / D-step leg of the full GAN schedule")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
