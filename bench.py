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
import json
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
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(gpu_index)], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, val in zip(names, r[2:6]):
                if val.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


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
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident(i):
        gan.optimize_parameters(LR, HR, Z, 1 + i)

    def step_e2e(i):
        lr, hr, z = (v.to(dev, non_blocking=True) for v in host)
        gan.optimize_parameters(lr, hr, z, 1 + i)
        return float(gan.get_G_train_loss_dict_ref()["total"])  # D2H read of the step's loss

    for i in range(args.warmup):
        step_resident(i)
    # roofline leg: time the dominant kernel (hr_convs.0 forward: 5x5x5, 144->144 @128x128x10, 69.6 % of G's
    # FLOPs) in-stream during the timed region
    is_g7 = lambda kind, s: kind == "fwd" and s.kx == 5 and s.cin == s.cout and s.x == HR_XY
    ops.set_kernel_timer(is_g7)
    launches0 = ops.launch_count()
    clocks = ClockSampler(local)
    ms = timed(step_resident, args.steps)
    clock_info = clocks.stop()
    launches = ops.launch_count() - launches0
    events = ops.kernel_timer_events()
    k_ms = [a.elapsed_time(b) for a, b in events]
    ops.set_kernel_timer(None)
    for i in range(min(2, args.warmup)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)

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
                   "parallelism": f"dp{world}", "steps_per_s": args.steps / (ms * 1e-3)},
        "e2e": {"value": vox / (ms_e2e / args.steps * 1e-3), "unit": "HR voxels/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": {"bound": "tensor", "kernel": "conv3d_tc_kernel (hr_convs.0 fwd, 5x5x5 144->144 @128x128x10, B=8)",
                     "achieved": achieved, "peak": tflops_peak, "unit": "TFLOP/s",
                     "frac": achieved / tflops_peak if tflops_peak else None, "traffic": None,
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
