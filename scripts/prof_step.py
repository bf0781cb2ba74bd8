"""Kernel-time breakdown of one training step via torch.profiler (CUPTI; cheap compared with ncu).
Groups by (kernel, grid) and prints the top entries; writes gpurun_out/step_kernels.json."""
import collections, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.config.config import Config
from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
from gan_sr_wind_field_b200.synthetic import make_batch

dev = torch.device("cuda:0")
ops.set_precision("bf16")
# usage: prof_step.py [g|d]   (d: the D step of the full GAN schedule, BASELINE config #3)
WHICH = sys.argv[1] if len(sys.argv) > 1 else "g"
cfg = Config(bench.INI)
cfg.is_train, cfg.gpu_id, cfg.device = True, 0, dev
if WHICH == "d":
    cfg.training.adversarial_loss_weight = 0.0005
    cfg.training.d_g_train_ratio, cfg.training.d_g_train_period = 1, 1
torch.manual_seed(2001)
gan = wind_field_GAN_3D(cfg)
LR, HR, Z, x, y = make_batch(8, 128, 10, 8, seed=2001, device=dev)
t = cfg.training
gan.feed_xy_niter(x, y, torch.tensor(t.niter, device=dev), t.d_g_train_ratio, t.d_g_train_period)
IT = 5 if WHICH == "d" else 4   # with period 1 / ratio 1 odd iterations are D steps
for i in range(4):
    gan.optimize_parameters(LR, HR, Z, i)
torch.cuda.synchronize()
import time
for i in range(3):
    t0 = time.perf_counter()
    gan.optimize_parameters(LR, HR, Z, IT)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"step: CPU enqueue {1e3*(t1-t0):.1f} ms (includes the one host read before backward), total {1e3*(t2-t0):.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    gan.optimize_parameters(LR, HR, Z, IT)
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "step_trace.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
tr = json.load(open(path))
# kernel timeline (start, duration, stream) of the step, for offline phase / gap analysis
tl = []
for e in tr["traceEvents"]:
    if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
        nm = re.sub(r"ws::[^:]*::", "", e["name"].replace("void ", ""))
        nm = re.sub(r"\(.*", "", nm)[:60]
        tl.append((e["ts"], e["dur"], e.get("args", {}).get("stream", -1), nm, "x".join(map(str, e.get("args", {}).get("grid", [])))))
tl.sort()
t0_ = tl[0][0] if tl else 0
with open(os.path.join(ROOT, "gpurun_out", os.environ.get("TIMELINE_NAME", "step_timeline.csv")), "w") as fh:
    fh.write("start_us,dur_us,stream,kernel,grid\n")
    for ts, dur, st_, nm, grid in tl:
        fh.write(f"{ts - t0_:.1f},{dur:.1f},{st_},{nm},{grid}\n")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
tmin, tmax = 1e30, 0
for e in tr["traceEvents"]:
    if e.get("cat") == "kernel":
        name = e["name"].replace("void ", "")
        name = re.sub(r"ws::[^:]*::", "", name)          # ws::<unnamed>:: / ws::(anonymous namespace)::
        name = re.sub(r"\(.*", "", name)
        m_ = re.match(r"([A-Za-z_0-9:]+)(<[^>]*>)?", name)
        name = (m_.group(1).split("::")[-1] + (m_.group(2) or ""))[:44] if m_ else name[:44]
        grid = tuple(e.get("args", {}).get("grid", []))
        agg[(name, grid)][0] += 1
        agg[(name, grid)][1] += e["dur"]
        tot += e["dur"]
        tmin = min(tmin, e["ts"]); tmax = max(tmax, e["ts"] + e["dur"])
print(f"kernels: {sum(v[0] for v in agg.values())}, sum of kernel time {tot/1e3:.1f} ms, GPU span {(tmax-tmin)/1e3:.1f} ms")
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
for (name, grid), (n, us) in rows[:45]:
    print(f"{name:42s} {str(grid):18s} n={n:4d} {us/1e3:8.3f} ms {100*us/tot:5.1f}%")
json.dump([dict(kernel=k[0], grid=list(k[1]), launches=v[0], ms=v[1] / 1e3) for k, v in rows],
          open(os.path.join(ROOT, "gpurun_out", "step_kernels.json" if WHICH == "g" else "dstep_kernels.json"), "w"), indent=1)
os.remove(path)
