"""CPU-side (Python / launch) profile of one training step: cProfile over 5 steps."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.config.config import Config
from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
from gan_sr_wind_field_b200.synthetic import make_batch
dev = torch.device("cuda:0")
ops.set_precision("bf16")
cfg = Config(bench.INI); cfg.is_train, cfg.gpu_id, cfg.device = True, 0, dev
torch.manual_seed(2001)
gan = wind_field_GAN_3D(cfg)
LR, HR, Z, x, y = make_batch(8, 128, 10, 8, seed=2001, device=dev)
t = cfg.training
gan.feed_xy_niter(x, y, torch.tensor(t.niter, device=dev), t.d_g_train_ratio, t.d_g_train_period)
for i in range(3):
    gan.optimize_parameters(LR, HR, Z, 1 + i)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(5):
    gan.optimize_parameters(LR, HR, Z, 5 + i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / 5:.1f} ms/step, wall {1e3 * (t2 - t0) / 5:.1f} ms/step (no profiler)")
st0 = torch.cuda.memory_stats()
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    gan.optimize_parameters(LR, HR, Z, 5 + i)
    torch.cuda.synchronize()
pr.disable()
st1 = torch.cuda.memory_stats()
for k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "allocation.all.allocated", "segment.all.allocated"):
    print(k, st1.get(k, 0) - st0.get(k, 0))
print("reserved GB", st1["reserved_bytes.all.current"] / 1e9, "peak allocated GB", st1["allocated_bytes.all.peak"] / 1e9)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
