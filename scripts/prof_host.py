"""Host-side cost of one training step with the GPU work made negligible (tiny volume, same architecture):
what remains is Python + launch overhead.  cProfile over 5 steps."""
import cProfile, os, pstats, sys, io, time
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
hr = 128  # the discriminator head is sized for 128x128; B=1 keeps the GPU far ahead of the host
LR, HR, Z, x, y = make_batch(1, hr, 10, 8, seed=2001, device=dev)
t = cfg.training
gan.feed_xy_niter(x, y, torch.tensor(t.niter, device=dev), t.d_g_train_ratio, t.d_g_train_period)
for i in range(3):
    gan.optimize_parameters(LR, HR, Z, 1 + i)
torch.cuda.synchronize()
l0 = ops.launch_count()
t0 = time.perf_counter()
for i in range(10):
    gan.optimize_parameters(LR, HR, Z, 5 + i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e2 * (t1 - t0):.1f} ms/step, wall {1e2 * (t2 - t0):.1f} ms/step, launches/step {(ops.launch_count() - l0) / 10}")
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    gan.optimize_parameters(LR, HR, Z, 5 + i)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
print(s.getvalue()[:9000])
