"""Four G training steps of the bench workload (upscale8, B=8, bf16), nothing else: the target of the ncu launch list."""
import os, sys
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for i in range(n):
    if i == n - 1:  # `ncu --profile-from-start off` then sees exactly one steady-state step
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    gan.optimize_parameters(LR, HR, Z, 1 + i)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", ops.launch_count())
