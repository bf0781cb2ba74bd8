"""Debug: 3 data-parallel G steps; prints checksums of parameters and last gradients (compare runs with
WINDSR_AUX_STREAM=1 / 0 and ranks with each other)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.config.config import Config
from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
from gan_sr_wind_field_b200.synthetic import make_batch
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
ops.set_precision("bf16")
cfg = Config(bench.INI); cfg.is_train, cfg.gpu_id, cfg.device = True, local, dev
cfg.generator.dropout_probability = 0.0 if hasattr(cfg, "generator") else None
torch.manual_seed(2001)
gan = wind_field_GAN_3D(cfg)
LR, HR, Z, x, y = make_batch(4, 128, 10, 8, seed=2001 + rank, device=dev)
t = cfg.training
t.use_instance_noise = False
gan.feed_xy_niter(x, y, torch.tensor(t.niter, device=dev), t.d_g_train_ratio, t.d_g_train_period)
for i in range(3):
    gan.optimize_parameters(LR, HR, Z, 1 + i)
torch.cuda.synchronize()
ps = list(gan.G.parameters())
psum = sum(float(p.detach().double().abs().sum()) for p in ps)
gsum = sum(float(p.grad.double().abs().sum()) for p in ps if p.grad is not None)
g_rdb = float(dict(gan.G.named_parameters())["model.1.module.7.RDBs.1.conv2.conv.0.weight"].grad.double().norm())
g_lff = float(dict(gan.G.named_parameters())["model.1.module.7.RDBs.1.LFF.weight"].grad.double().norm())
print(f"[rank {rank}] aux={os.environ.get('WINDSR_AUX_STREAM','1')} |params| {psum:.6f} |grads| {gsum:.6e} "
      f"rdb conv2 grad {g_rdb:.6e} LFF grad {g_lff:.6e} loss {float(gan.get_G_train_loss_dict_ref()['total']):.6f}", flush=True)
dist.destroy_process_group()
