"""Debug (not shipped): D train-mode fwd/bwd in fp32 on the GPU vs the CPU oracle, per parameter."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from tests.util import load_npz, sd_from, rel_l2, GOLDEN
from oracle import wind_oracle as wo
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.config.config import Config
from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
z = load_npz("gan_step.npz")
cfg = Config(os.path.join(GOLDEN, "configs", "tiny_gan.ini"))
cfg.is_train, cfg.gpu_id, cfg.device = True, 0, torch.device("cuda:0")
gan = wind_field_GAN_3D(cfg)
gan.D.load_state_dict(sd_from(z, "D0/"))
HR = torch.from_numpy(z["HR"])
torch.manual_seed(0)
fake = HR + 0.05 * torch.randn_like(HR)
sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k and "num_batches" not in k) for k, v in sd_from(z, "D0/").items()}
yp = wo.discriminator_forward(sd, HR, True).squeeze(); fy = wo.discriminator_forward(sd, fake, True).squeeze()
lab = torch.full_like(yp, 0.9); fl = torch.zeros_like(yp)
loss = wo.discriminator_loss(yp, fy, lab, fl); loss.backward()
D = gan.D; D.train()
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
with ops.precision(mode):
    a = D(HR.cuda()).squeeze(); b = D(fake.cuda()).squeeze()
    print("fwd", rel_l2(a, yp), rel_l2(b, fy))
    l2 = wo.discriminator_loss(a, b, lab.cuda(), fl.cuda()); l2.backward()
print("loss", float(loss), float(l2))
for k, p in D.named_parameters():
    if p.grad is not None and sd[k].grad is not None:
        print(f"{k:40s} {rel_l2(p.grad, sd[k].grad):.3e}  |g|={float(sd[k].grad.norm()):.3e}")
