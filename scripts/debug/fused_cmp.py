"""Debug: one G step with fused vs foreach Adam from identical state; compare every parameter/buffer."""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from tests.util import load_npz, sd_from, rel_l2, GOLDEN
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.config.config import Config
from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
z = load_npz("gan_step.npz")
def run(fused):
    os.environ["WINDSR_FUSED_ADAM"] = "1" if fused else "0"
    cfg = Config(os.path.join(GOLDEN, "configs", "tiny_gan.ini"))
    cfg.is_train, cfg.gpu_id, cfg.device = True, 0, torch.device("cuda:0")
    gan = wind_field_GAN_3D(cfg)
    gan.G.load_state_dict(sd_from(z, "G0/")); gan.D.load_state_dict(sd_from(z, "D0/"))
    LR, HR, Z, x, y = (torch.from_numpy(z[k]).cuda() for k in ("LR", "HR", "Z", "x", "y"))
    gan.feed_xy_niter(x, y, torch.tensor(cfg.training.niter, device="cuda"), cfg.training.d_g_train_ratio, cfg.training.d_g_train_period)
    with ops.precision("fp32"):
        gan.optimize_parameters(LR, HR, Z, 1)
    return gan
a, b = run(True), run(False)
for nm in ("G", "D"):
    sa, sb = getattr(a, nm).state_dict(), getattr(b, nm).state_dict()
    for k in sa:
        d = (sa[k].float() - sb[k].float()).abs().max().item()
        if d > 1e-7:
            p = dict(getattr(a, nm).named_parameters()).get(k)
            print(nm, k, d, tuple(sa[k].shape), sa[k].stride(), None if p is None or p.grad is None else p.grad.stride())
print("done")
