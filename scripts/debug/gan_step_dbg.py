"""Debug: golden GAN step in fp32, report where the D-step loss deviates."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from tests.util import load_npz, sd_from, rel_l2, GOLDEN
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.config.config import Config
from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
z = load_npz("gan_step.npz")
cfg = Config(os.path.join(GOLDEN, "configs", "tiny_gan.ini"))
cfg.is_train, cfg.gpu_id, cfg.device = True, 0, torch.device("cuda:0")
gan = wind_field_GAN_3D(cfg)
gan.G.load_state_dict(sd_from(z, "G0/")); gan.D.load_state_dict(sd_from(z, "D0/"))
LR, HR, Z, x, y = (torch.from_numpy(z[k]).cuda() for k in ("LR", "HR", "Z", "x", "y"))
gan.feed_xy_niter(x, y, torch.tensor(cfg.training.niter, device="cuda"), cfg.training.d_g_train_ratio, cfg.training.d_g_train_period)
with ops.precision("fp32"):
    gan.optimize_parameters(LR, HR, Z, 1)
    print("found_inf", getattr(gan.optimizer_G, "found_inf", None))
    pg = dict(gan.G.named_parameters())
    worst = 0
    for k in z.files:
        if k.startswith("G_step/param/"):
            name = k[13:]
            upd = pg[name].detach().cpu() - torch.from_numpy(z[f"G0/{name}"])
            ref_upd = torch.from_numpy(z[k]) - torch.from_numpy(z[f"G0/{name}"])
            r = rel_l2(upd, ref_upd)
            if r > worst:
                worst = r; print(name, r, float(upd.abs().max()), float(ref_upd.abs().max()))
    print("G.training", gan.G.training, "D.training", gan.D.training, [m.training for m in gan.D.modules()][:5])
    gan.optimize_parameters(LR, HR, Z, 2)
    print("D loss", float(gan.get_D_loss_dict_ref()["train_loss"]), float(z["D_step/loss"]))
    print("G.training", gan.G.training, "D.training", gan.D.training)
