"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference through
oracle/refshim.py) on seeded synthetic inputs.  The reference ships no golden vectors of its own;
these files are the pin for oracle/wind_oracle.py and, through it, for the CUDA path.

    python tests/golden/make_golden.py        # needs /root/reference; rewrites the .npz files
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim, wind_oracle as wo  # noqa: E402

refshim.activate()
from CNN_models.Discriminator_3D import Discriminator_3D  # noqa: E402
from CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D  # noqa: E402
import config.config as ref_config  # noqa: E402
import process_data  # noqa: E402
import tools.initialization as init  # noqa: E402
from GAN_models.wind_field_GAN_3D import get_norm_factors_of_gradients, wind_field_GAN_3D  # noqa: E402


def npd(sd, prefix):
    return {f"{prefix}{k}": v.detach().clone().numpy() for k, v in sd.items()}


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(arrays)} arrays")


def golden_generator():
    torch.manual_seed(11)
    G = Generator_3D(4, 3, 16, 2, upscale=4, hr_kern_size=5, number_of_RDB_convs=5, RDB_gc=8, lff_kern_size=1,
                     terrain_number_of_features=8, dropout_probability=0.0)
    init.init_weights(G, 0.5)
    G.train()
    LR, HR, Z, x, y = wo.synthetic_batch(2, hr_xy=16, nz=10, scale=4, seed=5)
    LR.requires_grad_(True)
    out = G(LR, Z)
    r = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    loss = (out * r).sum()
    names = ["model.0.0.weight", "model.1.module.0.RDBs.1.conv2.conv.0.weight", "model.1.module.1.RDBs.2.LFF.weight",
             "model.1.module.1.RDBs.2.LFF.bias", "model.1.module.2.0.weight", "model.3.1.0.weight",
             "terrain_convs.0.0.weight", "terrain_convs.1.0.weight", "hr_convs.0.0.weight", "hr_convs.2.weight",
             "hr_convs.2.bias"]
    params = dict(G.named_parameters())
    grads = torch.autograd.grad(loss, [LR] + [params[n] for n in names])
    arrays = npd(G.state_dict(), "sd/")
    arrays.update(LR=LR.detach().numpy(), Z=Z.numpy(), r=r.numpy(), out=out.detach().numpy(),
                  grad_LR=grads[0].numpy())
    for n, g in zip(names, grads[1:]):
        arrays[f"grad/{n}"] = g.numpy()
    save("generator_small.npz", **arrays)


def golden_discriminator():
    for slicing, hx, tag in ((True, 64, "slicing"), (False, 128, "full")):
        torch.manual_seed(13)
        D = Discriminator_3D(3, 4, enable_slicing=slicing, dropout_probability=0.0)
        init.init_weights(D, 0.2)
        xin = torch.randn(1 if hx == 128 else 2, 3, hx, hx, 10,
                          generator=torch.Generator().manual_seed(7)).half().float()
        sd0 = {k: v.clone() for k, v in D.state_dict().items()}
        D.train()
        xin.requires_grad_(True)
        out = D(xin)
        names = ["features.0.0.0.weight", "features.1.1.0.weight", "features.1.1.1.weight", "features.1.1.1.bias",
                 "features.3.1.0.weight", "classifier.2.weight"]
        params = dict(D.named_parameters())
        grads = torch.autograd.grad(out.sum(), [xin] + [params[n] for n in names])
        D.eval()
        out_eval = D(xin.detach())
        arrays = npd(sd0, "sd/")
        arrays.update(npd({k: v for k, v in D.state_dict().items() if "running" in k}, "after/"))
        arrays.update(x=xin.detach().numpy().astype(np.float16), out_train=out.detach().numpy(),
                      out_eval_after=out_eval.detach().numpy(), grad_x_sub=grads[0][:, :, ::4, ::4, :].numpy())
        for n, g in zip(names, grads[1:]):
            arrays[f"grad/{n}"] = g.numpy()
        save(f"discriminator_{tag}.npz", **arrays)


def golden_windloss():
    LR, HR, Z, x, y = wo.synthetic_batch(2, hr_xy=12, nz=10, scale=4, seed=9)
    HR = HR[:, :, :, :10].contiguous()  # X = 12, Y = 10: non-square on purpose
    Z = Z[:, :, :, :10].contiguous()
    y = y[:10].contiguous()
    gen = torch.Generator().manual_seed(21)
    SR = (HR + 0.1 * torch.randn(HR.shape, generator=gen)).requires_grad_(True)
    gh = process_data.calculate_gradient_of_wind_field(HR, x, y, Z)
    gs = process_data.calculate_gradient_of_wind_field(SR, x, y, Z)
    norms = get_norm_factors_of_gradients(gh, gs)
    mse = torch.nn.MSELoss()
    terms = [torch.nn.L1Loss()(HR, SR),
             mse(gs[:, :6] / norms[0], gh[:, :6] / norms[0]),
             mse(gs[:, 6:] / norms[1], gh[:, 6:] / norms[1]),
             mse((gh[:, 0] + gh[:, 4] + gh[:, 8]) / norms[2], (gs[:, 0] + gs[:, 4] + gs[:, 8]) / norms[2]),
             mse((gh[:, 0] + gh[:, 4]) / norms[3], (gs[:, 0] + gs[:, 4]) / norms[3])]
    w = [0.136, 3.064, 0.2, 0.366, 0.721]
    total = sum(wi * ti for wi, ti in zip(w, terms))
    (dsr,) = torch.autograd.grad(total, SR)
    # second case: SR gradients 150x larger so the normalisers take the SR_max/100 branch
    SR2 = (HR + 150.0 * torch.randn(HR.shape, generator=gen)).requires_grad_(True)
    gs2 = process_data.calculate_gradient_of_wind_field(SR2, x, y, Z)
    norms2 = get_norm_factors_of_gradients(gh, gs2)
    terms2 = [torch.nn.L1Loss()(HR, SR2),
              mse(gs2[:, :6] / norms2[0], gh[:, :6] / norms2[0]),
              mse(gs2[:, 6:] / norms2[1], gh[:, 6:] / norms2[1]),
              mse((gh[:, 0] + gh[:, 4] + gh[:, 8]) / norms2[2], (gs2[:, 0] + gs2[:, 4] + gs2[:, 8]) / norms2[2]),
              mse((gh[:, 0] + gh[:, 4]) / norms2[3], (gs2[:, 0] + gs2[:, 4]) / norms2[3])]
    total2 = sum(wi * ti for wi, ti in zip(w, terms2))
    (dsr2,) = torch.autograd.grad(total2, SR2)
    save("windloss.npz", HR=HR.numpy(), SR=SR.detach().numpy(), Z=Z.numpy(), x=x.numpy(), y=y.numpy(),
         jac_HR=gh.numpy(), jac_SR=gs.detach().numpy(), norms=np.array([float(n) for n in norms]),
         terms=np.array([float(t) for t in terms]), weights=np.array(w), total=np.array(float(total)),
         dSR=dsr.numpy(), SR2=SR2.detach().numpy(), norms2=np.array([float(n) for n in norms2]),
         terms2=np.array([float(t) for t in terms2]), total2=np.array(float(total2)), dSR2=dsr2.numpy())


def golden_gan_step():
    cfg = ref_config.Config(os.path.join(HERE, "configs", "tiny_gan.ini"))
    cfg.is_train = True
    cfg.gpu_id = None
    cfg.device = torch.device("cpu")
    torch.manual_seed(2001)
    gan = wind_field_GAN_3D(cfg)
    LR, HR, Z, x, y = wo.synthetic_batch(2, hr_xy=64, nz=10, scale=4, seed=17)
    gan.feed_xy_niter(x, y, torch.tensor(cfg.training.niter), cfg.training.d_g_train_ratio,
                      cfg.training.d_g_train_period)
    arrays = npd(gan.G.state_dict(), "G0/")
    arrays.update(npd(gan.D.state_dict(), "D0/"))
    arrays.update(LR=LR.numpy(), HR=HR.numpy().astype(np.float32), Z=Z.numpy(), x=x.numpy(), y=y.numpy())
    watch_G = ["model.0.0.weight", "model.1.module.1.RDBs.0.conv1.conv.0.weight", "hr_convs.2.weight",
               "hr_convs.2.bias", "terrain_convs.0.0.weight"]
    watch_D = ["features.0.0.0.weight", "features.2.1.1.weight", "classifier.2.weight"]
    # it = 1 -> G step (train_period 0); it = 2 -> D step (train_period 1)
    gan.optimize_parameters(LR, HR, Z, 1)
    for k, v in gan.get_G_train_loss_dict_ref().items():
        arrays[f"G_step/loss/{k}"] = np.array(float(v))
    pg = dict(gan.G.named_parameters())
    for n in watch_G:
        arrays[f"G_step/grad/{n}"] = pg[n].grad.numpy().copy()
        arrays[f"G_step/param/{n}"] = pg[n].detach().numpy().copy()
    gan.optimize_parameters(LR, HR, Z, 2)
    arrays["D_step/loss"] = np.array(float(gan.get_D_loss_dict_ref()["train_loss"]))
    pd_ = dict(gan.D.named_parameters())
    for n in watch_D:
        arrays[f"D_step/grad/{n}"] = pd_[n].grad.numpy().copy()
        arrays[f"D_step/param/{n}"] = pd_[n].detach().numpy().copy()
    arrays.update(npd({k: v for k, v in gan.D.state_dict().items() if "running" in k}, "D_step/after/"))
    save("gan_step.npz", **arrays)


if __name__ == "__main__":
    torch.set_num_threads(8)
    golden_generator()
    golden_discriminator()
    golden_windloss()
    golden_gan_step()
