"""GPU parity of the drop-in modules, the fused wind loss and the GAN step against the golden fixtures produced
by the unmodified reference (tests/golden/make_golden.py) and against the CPU oracle at full size.
Tolerances are the north star's: rel-L2 <= 1e-5 (FP32 mode) / 2e-2 (BF16 mode); where a quantity is the result
of a long fp32 reduction chain re-associated by the GPU (gradients through ~50 layers, BatchNorm statistics) the
FP32 bound is stated next to the assertion."""
import os

import numpy as np
import pytest
import torch

from tests.util import (GOLDEN, TOL, bf16_operand_emulation, load_npz, rel_l2, sd_from,
                        small_generator_kwargs)

pytestmark = pytest.mark.gpu


def _generator(sd):
    from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    G = Generator_3D(**small_generator_kwargs()).cuda()
    G.load_state_dict(sd)
    return G


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_generator_forward_backward_vs_golden(mode):
    from gan_sr_wind_field_b200 import ops
    z = load_npz("generator_small.npz")
    G = _generator(sd_from(z, "sd/"))
    G.train()
    LR = torch.from_numpy(z["LR"]).cuda().requires_grad_(True)
    Z = torch.from_numpy(z["Z"]).cuda()
    with ops.precision(mode):
        out = G(LR, Z)
        assert out.shape == z["out"].shape and out.dtype == torch.float32 and out.is_contiguous()
        (out * torch.from_numpy(z["r"]).cuda()).sum().backward()
    tol = TOL[mode]
    assert rel_l2(out, z["out"]) <= tol
    params = dict(G.named_parameters())
    errs = {k[5:]: rel_l2(params[k[5:]].grad, z[k]) for k in z.files if k.startswith("grad/")}
    errs["LR"] = rel_l2(LR.grad, z["grad_LR"])
    if mode == "fp32":
        assert all(e <= 5e-5 for e in errs.values()), errs  # ~50-layer chain of re-associated fp32 sums
        return
    # BF16 / TF32: this fixture is deliberately hot (init scale 0.5, |SR| up to ~90, raw-altitude terrain features
    # next to O(1) wind features), so even the minimal bf16-operand scheme is 5-19 % away from fp32 in the deep
    # gradients.  Bound the CUDA path by that intrinsic envelope, measured here on the CPU.
    if mode == "tf32":
        from tests.parity_util import operand_rounding
        emulate = lambda: operand_rounding("tf32")
    else:
        emulate = bf16_operand_emulation
    from oracle import wind_oracle as wo
    sd = sd_from(z, "sd/")
    p_cpu = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lr_cpu = torch.from_numpy(z["LR"]).requires_grad_(True)
    names = [k for k in errs if k != "LR"]
    with emulate():
        o = wo.generator_forward(p_cpu, lr_cpu, torch.from_numpy(z["Z"]))
        g = torch.autograd.grad((o * torch.from_numpy(z["r"])).sum(), [lr_cpu] + [p_cpu[n] for n in names])
    env = {"LR": rel_l2(g[0], z["grad_LR"])}
    env.update({n: rel_l2(gi, z[f"grad/{n}"]) for n, gi in zip(names, g[1:])})
    assert all(errs[k] <= 1.5 * env[k] + tol for k in errs), (errs, env)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["slicing", "full"])
def test_discriminator_vs_golden(tag, mode):
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.CNN_models.Discriminator_3D import Discriminator_3D
    z = load_npz(f"discriminator_{tag}.npz")
    D = Discriminator_3D(3, 4, enable_slicing=(tag == "slicing"), dropout_probability=0.0).cuda()
    D.load_state_dict(sd_from(z, "sd/"))
    x = torch.from_numpy(z["x"].astype(np.float32)).cuda().requires_grad_(True)
    tol = TOL[mode]
    with ops.precision(mode):
        D.train()
        out = D(x)
        out.sum().backward()
        assert rel_l2(out, z["out_train"]) <= (1e-4 if mode == "fp32" else 5 * tol)  # 10 BatchNorms in series
        sd = D.state_dict()
        for k in z.files:
            if k.startswith("after/"):
                assert rel_l2(sd[k[6:]], z[k]) <= (1e-5 if mode == "fp32" else tol), k
        params = dict(D.named_parameters())
        errs = {k[5:]: rel_l2(params[k[5:]].grad, z[k]) for k in z.files if k.startswith("grad/")}
        errs["x"] = rel_l2(x.grad[:, :, ::4, ::4, :], z["grad_x_sub"])
        if mode == "fp32":
            # fp32 mode: weight gradients are 1.6e5-term sums with heavy cancellation accumulated by split-K fp32
            # atomics (order varies run to run) behind ten train-mode BatchNorms: observed 2e-6 .. 2e-4
            assert all(e <= 1e-3 for e in errs.values()), errs
        else:
            # gradients through 10 train-mode BatchNorms of a 4-feature net: bound by the intrinsic bf16 envelope
            from oracle import wind_oracle as wo
            sd0 = sd_from(z, "sd/")
            p_cpu = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
                     for k, v in sd0.items()}
            x_cpu = torch.from_numpy(z["x"].astype(np.float32)).requires_grad_(True)
            names = [k for k in errs if k != "x"]
            with bf16_operand_emulation():
                o = wo.discriminator_forward(p_cpu, x_cpu, True)
                g = torch.autograd.grad(o.sum(), [x_cpu] + [p_cpu[n] for n in names])
            env = {"x": rel_l2(g[0][:, :, ::4, ::4, :], z["grad_x_sub"])}
            env.update({n: rel_l2(gi, z[f"grad/{n}"]) for n, gi in zip(names, g[1:])})
            assert all(errs[k] <= 1.5 * env[k] + 5 * tol for k in errs), (errs, env)
        D.eval()
        with torch.no_grad():
            out_eval = D(x.detach())
        assert rel_l2(out_eval, z["out_eval_after"]) <= (1e-4 if mode == "fp32" else 5 * tol)


def test_wind_gradient_and_loss_vs_golden():
    from gan_sr_wind_field_b200 import ops
    z = load_npz("windloss.npz")
    HR, Z, x, y = (torch.from_numpy(z[k]).cuda() for k in ("HR", "Z", "x", "y"))
    # the 9-channel Jacobian itself (calculate_gradient_of_wind_field): same operation order as torch -> ~1 ulp
    assert rel_l2(ops.wind_gradient(HR, x, y, Z), z["jac_HR"]) <= 1e-6
    assert rel_l2(ops.wind_gradient(torch.from_numpy(z["SR"]).cuda(), x, y, Z), z["jac_SR"]) <= 1e-6
    w = z["weights"].tolist()
    cnt = float(HR.shape[0] * HR.shape[2] * HR.shape[3] * HR.shape[4])
    for sr_key, sfx in (("SR", ""), ("SR2", "2")):
        SR = torch.from_numpy(z[sr_key]).cuda().requires_grad_(True)
        s = ops.windloss_slots(HR, SR, Z, x, y)
        norm = [torch.maximum(s[6 + 2 * k], s[7 + 2 * k] / 100) for k in range(4)]
        assert np.allclose([float(n) for n in norm], z["norms" + sfx], rtol=1e-6)
        terms = [s[4] / (3 * cnt), s[0] / (6 * cnt) / norm[0] ** 2, s[1] / (3 * cnt) / norm[1] ** 2,
                 s[2] / cnt / norm[2] ** 2, s[3] / cnt / norm[3] ** 2]
        assert np.allclose([float(t) for t in terms], z["terms" + sfx], rtol=1e-5)
        total = sum(wi * ti for wi, ti in zip(w, terms))
        total.backward()
        assert abs(float(total) - float(z["total" + sfx])) <= 1e-5 * abs(float(z["total" + sfx]))
        assert rel_l2(SR.grad, z["dSR" + sfx]) <= 1e-5  # includes the path through SR_max/100 for SR2


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_gan_step_vs_golden(mode):
    """One G step and one D step of wind_field_GAN_3D.optimize_parameters against the reference's."""
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.config.config import Config
    from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
    z = load_npz("gan_step.npz")
    cfg = Config(os.path.join(GOLDEN, "configs", "tiny_gan.ini"))
    cfg.is_train, cfg.gpu_id, cfg.device = True, 0, torch.device("cuda:0")
    gan = wind_field_GAN_3D(cfg)
    gan.G.load_state_dict(sd_from(z, "G0/"))
    gan.D.load_state_dict(sd_from(z, "D0/"))
    LR, HR, Z, x, y = (torch.from_numpy(z[k]).cuda() for k in ("LR", "HR", "Z", "x", "y"))
    gan.feed_xy_niter(x, y, torch.tensor(cfg.training.niter, device="cuda"), cfg.training.d_g_train_ratio,
                      cfg.training.d_g_train_period)
    tol = 1e-4 if mode == "fp32" else TOL[mode]
    with ops.precision(mode):
        gan.optimize_parameters(LR, HR, Z, 1)  # G step
        for k, v in gan.get_G_train_loss_dict_ref().items():
            ref = float(z[f"G_step/loss/{k}"])
            assert abs(float(v) - ref) <= tol * max(abs(ref), 1e-3), (k, float(v), ref)
        pg = dict(gan.G.named_parameters())
        for k in z.files:
            if k.startswith("G_step/grad/"):
                assert rel_l2(pg[k[12:]].grad, z[k]) <= (2e-4 if mode == "fp32" else 2.5 * tol), k
        gan.optimize_parameters(LR, HR, Z, 2)  # D step
        ref = float(z["D_step/loss"])
        assert abs(float(gan.get_D_loss_dict_ref()["train_loss"]) - ref) <= (1e-4 if mode == "fp32" else 0.05) * abs(ref)
        if mode == "fp32":
            pd_ = dict(gan.D.named_parameters())
            # The D step sees fake_HR from the generator AFTER its first Adam step, which moves every weight by
            # ~lr*sign(g): weights whose gradient is at rounding-noise level land lr apart between any two fp32
            # implementations (update rel_l2 bound below is 5e-2 for that reason), and that input perturbation
            # shows up ~1e-2 in these gradients.  The D kernels themselves are held to 1e-4..1e-3 by
            # test_discriminator_vs_golden (measured ~3e-6 per parameter in fp32).
            for k in z.files:
                if k.startswith("D_step/grad/"):
                    assert rel_l2(pd_[k[12:]].grad, z[k]) <= 2e-2, k
            # Adam's first step moves every weight by ~lr*sign(g): compare the update, not the tiny gradients
            for k in z.files:
                if k.startswith("G_step/param/"):
                    name = k[13:]
                    upd = pg[name].detach().cpu() - torch.from_numpy(z[f"G0/{name}"])
                    ref_upd = torch.from_numpy(z[k]) - torch.from_numpy(z[f"G0/{name}"])
                    assert rel_l2(upd, ref_upd) <= 0.05, name


@pytest.mark.parametrize("mode", ["bf16", "tf32", "fp32"])
def test_full_size_upscale8_generator_vs_oracle(mode):
    """BASELINE.json config #1 shapes: the shipped upscale8 architecture (128 features, 16 RRDBs, 5x5x5 HR convs),
    LR (1,4,16,16,10) -> SR (1,3,128,128,10), against the CPU oracle with the same weights."""
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    from gan_sr_wind_field_b200.tools import initialization
    from oracle import wind_oracle as wo
    torch.manual_seed(2001)
    G = Generator_3D(4, 3, 128, 16, upscale=8, hr_kern_size=5, lff_kern_size=1, dropout_probability=0.1)
    initialization.init_weights(G, 0.1)
    G.eval()
    LR, HR, Z, x, y = wo.synthetic_batch(1, hr_xy=128, nz=10, scale=8, seed=2001)
    with torch.no_grad():
        ref = wo.generator_forward(G.state_dict(), LR, Z)
    G.cuda()
    with ops.precision(mode), torch.no_grad():
        out = G(LR.cuda(), Z.cuda())
    assert out.shape == (1, 3, 128, 128, 10)
    assert rel_l2(out, ref) <= TOL[mode]


def test_size_independent_properties_full_size():
    """At the full 144-channel 128x128x10 shape (too slow for the CPU oracle in backward): linearity of the
    tensor-core conv in its input, and <dy, conv(x)> == <dgrad(dy), x> == <wgrad(x, dy), w> (adjointness)."""
    from gan_sr_wind_field_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    n, c, vol = 1, 144, (128, 128, 10)
    with ops.precision("bf16"):
        mk = lambda *s: torch.randn(*s, generator=g, device="cuda")
        xa = ops.empty_cl(n, c, *vol, torch.bfloat16, "cuda")
        xb = ops.empty_cl(n, c, *vol, torch.bfloat16, "cuda")
        xa.copy_(mk(n, c, *vol))
        xb.copy_(mk(n, c, *vol))
        w = (mk(c, c, 5, 5, 5) / (c * 125) ** 0.5).bfloat16().float()  # bf16-exact weights
        shape = ops.make_shape(xa.shape, c, (5, 5, 5), 1, 2)
        ya, yb, yab = (ops.empty_cl(n, c, *vol, torch.float32, "cuda") for _ in range(3))
        ops.conv_fwd(xa, w, None, shape, ya)
        ops.conv_fwd(xb, w, None, shape, yb)
        xs = ops.empty_cl(n, c, *vol, torch.bfloat16, "cuda")
        xs.copy_((xa.float() + xb.float()) * 0.5)  # may round; compare against the rounded sum
        ops.conv_fwd(xs, w, None, shape, yab)
        xs_exact = xs.float()
        lin = rel_l2(yab, (ya + yb) * 0.5)
        assert lin <= 1e-2, lin
        dy = ops.empty_cl(n, c, *vol, torch.bfloat16, "cuda")
        dy.copy_(mk(n, c, *vol))
        dx = ops.empty_cl(n, c, *vol, torch.float32, "cuda")
        ops.conv_dgrad(dy, w, None, shape, dx)
        dw, _ = ops.conv_wgrad(xa, dy, shape)
        lhs = float((dy.float() * ya).sum())
        assert abs(float((dx * xa.float()).sum()) - lhs) <= 2e-3 * abs(lhs)
        assert abs(float((dw * w).sum()) - lhs) <= 2e-3 * abs(lhs)


def test_bf16_gradients_at_shipped_init_scale():
    """BF16 parameter gradients of a mid-size generator at the SHIPPED weight-init scale (0.1) and input ranges,
    against fp32 CPU autograd on the oracle: the north star's 2e-2 where the net is in its operating regime."""
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    from gan_sr_wind_field_b200.tools import initialization
    from oracle import wind_oracle as wo
    torch.manual_seed(7)
    G = Generator_3D(4, 3, 64, 3, upscale=4, hr_kern_size=5, number_of_RDB_convs=5, RDB_gc=32, lff_kern_size=1,
                     terrain_number_of_features=16, dropout_probability=0.0)
    initialization.init_weights(G, 0.1)
    LR, HR, Z, x, y = wo.synthetic_batch(2, hr_xy=32, nz=10, scale=4, seed=3)
    p_cpu = {k: v.detach().clone().requires_grad_(True) for k, v in G.state_dict().items()}
    out_ref = wo.generator_forward(p_cpu, LR, Z)
    names = ["model.0.0.weight", "model.1.module.1.RDBs.1.conv3.conv.0.weight", "model.1.module.2.RDBs.2.LFF.weight",
             "model.1.module.3.0.weight", "model.2.1.0.weight", "terrain_convs.1.0.weight", "hr_convs.0.0.weight",
             "hr_convs.2.weight"]
    g_ref = torch.autograd.grad(torch.nn.functional.l1_loss(out_ref, HR), [p_cpu[n] for n in names])
    G.cuda().train()
    with ops.precision("bf16"):
        out = G(LR.cuda(), Z.cuda())
        torch.nn.functional.l1_loss(out, HR.cuda()).backward()
    assert rel_l2(out, out_ref) <= TOL["bf16"]
    params = dict(G.named_parameters())
    errs = {n: rel_l2(params[n].grad, g) for n, g in zip(names, g_ref)}
    with bf16_operand_emulation():
        p2 = {k: v.detach().clone().requires_grad_(True) for k, v in G.cpu().state_dict().items()}
        o2 = wo.generator_forward(p2, LR, Z)
        g2 = torch.autograd.grad(torch.nn.functional.l1_loss(o2, HR), [p2[n] for n in names])
    env = {n: rel_l2(a, b) for n, a, b in zip(names, g2, g_ref)}
    assert all(errs[n] <= max(TOL["bf16"], 1.5 * env[n] + 5e-3) for n in names), (errs, env)


def test_nan_guard_drops_physics_terms_like_the_reference():
    """Degenerate altitude levels (two equal z levels -> division by zero in calculate_div_z) make the physics
    terms NaN/Inf; the reference then optimises adversarial + pixel (+ feature) only
    (wind_field_GAN_3D.py:434-454).  The device-side guard must give a finite total equal to the oracle's and
    finite parameter gradients, and still take the optimiser step."""
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.config.config import Config
    from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
    from oracle import wind_oracle as wo
    z = load_npz("gan_step.npz")
    cfg = Config(os.path.join(GOLDEN, "configs", "tiny_gan.ini"))
    cfg.is_train, cfg.gpu_id, cfg.device = True, 0, torch.device("cuda:0")
    gan = wind_field_GAN_3D(cfg)
    G0, D0 = sd_from(z, "G0/"), sd_from(z, "D0/")
    gan.G.load_state_dict(G0)
    gan.D.load_state_dict(D0)
    LR, HR, Z, x, y = (torch.from_numpy(z[k]) for k in ("LR", "HR", "Z", "x", "y"))
    Z = Z.clone()
    Z[..., 4] = Z[..., 3]  # zero spacing between two levels
    gan.feed_xy_niter(x.cuda(), y.cuda(), torch.tensor(100, device="cuda"), 1, 2)
    w0 = gan.G.hr_convs[2].weight.detach().clone()
    with ops.precision("fp32"):
        gan.optimize_parameters(LR.cuda(), HR.cuda(), Z.cuda(), 1)
    losses = {k: float(v) for k, v in gan.get_G_train_loss_dict_ref().items()}
    assert not np.isfinite(losses["z_gradient"]) or not np.isfinite(losses["divergence"])
    # oracle with the reference's guard
    SR = wo.generator_forward(G0, LR, Z)
    real = torch.full((2,), 0.9 + 0.1 * 1 / 100)
    adv = wo.adversarial_G(wo.discriminator_forward(D0, HR, False).squeeze(),
                           wo.discriminator_forward(D0, SR, False).squeeze(), real, torch.zeros(2))
    total, _ = wo.generator_loss(HR, SR, Z, x, y, dict(pixel=0.136, xy=3.064, z=0.2, div=0.366, dxy=0.721, adv=0.05),
                                 adv=adv)
    assert np.isfinite(losses["total"]) and abs(losses["total"] - float(total)) <= 1e-4 * abs(float(total))
    assert all(torch.isfinite(p.grad).all() for p in gan.G.parameters() if p.grad is not None)
    assert not torch.equal(gan.G.hr_convs[2].weight.detach(), w0)  # the step was taken


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_packed_weights_follow_fused_optimizer(mode):
    """Single-kernel optimizers update parameters without bumping Tensor._version: the packed operand copies of
    Conv3d and of the RDB executor must still be refreshed (ops.invalidate_packed_weights via the global
    optimizer-step hook)."""
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.CNN_models.torch_blocks import RRDB
    torch.manual_seed(5)
    blk = RRDB(32, 16, 3, 1, lrelu_negative_slope=0.2, RDB_residual_scaling=0.2, RRDB_residual_scaling=0.2,
               mode="3D").cuda()
    x = torch.randn(1, 32, 8, 8, 4, device="cuda")
    opt = torch.optim.Adam(blk.parameters(), lr=0.05, fused=True)
    with ops.precision(mode):
        y0 = blk(x)
        y0.square().mean().backward()
        opt.step()
        with torch.no_grad():
            y1 = blk(x).float().clone()
        fresh = RRDB(32, 16, 3, 1, lrelu_negative_slope=0.2, RDB_residual_scaling=0.2, RRDB_residual_scaling=0.2,
                     mode="3D").cuda()
        fresh.load_state_dict(blk.state_dict())
        with torch.no_grad():
            y2 = fresh(x).float()
    assert rel_l2(y1, y0.detach().float()) > 1e-3  # the step moved the output
    assert rel_l2(y1, y2) <= 1e-6  # and the cached operands saw it


@pytest.mark.gpu
def test_full_size_upscale16_generator_vs_oracle():
    """BASELINE.json config #4 shapes: the shipped upscale16 architecture (one more UpConv stage), LR (1,4,8,8,10) ->
    SR (1,3,128,128,10), bf16, against the CPU oracle with the same weights."""
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    from gan_sr_wind_field_b200.tools import initialization
    from oracle import wind_oracle as wo
    torch.manual_seed(2001)
    G = Generator_3D(4, 3, 128, 16, upscale=16, hr_kern_size=5, lff_kern_size=1, dropout_probability=0.1)
    initialization.init_weights(G, 0.1)
    G.eval()
    LR, HR, Z, x, y = wo.synthetic_batch(1, hr_xy=128, nz=10, scale=16, seed=2001)
    assert LR.shape == (1, 4, 8, 8, 10)
    with torch.no_grad():
        ref = wo.generator_forward(G.state_dict(), LR, Z)
    G.cuda()
    with ops.precision("bf16"), torch.no_grad():
        out = G(LR.cuda(), Z.cuda())
    assert out.shape == (1, 3, 128, 128, 10)
    assert rel_l2(out, ref) <= TOL["bf16"]


@pytest.mark.gpu
def test_trunk_size_rrdb_forward_backward_vs_oracle():
    """One RRDB at the shipped trunk size (128 features, gc 32, 4 dense convs + LFF per RDB, 16x16x10 volume, B=2)
    in bf16: forward (x-folded dense convs), dL/dx and every parameter gradient (merged dense-conv weight gradient)
    against the CPU oracle in fp32, each bounded by the intrinsic bf16-operand envelope measured on the CPU."""
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.CNN_models.torch_blocks import RRDB
    from oracle import wind_oracle as wo
    from tests.util import bf16_operand_emulation
    torch.manual_seed(11)
    blk = RRDB(128, 32, 5, 1, lrelu_negative_slope=0.2, RDB_residual_scaling=0.2, RRDB_residual_scaling=0.2, mode="3D")
    x = torch.randn(2, 128, 16, 16, 10)
    gy = torch.randn(2, 128, 16, 16, 10)

    def oracle_run():
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
        xi = x.clone().requires_grad_(True)
        y = wo.rrdb(sd, "", xi, n_rdb=3, n_dense=4)
        y.backward(gy)
        return y.detach(), xi.grad, {k: v.grad for k, v in sd.items()}

    y_ref, dx_ref, g_ref = oracle_run()
    with bf16_operand_emulation():
        y_emu, dx_emu, g_emu = oracle_run()
    blk.cuda()
    xg = x.cuda().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    with ops.precision("bf16"):
        y = blk(xg)
        y.backward(gy.cuda())
    bound = lambda emu, ref: 1.5 * rel_l2(emu, ref) + 2e-2
    assert rel_l2(y, y_ref) <= bound(y_emu, y_ref)
    assert rel_l2(xg.grad, dx_ref) <= bound(dx_emu, dx_ref)
    worst = 0.0
    for k, p in blk.named_parameters():
        err, lim = rel_l2(p.grad, g_ref[k]), bound(g_emu[k], g_ref[k])
        worst = max(worst, err / lim)
        assert err <= lim, (k, err, lim)
    print(f"trunk RRDB bf16: fwd {rel_l2(y, y_ref):.2e}, dx {rel_l2(xg.grad, dx_ref):.2e}, worst grad/bound {worst:.2f}")


@pytest.mark.gpu
@pytest.mark.parametrize("groups", ["1", "2"])
def test_batched_trunk_equals_per_block_path(groups, monkeypatch):
    """ops.TrunkFn (one autograd node for a run of RRDBs, weight gradients of all blocks as batched launches,
    ws_trunk_wgrad) against the per-block path (ops.RDBFn: one merged + one LFF GEMM per block) on a 3-RRDB trunk at
    the shipped trunk size.  The reference chain runs the RRDBs one at a time through the per-block path and hands
    dL/dx from one to the next by hand: the forward and dL/dx must agree BIT FOR BIT (same kernels, same joins), the
    parameter gradients to fp32-summation-order noise (both multiply the same bf16 operands; the batched GEMM has no
    split-K).  The per-block path inside ONE autograd graph is compared too, more loosely: its joins of the RRDB skip
    gradients are done by autograd."""
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.CNN_models.torch_blocks import RRDB, run_trunk
    torch.manual_seed(5)
    mods = [RRDB(128, 32, 5, 1, lrelu_negative_slope=0.2, RDB_residual_scaling=0.2, RRDB_residual_scaling=0.2,
                 mode="3D").cuda() for _ in range(3)]
    x = torch.randn(2, 128, 16, 16, 10, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    gy = torch.randn(2, 128, 16, 16, 10, device="cuda")
    params = [p for m in mods for p in m.parameters()]
    monkeypatch.setenv("WINDSR_TRUNK_GROUPS", groups)

    def run(batched):
        monkeypatch.setenv("WINDSR_TRUNK_BATCH", "1" if batched else "0")
        for p in params:
            p.grad = None
        xi = x.clone().requires_grad_(True)
        with ops.precision("bf16"):
            y = run_trunk(mods, xi)
            assert (type(y.grad_fn).__name__ == "TrunkFnBackward") == batched
            y.backward(gy)
        ops.aux_join()
        torch.cuda.synchronize()
        return y.detach().clone(), xi.grad.clone(), [p.grad.clone() for p in params]

    def chain():
        monkeypatch.setenv("WINDSR_TRUNK_BATCH", "0")
        for p in params:
            p.grad = None
        with ops.precision("bf16"):
            hs = [x.clone()]
            with torch.no_grad():
                for m in mods:
                    hs.append(run_trunk([m], hs[-1]))
            d = gy
            for i in range(len(mods) - 1, -1, -1):
                xi = hs[i].clone().requires_grad_(True)
                run_trunk([mods[i]], xi).backward(d)
                ops.aux_join()
                d = xi.grad
        torch.cuda.synchronize()
        return hs[-1], d, [p.grad.clone() for p in params]

    y0, dx0, g0 = chain()
    y1, dx1, g1 = run(True)
    assert torch.equal(y0, y1)
    assert torch.equal(dx1, dx0)
    worst = max(rel_l2(a, b) for a, b in zip(g1, g0))
    assert worst <= 2e-5, worst
    for a, p in zip(g1, params):
        assert a.shape == p.shape and a.is_contiguous()
    y2, dx2, g2 = run(False)
    assert torch.equal(y2, y1)
    assert rel_l2(dx2, dx1) <= 1e-4
    assert max(rel_l2(a, b) for a, b in zip(g2, g1)) <= 5e-3
