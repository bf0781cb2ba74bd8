"""Pins the CPU oracle (oracle/wind_oracle.py, oracle/np_conv3d.py).

The reference ships no tests / golden vectors, so the pin is the reference itself:
  * tests/golden/*.npz — outputs of the UNMODIFIED reference modules (tests/golden/make_golden.py) — checked
    everywhere;
  * the live reference under /root/reference, when it is mounted (this container; absent on the GPU box).
fp32 CPU arithmetic is the same torch primitive on both sides, so the bar here is (near) bit-exactness.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import np_conv3d, refshim, wind_oracle as wo
from tests.util import load_npz, rel_l2, sd_from

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
needs_reference = pytest.mark.skipif(not refshim.available(), reason="reference not mounted")


# ---- golden fixtures -------------------------------------------------------------------------------------
def test_generator_matches_golden():
    z = load_npz("generator_small.npz")
    sd = sd_from(z, "sd/")
    LR = torch.from_numpy(z["LR"]).requires_grad_(True)
    Z = torch.from_numpy(z["Z"])
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = wo.generator_forward(params, LR, Z)
    assert rel_l2(out, z["out"]) < 1e-6
    names = [k[len("grad/"):] for k in z.files if k.startswith("grad/")]
    loss = (out * torch.from_numpy(z["r"])).sum()
    grads = torch.autograd.grad(loss, [LR] + [params[n] for n in names])
    assert rel_l2(grads[0], z["grad_LR"]) < 1e-5
    for n, g in zip(names, grads[1:]):
        assert rel_l2(g, z[f"grad/{n}"]) < 1e-5, n


@pytest.mark.parametrize("tag", ["slicing", "full"])
def test_discriminator_matches_golden(tag):
    z = load_npz(f"discriminator_{tag}.npz")
    sd = sd_from(z, "sd/")
    x = torch.from_numpy(z["x"].astype(np.float32)).requires_grad_(True)
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
              for k, v in sd.items()}
    upd = {}
    out = wo.discriminator_forward(params, x, True, updated_stats=upd)
    assert rel_l2(out, z["out_train"]) < 1e-5
    names = [k[len("grad/"):] for k in z.files if k.startswith("grad/")]
    grads = torch.autograd.grad(out.sum(), [x] + [params[n] for n in names])
    assert rel_l2(grads[0][:, :, ::4, ::4, :], z["grad_x_sub"]) < 1e-4
    for n, g in zip(names, grads[1:]):
        assert rel_l2(g, z[f"grad/{n}"]) < 1e-4, n
    for k, v in upd.items():
        assert rel_l2(v, z[f"after/{k}"]) < 1e-6, k
    sd_eval = dict(sd)
    sd_eval.update(upd)
    out_eval = wo.discriminator_forward(sd_eval, x.detach(), False)
    assert rel_l2(out_eval, z["out_eval_after"]) < 1e-5


def test_windloss_matches_golden():
    z = load_npz("windloss.npz")
    HR, Z, x, y = (torch.from_numpy(z[k]) for k in ("HR", "Z", "x", "y"))
    w = dict(zip(("pixel", "xy", "z", "div", "dxy"), z["weights"].tolist()))
    for sr_key, suffix in (("SR", ""), ("SR2", "2")):
        SR = torch.from_numpy(z[sr_key]).requires_grad_(True)
        if not suffix:
            assert rel_l2(wo.wind_gradient(HR, x, y, Z), z["jac_HR"]) < 1e-6
            assert rel_l2(wo.wind_gradient(SR, x, y, Z), z["jac_SR"]) < 1e-6
        gh, gs = wo.wind_gradient(HR, x, y, Z), wo.wind_gradient(SR, x, y, Z)
        assert np.allclose([float(n) for n in wo.norm_factors(gh, gs)], z["norms" + suffix], rtol=1e-6)
        total, parts = wo.generator_loss(HR, SR, Z, x, y, w)
        got = [float(parts[k]) / w[kk] for k, kk in (("pix", "pixel"), ("xy_gradient", "xy"), ("z_gradient", "z"),
                                                      ("divergence", "div"), ("xy_divergence", "dxy"))]
        assert np.allclose(got, z["terms" + suffix], rtol=2e-5)
        assert abs(float(total) - float(z["total" + suffix])) <= 2e-5 * abs(float(z["total" + suffix]))
        (d,) = torch.autograd.grad(total, SR)
        assert rel_l2(d, z["dSR" + suffix]) < 1e-5
    # the SR2 case really exercises the SR_max/100 branch of every normaliser that can take it
    assert z["norms2"][0] > z["norms"][0]


def test_np_conv3d_pins_torch_conv_semantics():
    g = torch.Generator().manual_seed(0)
    for (cin, cout, vol, k, s, p) in [(3, 4, (5, 6, 7), (3, 3, 3), 1, 1), (2, 3, (8, 8, 6), (4, 4, 3), (2, 2, 1), 1),
                                      (2, 2, (8, 6, 10), (4, 4, 3), (2, 2, 2), 1), (4, 2, (4, 4, 10), (3, 3, 3), (1, 1, 2), 1),
                                      (3, 3, (6, 6, 5), (5, 5, 5), 1, 2), (5, 4, (3, 3, 3), (1, 1, 1), 1, 0)]:
        x = torch.randn(2, cin, *vol, generator=g, dtype=torch.float64).requires_grad_(True)
        w = torch.randn(cout, cin, *k, generator=g, dtype=torch.float64).requires_grad_(True)
        y = F.conv3d(x, w, stride=s, padding=p)
        dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        gx, gw = torch.autograd.grad(y, (x, w), dy)
        assert np.allclose(np_conv3d.conv3d_fwd(x.detach(), w.detach(), s, p), y.detach().numpy(), atol=1e-10)
        assert np.allclose(np_conv3d.conv3d_dgrad(dy, w.detach(), x.shape, s, p), gx.numpy(), atol=1e-10)
        assert np.allclose(np_conv3d.conv3d_wgrad(x.detach(), dy, w.shape, s, p), gw.numpy(), atol=1e-10)
    xx = torch.randn(2, 3, 4, 5, 6)
    assert np.array_equal(np_conv3d.upsample_nearest_xy(xx.numpy()),
                          torch.nn.Upsample(scale_factor=(2, 2, 1), mode="nearest")(xx).numpy())
    assert torch.equal(wo.upsample_nearest_xy(xx), torch.nn.Upsample(scale_factor=(2, 2, 1), mode="nearest")(xx))


def test_gan_step_losses_match_golden():
    """The oracle's loss restatement reproduces the reference's logged G-step / D-step losses."""
    z = load_npz("gan_step.npz")
    G0, D0 = sd_from(z, "G0/"), sd_from(z, "D0/")
    LR, HR, Z, x, y = (torch.from_numpy(z[k]) for k in ("LR", "HR", "Z", "x", "y"))
    it, niter = 1, 100
    real = torch.full((2,), 0.9 + 0.1 * it / niter)
    fake = torch.zeros(2)
    SR = wo.generator_forward(G0, LR, Z)
    y_pred = wo.discriminator_forward(D0, HR, False).squeeze()
    y_fake = wo.discriminator_forward(D0, SR, False).squeeze()
    adv = wo.adversarial_G(y_pred, y_fake, real, fake)
    w = dict(pixel=0.136, xy=3.064, z=0.2, div=0.366, dxy=0.721, adv=0.05)
    total, parts = wo.generator_loss(HR, SR, Z, x, y, w, adv=adv)
    assert abs(float(total) - float(z["G_step/loss/total"])) <= 1e-5 * abs(float(z["G_step/loss/total"]))
    for k, v in parts.items():
        assert abs(float(v) - float(z[f"G_step/loss/{k}"])) <= 1e-5 * max(1e-6, abs(float(z[f"G_step/loss/{k}"]))), k


# ---- live reference --------------------------------------------------------------------------------------
@needs_reference
def test_oracle_equals_live_reference_generator_and_discriminator():
    refshim.activate()
    from CNN_models.Discriminator_3D import Discriminator_3D
    from CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    import tools.initialization as init
    torch.manual_seed(3)
    G = Generator_3D(4, 3, 16, 1, upscale=2, hr_kern_size=5, RDB_gc=8, terrain_number_of_features=8)
    init.init_weights(G, 0.3)
    G.eval()
    LR, HR, Z, x, y = wo.synthetic_batch(1, hr_xy=16, nz=10, scale=2, seed=4)
    with torch.no_grad():
        assert torch.equal(wo.generator_forward(G.state_dict(), LR, Z), G(LR, Z))
    D = Discriminator_3D(3, 4, enable_slicing=True)
    init.init_weights(D, 0.2)
    xin = torch.randn(2, 3, 64, 64, 10)
    sd = {k: v.clone() for k, v in D.state_dict().items()}
    D.train()
    with torch.no_grad():
        assert rel_l2(wo.discriminator_forward(sd, xin, True), D(xin)) < 1e-6


@needs_reference
def test_oracle_equals_live_reference_stencils():
    refshim.activate()
    import process_data
    from GAN_models.wind_field_GAN_3D import get_norm_factors_of_gradients
    LR, HR, Z, x, y = wo.synthetic_batch(2, hr_xy=16, nz=10, scale=4, seed=8)
    SR = HR + 0.1 * torch.randn_like(HR)
    a = process_data.calculate_gradient_of_wind_field(HR, x, y, Z)
    b = wo.wind_gradient(HR, x, y, Z)
    assert rel_l2(b, a) < 1e-7
    na = get_norm_factors_of_gradients(a, process_data.calculate_gradient_of_wind_field(SR, x, y, Z))
    nb = wo.norm_factors(b, wo.wind_gradient(SR, x, y, Z))
    assert all(abs(float(p) - float(q)) <= 1e-6 * abs(float(p)) for p, q in zip(na, nb))
    # FP64: the closed-form stencil agrees with torch.gradient to rounding (8.9e-16)
    a64 = process_data.calculate_gradient_of_wind_field(HR.double(), x.double(), y.double(), Z.double())
    assert rel_l2(wo.wind_gradient(HR.double(), x.double(), y.double(), Z.double()), a64) < 1e-13


def _prep_case(z, i):
    """(fields cropped, flags) of case i of prepare_batch.npz"""
    layout, inc_p, inc_z, inc_ag, xs, ys, k, fx, fy = (int(v) for v in z["cases"][i])
    return dict(inc_p=bool(inc_p), inc_z=bool(inc_z), inc_ag=bool(inc_ag), xs=xs, ys=ys, k=k, fx=fx, fy=fy)


def test_data_oracle_matches_golden_bit_exact():
    """oracle/data_oracle.py (crop + reformat + augment) against outputs of the reference's reformat_to_torch and
    augmentation statements (tests/golden/make_golden_r02.py): bit-exact, all 48 (layout, rot, flip) cases."""
    from oracle import data_oracle as do
    z = load_npz("prepare_batch.npz")
    c = {k[6:]: float(z[k]) for k in z.files if k.startswith("const/")}
    size, cf = int(z["slice_size"]), int(z["coarseness"])
    for i in range(len(z["cases"])):
        m = _prep_case(z, i)
        u, v, w, p, zz, zag = do.crop([z[k] for k in ("u", "v", "w", "p", "z", "zag")], m["xs"], m["ys"], size)
        LR, HR, Z = do.reformat(u, v, w, p, zz, zag, c["Z_MIN"], c["Z_MAX"], c["Z_ABOVE_GROUND_MAX"], c["UVW_MAX"],
                                c["P_MIN"], c["P_MAX"], cf, m["inc_p"], m["inc_z"], m["inc_ag"])
        LR, HR, Z = do.augment(LR, HR, Z, m["k"], m["fx"], m["fy"])
        for name, t in (("LR", LR), ("HR", HR), ("Z", Z)):
            assert torch.equal(t, torch.from_numpy(z[f"case{i}/{name}"])), (i, name, m)


def test_metrics_oracle_matches_golden():
    from oracle import data_oracle as do
    z = load_npz("metrics.npz")
    HR, SR, LR = (torch.from_numpy(z[k]) for k in ("HR", "SR", "LR"))
    psnr, tri_psnr, tri_l1 = do.validation_metrics(LR, HR, SR, int(z["scale"]), "l1")
    assert abs(float(psnr) - float(z["psnr"])) <= 1e-5 and abs(float(tri_psnr) - float(z["tri_psnr"])) <= 1e-5
    assert abs(float(tri_l1) - float(z["tri_l1"])) <= 1e-6
    assert abs(float(do.validation_metrics(LR, HR, SR, int(z["scale"]), "l2")[2]) - float(z["tri_l2"])) <= 1e-6


@needs_reference
def test_data_oracle_equals_live_reference_reformat():
    refshim.activate()
    import process_data
    from oracle import data_oracle as do
    rng = np.random.default_rng(2)
    f = [rng.normal(size=(8, 12, 4)) * 5 for _ in range(6)]
    for flags in ((False, False, False), (False, True, False), (True, True, True), (True, False, False)):
        a = process_data.reformat_to_torch(*f, -2.71, 550.44, 68.46, 32.33, 9e4, 1.05e5, coarseness_factor=4,
                                           include_pressure=flags[0], include_z_channel=flags[1],
                                           include_above_ground_channel=flags[2])
        b = do.reformat(*f, -2.71, 550.44, 68.46, 32.33, 9e4, 1.05e5, 4, *flags)
        assert all(torch.equal(x, y) for x, y in zip(a, b))
