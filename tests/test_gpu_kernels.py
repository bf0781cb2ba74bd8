"""GPU parity of the individual kernels behind the C-ABI against the oracle primitives (torch CPU fp32 conv3d,
whose semantics oracle/np_conv3d.py pins) on identical seeded inputs.  Bar: rel-L2 <= 1e-5 in FP32 mode,
<= 2e-2 in BF16 mode (north star); index/gather work bit-exact."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import TOL, rel_l2

pytestmark = pytest.mark.gpu

# (name, n, cin, cout, (x,y,z), kernel, stride, pad): every distinct layer geometry of SURVEY §8-a, spatially
# scaled down so the CPU oracle finishes in seconds
LAYERS = [
    ("G1_feature", 2, 4, 128, (8, 8, 10), 3, 1, 1),
    ("G2_rdb0", 2, 128, 32, (8, 8, 10), 3, 1, 1),
    ("G2_rdb3", 2, 224, 32, (8, 8, 10), 3, 1, 1),
    ("G3_lff", 2, 256, 128, (8, 8, 10), 1, 1, 0),
    ("G4_lr", 2, 128, 128, (8, 8, 10), 3, 1, 1),
    ("G5_upconv", 1, 128, 128, (16, 16, 10), 3, 1, 1),
    ("G6a_terrain", 1, 1, 16, (16, 16, 10), 3, 1, 1),
    ("G6b_terrain", 1, 16, 16, (16, 16, 10), 3, 1, 1),
    ("G7_hr0", 1, 144, 144, (13, 9, 10), 5, 1, 2),
    ("G8_hr2", 1, 144, 3, (12, 12, 10), 5, 1, 2),
    ("D1", 1, 3, 32, (16, 16, 10), 3, 1, 1),
    ("D2_strided", 1, 32, 32, (16, 16, 10), (4, 4, 3), (2, 2, 1), 1),
    ("D6_strided", 1, 128, 128, (8, 8, 10), (4, 4, 3), (2, 2, 1), 1),
    ("D7", 1, 128, 256, (8, 8, 10), 3, 1, 1),
    ("D10_halve_z", 2, 256, 256, (8, 8, 10), (4, 4, 3), (2, 2, 2), 1),
    ("Dslice_z2", 1, 64, 64, (4, 4, 10), 3, (1, 1, 2), 1),
]


def _case(n, cin, cout, vol, k, s, p, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, *vol, generator=g)
    kk = (k, k, k) if isinstance(k, int) else k
    w = torch.randn(cout, cin, *kk, generator=g) / (cin * kk[0] * kk[1] * kk[2]) ** 0.5
    return x, w


def _to_act(t, dtype):
    from gan_sr_wind_field_b200 import ops
    out = ops.empty_cl(*t.shape, dtype, "cuda")
    out.copy_(t.cuda())
    return out


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("layer", LAYERS, ids=[l[0] for l in LAYERS])
def test_conv_fwd_dgrad_wgrad(layer, mode):
    from gan_sr_wind_field_b200 import ops
    name, n, cin, cout, vol, k, s, p = layer
    x, w = _case(n, cin, cout, vol, k, s, p)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    if mode == "bf16":  # the reference sees the same bf16-rounded operands; tolerance covers the rest
        pass
    y_ref = F.conv3d(xr, wr, stride=s, padding=p)
    dy = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(1))
    gx_ref, gw_ref = torch.autograd.grad(y_ref, (xr, wr), dy)
    with ops.precision(mode):
        dt = ops.act_dtype()
        xa = _to_act(x, dt)
        wc = w.cuda()
        shape = ops.make_shape(x.shape, cout, w.shape[2:], s, p)
        xo, yo, zo = ops.out_dims(shape)
        assert (xo, yo, zo) == tuple(y_ref.shape[2:])
        y = ops.empty_cl(n, cout, xo, yo, zo, dt, "cuda")
        ops.conv_fwd(xa, wc, None, shape, y)
        dya = _to_act(dy, dt)
        dx = ops.empty_cl(*x.shape, torch.float32, "cuda")
        ops.conv_dgrad(dya, wc, None, shape, dx)
        dw, _ = ops.conv_wgrad(xa, dya, shape)
        torch.cuda.synchronize()
    tol = TOL[mode]
    errs = dict(fwd=rel_l2(y.float(), y_ref), dgrad=rel_l2(dx, gx_ref), wgrad=rel_l2(dw, gw_ref))
    assert all(e <= tol for e in errs.values()), f"{name} {mode}: {errs}"


def test_tensor_core_path_is_selected():
    """BF16 mode must resolve the heavy layers to the tcgen05 kernels (not the CUDA-core family)."""
    from gan_sr_wind_field_b200 import _lib, ops
    with ops.precision("bf16"):
        x = ops.empty_cl(1, 144, 8, 8, 10, torch.bfloat16, "cuda")
        shape = ops.make_shape(x.shape, 144, (5, 5, 5), 1, 2)
        dy = ops.empty_cl(1, 144, 8, 8, 10, torch.bfloat16, "cuda")
        assert ops.fwd_path(shape, x) == _lib.PATH_TCGEN05
        assert ops.dgrad_path(shape, dy) == _lib.PATH_TCGEN05
        assert ops.wgrad_path(shape, x, dy) == _lib.PATH_TCGEN05
    with ops.precision("fp32"):
        xf = ops.empty_cl(1, 144, 8, 8, 10, torch.float32, "cuda")
        assert ops.fwd_path(shape, xf) == _lib.PATH_SIMT
    with ops.precision("tf32"):  # kind::tf32 on fp32 activations; strided convs stay on the CUDA cores
        dyf = ops.empty_cl(1, 144, 8, 8, 10, torch.float32, "cuda")
        assert ops.fwd_path(shape, xf) == _lib.PATH_TCGEN05
        assert ops.dgrad_path(shape, dyf) == _lib.PATH_TCGEN05
        assert ops.wgrad_path(shape, xf, dyf) == _lib.PATH_TCGEN05
        strided = ops.make_shape(xf.shape, 144, (4, 4, 3), (2, 2, 1), 1)
        assert ops.fwd_path(strided, xf) == _lib.PATH_SIMT


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_fused_epilogue(mode):
    """bias + LeakyReLU + channel (dropout) scale + two residuals + slice write + second output."""
    from gan_sr_wind_field_b200 import ops
    n, cin, cout, vol = 2, 64, 48, (6, 5, 10)
    x, w = _case(n, cin, cout, vol, 3, 1, 1, seed=3)
    g = torch.Generator().manual_seed(4)
    bias = torch.randn(cout, generator=g)
    cs = (torch.rand(n, cout, generator=g) > 0.3).float() * 1.25
    r1 = torch.randn(n, cout, *vol, generator=g)
    r2 = torch.randn(n, cout, *vol, generator=g)
    ref = F.leaky_relu(F.conv3d(x, w, bias, padding=1), 0.2) * cs[:, :, None, None, None]
    ref = 0.04 * ref + 0.2 * r1 + r2
    with ops.precision(mode):
        dt = ops.act_dtype()
        xa = _to_act(x, dt)
        big = ops.zeros_cl(n, cout + 16, *vol, dt, "cuda")
        out2 = torch.empty(n, cout, *vol, dtype=torch.float32, device="cuda")  # NCXYZ contiguous
        shape = ops.make_shape(x.shape, cout, (3, 3, 3), 1, 1)
        ops.conv_fwd(xa, w.cuda(), None, shape, big[:, 16:], bias=bias.cuda(), chan_scale=cs.cuda().reshape(-1),
                     slope=0.2, alpha=0.04, res1=_to_act(r1, torch.float32), beta1=0.2, res2=r2.cuda(), beta2=1.0,
                     out2=out2)
        torch.cuda.synchronize()
    assert rel_l2(big[:, 16:].float(), ref) <= TOL[mode]
    assert rel_l2(out2, ref) <= TOL[mode]
    assert float(big[:, :16].float().abs().max()) == 0.0  # neighbouring channels of the concat buffer untouched


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("channels_last", [True, False])
def test_upsample_bit_exact(dtype, channels_last):
    from gan_sr_wind_field_b200 import ops
    x = torch.randn(2, 24, 5, 7, 10).to(dtype)
    ref = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
    xa = _to_act(x, dtype) if channels_last else x.cuda()
    out = ops.UpsampleFn.apply(xa)
    assert torch.equal(out.cpu(), ref)  # bit-exact index map out[x,y,z] = in[x//2, y//2, z]
    dy = torch.randn(ref.shape).to(dtype)
    din = torch.empty_like(xa)
    ops.upsample_bwd(_to_act(dy, dtype) if channels_last else dy.cuda(), din)
    d = dy.float()
    ref_b = d[:, :, 0::2, 0::2] + d[:, :, 0::2, 1::2] + d[:, :, 1::2, 0::2] + d[:, :, 1::2, 1::2]
    assert rel_l2(din.float(), ref_b) <= (1e-6 if dtype == torch.float32 else 8e-3)


def test_copy_slice_and_layout_bit_exact():
    """dense-concat slice writes / NCXYZ <-> channels-last moves are pure index work: bit-exact."""
    from gan_sr_wind_field_b200 import ops
    x = torch.randn(2, 20, 4, 6, 10)
    buf = ops.zeros_cl(2, 52, 4, 6, 10, torch.float32, "cuda")
    ops.copy_(x.cuda(), buf[:, 32:])
    assert torch.equal(buf[:, 32:].cpu(), x) and float(buf[:, :32].abs().max()) == 0.0
    back = torch.empty(2, 20, 4, 6, 10, device="cuda")
    ops.copy_(buf[:, 32:], back)
    assert torch.equal(back.cpu(), x)
    xb = x.bfloat16()
    b16 = ops.empty_cl(2, 20, 4, 6, 10, torch.bfloat16, "cuda")
    ops.copy_(x.cuda(), b16)
    assert torch.equal(b16.cpu(), xb)  # round-to-nearest-even like torch


def test_empty_and_ragged_edges():
    """1-voxel-wide volumes, channel counts that are not multiples of the vector width, batch of 1."""
    from gan_sr_wind_field_b200 import ops
    for mode in ("fp32", "bf16"):
        for (cin, cout, vol, k, p) in [(3, 5, (1, 1, 10), 3, 1), (17, 19, (3, 2, 10), 3, 1), (40, 24, (2, 9, 5), 5, 2)]:
            x, w = _case(1, cin, cout, vol, k, 1, p, seed=9)
            ref = F.conv3d(x, w, padding=p)
            with ops.precision(mode):
                dt = ops.act_dtype()
                shape = ops.make_shape(x.shape, cout, w.shape[2:], 1, p)
                y = ops.empty_cl(1, cout, *ref.shape[2:], dt, "cuda")
                ops.conv_fwd(_to_act(x, dt), w.cuda(), None, shape, y)
            assert rel_l2(y.float(), ref) <= TOL[mode], (mode, cin, cout, vol)


@pytest.mark.gpu
def test_module_path_144_channels_remainder_xfold_wgrad():
    """hr_convs.0 as the training step runs it (module + autograd): 144 = 128 + 16 output channels, so the weight
    gradient is a 128-row tensor-core GEMM plus an x-folded (1,5,5) GEMM for the 16-channel remainder (ws_xunfold);
    forward, dL/dx and dL/dw against torch's fp32 CPU conv."""
    import torch.nn.functional as F
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.CNN_models.torch_blocks import Conv3d
    torch.manual_seed(9)
    conv = Conv3d(144, 144, 5, 1, 2, bias=False)
    x = torch.randn(1, 144, 13, 9, 10)
    gy = torch.randn(1, 144, 13, 9, 10)
    xr = x.clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, padding=2)
    yr.backward(gy)
    conv.cuda()
    with ops.precision("bf16"):
        xg = ops.empty_cl(1, 144, 13, 9, 10, torch.bfloat16, "cuda")
        xg.copy_(x.cuda())
        xg.requires_grad_(True)
        y = conv(xg)
        y.backward(gy.cuda().to(y.dtype))
    assert rel_l2(y, yr) <= TOL["bf16"]
    assert rel_l2(xg.grad, xr.grad) <= TOL["bf16"]
    assert rel_l2(conv.weight.grad, wr.grad) <= TOL["bf16"]
    # both halves of the output-channel split are right on their own
    assert rel_l2(conv.weight.grad[:128], wr.grad[:128]) <= TOL["bf16"]
    assert rel_l2(conv.weight.grad[128:], wr.grad[128:]) <= TOL["bf16"]


@pytest.mark.parametrize("fn_name", ["XFoldConvFn", "XYFoldConvFn"])
def test_folded_narrow_conv_matches_conv3d(fn_name):
    """hr_convs.2 (5x5x5, 144 -> 3, bias; Generator_3D_Resnet_ESRGAN.py:105-110) with the lateral taps folded into the
    channel dimension: forward, dL/dx, dL/dw, dL/db against torch's CPU conv3d on the same bf16-rounded operands."""
    from gan_sr_wind_field_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, cin, cout, vol = 2, 144, 3, (11, 13, 10)
    x = torch.randn(n, cin, *vol, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, 5, 5, 5, generator=g) / (cin * 125) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g)
    dy = torch.randn(n, cout, *vol, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv3d(xr, wr, br, padding=2)
    gx, gw, gb = torch.autograd.grad(ref, (xr, wr, br), dy)
    with ops.precision("bf16"):
        xa = _to_act(x, torch.bfloat16).requires_grad_(True)
        wc, bc = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        out = getattr(ops, fn_name).apply(xa, wc, bc, (2, 2, 2))
        out.backward(dy.cuda())
        torch.cuda.synchronize()
    assert out.shape == ref.shape and out.is_contiguous()
    errs = dict(fwd=rel_l2(out, ref), dx=rel_l2(xa.grad.float(), gx), dw=rel_l2(wc.grad, gw), db=rel_l2(bc.grad, gb))
    assert all(e <= TOL["bf16"] for e in errs.values()), errs
