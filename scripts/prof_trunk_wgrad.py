"""Batched trunk weight gradient (ws_trunk_wgrad) alone: correctness against the per-conv weight-gradient kernel on
a few blocks, then timing under the WS_TRUNK_WGRAD_FORCE="taps_per_cta,groups_per_cta" settings given on the command
line (default: the cost model's choice)."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gan_sr_wind_field_b200 import ops, _lib
from gan_sr_wind_field_b200._lib import load, view, check

dev = torch.device("cuda:0")
ops.set_precision("bf16")
lib = load()
R, n, X, Y, Z, f, gc, nconv = int(os.environ.get("R", 48)), 8, 16, 16, 10, 128, 32, 4
ctot = f + nconv * gc
torch.manual_seed(0)
slab = (torch.randn((R * n, X, Y, Z, ctot), device=dev) * 0.5).bfloat16().permute(0, 4, 1, 2, 3)
g = (torch.randn((R * n, X, Y, Z, nconv * gc), device=dev) * 0.5).bfloat16().permute(0, 4, 1, 2, 3)
gl = (torch.randn((R * n, X, Y, Z, f), device=dev) * 0.5).bfloat16().permute(0, 4, 1, 2, 3)
desc = _lib.WsRdbDesc(n, X, Y, Z, f, gc, nconv, 3, 1, 0.2, 0.2, 1.0, 0.0, ops.math_mode(), 0)
assert lib.ws_trunk_wgrad_supported(C.byref(desc))
rec = int(lib.ws_trunk_wgrad_record_floats(C.byref(desc)))
flat = torch.zeros((R, rec), device=dev)
wsp = torch.empty(int(lib.ws_trunk_wgrad_workspace_bytes(C.byref(desc), R)), dtype=torch.uint8, device=dev)


def run():
    bv, gv, glv = view(slab[:n]), view(g[:n]), view(gl[:n])
    check(lib.ws_trunk_wgrad(C.byref(desc), R, C.byref(bv), C.byref(gv), C.byref(glv), flat.data_ptr(), rec,
                             wsp.data_ptr(), wsp.numel(), _lib.stream_ptr()), "ws_trunk_wgrad")


run()
torch.cuda.synchronize()
worst = 0.0
for r in sorted({0, R // 2, R - 1}):
    off = 0
    b = slab[r * n:(r + 1) * n]
    for i in range(nconv + 1):
        cin = f + i * gc
        if i < nconv:
            shape = ops.make_shape((n, cin, X, Y, Z), gc, 3, 1, 1)
            dw, _ = ops.conv_wgrad(b[:, :cin], g[r * n:(r + 1) * n, i * gc:(i + 1) * gc], shape)
            db = None
        else:
            shape = ops.make_shape((n, ctot, X, Y, Z), f, 1, 1, 0)
            dw, db = ops.conv_wgrad(b, gl[r * n:(r + 1) * n], shape, want_bias=True)
        got = flat[r, off:off + dw.numel()].view(dw.shape)
        err = float((got - dw).norm() / dw.norm())
        worst = max(worst, err)
        off += dw.numel()
        if db is not None:
            gb = flat[r, off:off + db.numel()]
            err = float((gb - db).norm() / db.norm())
            worst = max(worst, err)
            off += db.numel()
    assert off == rec
print(f"trunk wgrad vs per-conv kernels: worst rel-L2 {worst:.2e}")
assert worst < 1e-4

flops = 2.0 * R * n * X * Y * Z * (27 * gc * sum(f + i * gc for i in range(nconv)) + ctot * f)
for force in (sys.argv[1:] or [""]):
    if force:
        os.environ["WS_TRUNK_WGRAD_FORCE"] = force
    else:
        os.environ.pop("WS_TRUNK_WGRAD_FORCE", None)
    for _ in range(3):
        run()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        run()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print(f"force={force or 'model'}: {ms:.3f} ms per call ({R} blocks), {flops / ms * 1e-9:.0f} TFLOP/s")
