"""Bring-up diagnostic (GPU box): per-layer rel-L2 of fwd / dgrad / wgrad for both kernel families, printed as
a table and written to gpurun_out/diag_<tag>.json.  Not a test: it never asserts, so one wrong kernel does not
hide the others.  Usage: python scripts/gpu_diag.py <fp32|bf16> [name-substring]"""
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from gan_sr_wind_field_b200 import _lib, ops
from tests.test_gpu_kernels import LAYERS, _case, _to_act
from tests.util import rel_l2

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
filt = sys.argv[2] if len(sys.argv) > 2 else ""
rows = []
print(torch.cuda.get_device_name(0), "tcgen05:", _lib.load().ws_device_supports_tcgen05(), flush=True)
for (name, n, cin, cout, vol, k, s, p) in LAYERS:
    if filt and filt not in name:
        continue
    x, w = _case(n, cin, cout, vol, k, s, p)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y_ref = F.conv3d(xr, wr, stride=s, padding=p)
    dy = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(1))
    gx_ref, gw_ref = torch.autograd.grad(y_ref, (xr, wr), dy)
    row = dict(name=name, mode=mode)
    with ops.precision(mode):
        dt = ops.act_dtype()
        xa, dya, wc = _to_act(x, dt), _to_act(dy, dt), w.cuda()
        shape = ops.make_shape(x.shape, cout, w.shape[2:], s, p)
        row["paths"] = (ops.fwd_path(shape, xa), ops.dgrad_path(shape, dya), ops.wgrad_path(shape, xa, dya))
        for what in ("fwd", "dgrad", "wgrad"):
            try:
                if what == "fwd":
                    y = ops.empty_cl(*y_ref.shape, dt, "cuda")
                    ops.conv_fwd(xa, wc, None, shape, y)
                    torch.cuda.synchronize()
                    row[what] = rel_l2(y.float(), y_ref)
                elif what == "dgrad":
                    dx = ops.empty_cl(*x.shape, torch.float32, "cuda")
                    ops.conv_dgrad(dya, wc, None, shape, dx)
                    torch.cuda.synchronize()
                    row[what] = rel_l2(dx, gx_ref)
                else:
                    dw, _ = ops.conv_wgrad(xa, dya, shape)
                    torch.cuda.synchronize()
                    row[what] = rel_l2(dw, gw_ref)
            except Exception as e:  # noqa: BLE001
                row[what] = f"ERR {type(e).__name__}: {str(e)[:200]}"
                traceback.print_exc()
    rows.append(row)
    print(f"{name:14s} {mode} paths={row['paths']} fwd={row['fwd']} dgrad={row['dgrad']} wgrad={row['wgrad']}",
          flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"diag_{mode}{'_' + filt if filt else ''}.json"), "w") as f:
    json.dump(rows, f, indent=1)
