"""Micro-driver for profiling: the hr_convs.0 forward (5x5x5, 144->144 @128x128x10, B=8, bf16) a few times.
Usage: python scripts/prof_conv.py [reps] [layer: g7|g5|rdb]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gan_sr_wind_field_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
layer = sys.argv[2] if len(sys.argv) > 2 else "g7"
cfgs = {"g7": (8, 144, 144, (128, 128, 10), 5, 2), "g5": (8, 128, 128, (128, 128, 10), 3, 1),
        "rdb": (8, 224, 32, (16, 16, 10), 3, 1), "g8": (8, 144, 3, (128, 128, 10), 5, 2),
        "dg": (8, 32, 224, (16, 16, 10), 3, 1), "lff": (8, 256, 128, (16, 16, 10), 1, 0),
        "rdb0": (8, 128, 32, (16, 16, 10), 3, 1),
        # x-folded look-alikes of the dense convs: (1,3,3) kernel, the 3 kx taps side by side on N (fp32 U)
        "rdbx": (8, 224, 96, (16, 16, 10), (1, 3, 3), (0, 1, 1)), "rdb0x": (8, 128, 96, (16, 16, 10), (1, 3, 3), (0, 1, 1)),
        # dgrad of the widest / narrowest dense conv
        "dg0": (8, 32, 128, (16, 16, 10), 3, 1)}
n, cin, cout, vol, k, p = cfgs[layer]
ops.set_precision("bf16")
g = torch.Generator(device="cuda").manual_seed(0)
x = ops.empty_cl(n, cin, *vol, torch.bfloat16, "cuda")
x.copy_(torch.randn(n, cin, *vol, generator=g, device="cuda"))
kk = (k, k, k) if isinstance(k, int) else k
w = torch.randn(cout, cin, *kk, generator=g, device="cuda") / (cin * kk[0] * kk[1] * kk[2]) ** 0.5
shape = ops.make_shape(x.shape, cout, kk, 1, p)
acc = layer in ("dg", "dg0")   # dense-conv dgrad look-alike: fp32 output accumulated in place
y = ops.zeros_cl(n, cout, *vol, torch.float32 if acc or layer.endswith("x") else torch.bfloat16, "cuda")
cache = ops.PackedWeights()
kw = dict(res1=y, beta1=1.0) if acc else dict(slope=0.2)
if len(sys.argv) > 3 and sys.argv[3] == "wgrad":
    # weight-gradient of the same layer (main 128-row M block only for cout = 144, as the step runs it)
    co = min(cout, 128)
    dy = ops.empty_cl(n, co, *vol, torch.bfloat16, "cuda")
    dy.copy_(torch.randn(n, co, *vol, generator=g, device="cuda"))
    shape_w = ops.make_shape(x.shape, co, kk, 1, p)
    ops.conv_wgrad(x, dy, shape_w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.conv_wgrad(x, dy, shape_w)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * n * vol[0] * vol[1] * vol[2] * cin * co * kk[0] * kk[1] * kk[2]
    print(f"{layer} wgrad ({cin}->{co}): {ms:.3f} ms/launch, {fl / ms / 1e9:.1f} TFLOP/s (dense-MAC convention)")
    sys.exit(0)
ops.conv_fwd(x, w, cache, shape, y, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.conv_fwd(x, w, cache, shape, y, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2.0 * n * vol[0] * vol[1] * vol[2] * cin * cout * kk[0] * kk[1] * kk[2]
print(f"{layer}: {ms:.3f} ms/launch, {fl / ms / 1e9:.1f} TFLOP/s (dense-MAC convention)")
