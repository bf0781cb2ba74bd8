"""BASELINE config #5: Conv3d microbench over every generator / discriminator layer shape of the upscale8 config
(SURVEY 8-a rows G1-G8, D1-D10), forward / data-gradient / weight-gradient, B = 8, bf16 — this package's layer
(module + autograd, i.e. what the training step runs) next to torch.nn.Conv3d (cuDNN, bf16, channels_last_3d) on the
same GPU, same inputs and weights.  Writes a markdown table.

fwd = module(x); dgrad = backward with the weight frozen; wgrad = (full backward) - dgrad.
Usage: python scripts/layer_sweep.py [out.md] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.CNN_models.torch_blocks import Conv3d

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "layer_sweep.md")
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B = 8
# row, cin, cout, kernel, stride, pad, input volume
LAYERS = [
    ("G1 feature_conv", 4, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), (16, 16, 10)),
    ("G2 RDB conv0", 128, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), (16, 16, 10)),
    ("G2 RDB conv1", 160, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), (16, 16, 10)),
    ("G2 RDB conv2", 192, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), (16, 16, 10)),
    ("G2 RDB conv3", 224, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), (16, 16, 10)),
    ("G3 LFF", 256, 128, (1, 1, 1), (1, 1, 1), (0, 0, 0), (16, 16, 10)),
    ("G4 lr_conv", 128, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), (16, 16, 10)),
    ("G5 UpConv 32^2", 128, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), (32, 32, 10)),
    ("G5 UpConv 64^2", 128, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), (64, 64, 10)),
    ("G5 UpConv 128^2", 128, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), (128, 128, 10)),
    ("G6 terrain 1->16", 1, 16, (3, 3, 3), (1, 1, 1), (1, 1, 1), (128, 128, 10)),
    ("G6 terrain 16->16", 16, 16, (3, 3, 3), (1, 1, 1), (1, 1, 1), (128, 128, 10)),
    ("G7 hr_convs.0", 144, 144, (5, 5, 5), (1, 1, 1), (2, 2, 2), (128, 128, 10)),
    ("G8 hr_convs.2", 144, 3, (5, 5, 5), (1, 1, 1), (2, 2, 2), (128, 128, 10)),
    ("D1", 3, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), (128, 128, 10)),
    ("D2", 32, 32, (4, 4, 3), (2, 2, 1), (1, 1, 1), (128, 128, 10)),
    ("D3", 32, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1), (64, 64, 10)),
    ("D4", 64, 64, (4, 4, 3), (2, 2, 1), (1, 1, 1), (64, 64, 10)),
    ("D5", 64, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), (32, 32, 10)),
    ("D6", 128, 128, (4, 4, 3), (2, 2, 1), (1, 1, 1), (32, 32, 10)),
    ("D7", 128, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1), (16, 16, 10)),
    ("D8", 256, 256, (4, 4, 3), (2, 2, 1), (1, 1, 1), (16, 16, 10)),
    ("D9", 256, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1), (8, 8, 10)),
    ("D10", 256, 256, (4, 4, 3), (2, 2, 2), (1, 1, 1), (8, 8, 10)),
]


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def three(make_out, x, params):
    """(fwd, dgrad, wgrad) ms of y = make_out(); x and params are leaf tensors."""
    with torch.no_grad():
        t_f = timed(make_out, REPS)
    x.requires_grad_(True)
    for p in params:
        p.requires_grad_(False)
    y = make_out()
    g = torch.randn_like(y)
    t_d = timed(lambda: torch.autograd.grad(y, x, g, retain_graph=True), REPS) if x.shape[1] > 4 else float("nan")
    x.requires_grad_(False)
    for p in params:
        p.requires_grad_(True)
    y = make_out()
    t_w = timed(lambda: torch.autograd.grad(y, params, g, retain_graph=True), REPS)
    return t_f, t_d, t_w


ops.set_precision("bf16")
torch.backends.cudnn.benchmark = True
rows = []
gen = torch.Generator(device="cuda").manual_seed(0)
for name, cin, cout, k, s, p, vol in LAYERS:
    # this package: bf16 channels-last activations for cin >= 16, the fp32 boundary tensor otherwise
    if cin >= 16:
        x = ops.empty_cl(B, cin, *vol, torch.bfloat16, "cuda")
        x.copy_(torch.randn(B, cin, *vol, generator=gen, device="cuda"))
    else:
        x = torch.randn(B, cin, *vol, generator=gen, device="cuda")
    m = Conv3d(cin, cout, k, s, p, bias=(name.startswith("G3") or name.startswith("G8"))).cuda()
    if name.startswith("G8"):
        mine = three(lambda: ops.XFoldConvFn.apply(x, m.weight, m.bias, m.padding), x, [m.weight])
    else:
        mine = three(lambda: m(x), x, [m.weight])
    # torch / cuDNN, bf16 channels_last_3d
    ref = torch.nn.Conv3d(cin, cout, k, s, p, bias=m.bias is not None).cuda().to(torch.bfloat16)
    ref = ref.to(memory_format=torch.channels_last_3d)
    xr = x.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    theirs = three(lambda: ref(xr), xr, [ref.weight])
    xo = [(vol[i] + 2 * p[i] - k[i]) // s[i] + 1 for i in range(3)]
    gflop = 2.0 * B * xo[0] * xo[1] * xo[2] * cin * cout * k[0] * k[1] * k[2] / 1e9
    rows.append((name, f"{cin}->{cout} k{k} s{s} @{vol}", gflop, mine, theirs))
    print(name, [f"{v:.3f}" for v in mine], [f"{v:.3f}" for v in theirs], flush=True)
    del m, ref, x, xr

lines = ["# Conv3d layer sweep (BASELINE config #5): upscale8 shapes, B=8, bf16, one B200",
         "",
         "`python scripts/layer_sweep.py` — ms per pass, CUDA events over %d launches after 3 warm-ups; "
         "TFLOP/s by the dense-MAC convention." % REPS,
         "ours = this package's layer as the training step runs it (module + autograd, tcgen05 kernels); "
         "torch = torch.nn.Conv3d bf16 channels_last_3d (cuDNN, `cudnn.benchmark=True`) on the same GPU.",
         "dgrad of the first layers (cin <= 4) is never needed by the step and is not timed.", "",
         "| layer | shape | GFLOP/pass | ours fwd | dgrad | wgrad | ours fwd TFLOP/s | torch fwd | dgrad | wgrad | speed-up fwd / dgrad / wgrad |",
         "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
tot_m, tot_t = 0.0, 0.0
for name, shape, gflop, mine, theirs in rows:
    su = " / ".join("-" if a != a or b != b else f"{b / a:.2f}x" for a, b in zip(mine, theirs))
    f = lambda v: "-" if v != v else f"{v:.3f}"
    lines.append(f"| {name} | {shape} | {gflop:.1f} | {f(mine[0])} | {f(mine[1])} | {f(mine[2])} | "
                 f"{gflop / mine[0]:.0f} | {f(theirs[0])} | {f(theirs[1])} | {f(theirs[2])} | {su} |")
open(out_path, "w").write("\n".join(lines) + "\n")
print("wrote", out_path)
