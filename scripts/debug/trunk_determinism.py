"""Run-to-run determinism of the trunk paths (per-block RDBFn / batched TrunkFn): dL/dx and parameter gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.CNN_models.torch_blocks import RRDB, run_trunk

torch.manual_seed(5)
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mods = [RRDB(128, 32, 5, 1, lrelu_negative_slope=0.2, RDB_residual_scaling=0.2, RRDB_residual_scaling=0.2,
             mode="3D").cuda() for _ in range(NB)]
x = torch.randn(2, 128, 16, 16, 10, device="cuda").contiguous(memory_format=torch.channels_last_3d)
gy = torch.randn(2, 128, 16, 16, 10, device="cuda")
params = [p for m in mods for p in m.parameters()]
names = [f"{i}.{n}" for i, m in enumerate(mods) for n, _ in m.named_parameters()]


def run(batched):
    os.environ["WINDSR_TRUNK_BATCH"] = "1" if batched else "0"
    for p in params:
        p.grad = None
    xi = x.clone().requires_grad_(True)
    with ops.precision("bf16"):
        y = run_trunk(mods, xi)
        y.backward(gy)
    ops.aux_join()
    torch.cuda.synchronize()
    return y.detach().clone(), xi.grad.clone(), [p.grad.clone() for p in params]


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


runs = {k: [run(k == "batched") for _ in range(3)] for k in ("per_block", "batched")}
for k, rs in runs.items():
    for i in (1, 2):
        d = rel(rs[i][1], rs[0][1])
        worst = max((rel(a, b), n) for a, b, n in zip(rs[i][2], rs[0][2], names))
        print(f"{k}: run {i} vs run 0: y equal {torch.equal(rs[i][0], rs[0][0])}, dx rel {d:.2e}, worst grad {worst[0]:.2e} ({worst[1]})")
a, b = runs["batched"][0], runs["per_block"][0]
print(f"batched vs per_block: dx rel {rel(a[1], b[1]):.2e}")
diffs = sorted(((rel(p, q), n) for p, q, n in zip(a[2], b[2], names)), reverse=True)
print("largest grad differences:", [(f"{d:.1e}", n) for d, n in diffs[:8]])
dd = (a[1] - b[1]).abs()
print("dx max abs diff", float(dd.max()), "at", [int(v) for v in torch.unravel_index(dd.argmax(), dd.shape)], "nonzero frac", float((dd > 0).float().mean()))

# ground truth for the join between RRDBs: chain single-RRDB backward passes by hand (single-RRDB results are identical
# in both paths)
if NB >= 2:
    os.environ["WINDSR_TRUNK_BATCH"] = "0"
    with ops.precision("bf16"):
        hs = [x.clone()]
        with torch.no_grad():
            for m in mods:
                hs.append(run_trunk([m], hs[-1]))
        d = gy
        for i in range(NB - 1, -1, -1):
            xi = hs[i].clone().requires_grad_(True)
            yi = run_trunk([mods[i]], xi)
            yi.backward(d)
            ops.aux_join()
            d = xi.grad
    torch.cuda.synchronize()
    print(f"manual chain vs per_block: dx rel {rel(d, runs['per_block'][0][1]):.2e}; vs batched: {rel(d, runs['batched'][0][1]):.2e}")
