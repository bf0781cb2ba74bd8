"""Per-tensor parity table at the SHIPPED configuration (upscale8 ini: 128 features, 16 RRDBs, init scale 0.1,
dropout mask injected, the reference's generator loss with the ini's weights), for every precision mode, against
``oracle/wind_oracle.py`` on the GPU in strict fp32 (tests/parity_util.py).  Writes profiles/r02_parity_table.md.

    python scripts/parity_table.py [--batch 8] [--out profiles/r02_parity_table.md]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import parity_util as pu  # noqa: E402
from tests.util import TOL, rel_l2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_parity_table.md"))
    args = ap.parse_args()
    from gan_sr_wind_field_b200 import ops
    from oracle import wind_oracle as wo
    B = args.batch
    gan, cfg = pu.shipped_gan()
    batch = tuple(t.cuda() for t in wo.synthetic_batch(B, hr_xy=128, nz=10, scale=8, seed=2001))
    g = torch.Generator().manual_seed(7)
    p = cfg.generator.dropout_probability
    ds = (torch.bernoulli(torch.full((B, 144), 1.0 - p), generator=g) / (1.0 - p)).cuda()
    sd = {k: v.detach().clone() for k, v in gan.G.state_dict().items()}
    w = pu.loss_weights(cfg)
    # arbiter: the oracle in float64; "ref fp32" columns = the oracle in strict fp32 against it, i.e. how far the
    # reference's own fp32 arithmetic is from the exact result (LeakyReLU sign flips of near-zero pre-activations)
    SR_ref, L_ref, g_ref, dLR_ref = pu.oracle_generator_step(sd, batch, w, ds, dtype=torch.float64)
    SR32, L32, g32, dLR32 = pu.oracle_generator_step(sd, batch, w, ds)
    ref32 = pu.grad_errors(g32, g_ref)
    ref32_rows = pu.summarize(ref32)
    lines = [f"# Parity table — upscale8 generator step at the shipped configuration, B = {B} (round 2)", "",
             "`python scripts/parity_table.py` on one B200.  Checker: `oracle/wind_oracle.py` executed on the GPU in strict "
             "**float64** (the arbiter), the restatement `tests/test_oracle_pinned.py` pins bit-exactly to the reference; "
             "*ref fp32* = the same oracle in strict fp32 (TF32 off) against the float64 result — the error the "
             "reference's own fp32 arithmetic has.  ",
             "Step: `G(LR, Z)` in train mode with a fixed Dropout3d mask, the reference's generator loss (pixel L1 0.136 + "
             "xy-gradient 3.064 + divergence 0.366 + xy-divergence 0.721, adversarial weight 0), backward to every "
             "parameter and to LR.  rel-L2 = ||ours - oracle|| / ||oracle|| per tensor; rows aggregate the tensors of one "
             "layer family (count, median, max).  *envelope* = the same oracle with its conv operands (and incoming "
             "gradients) rounded to the mode's operand format, fp32 accumulation: the error ANY implementation with "
             "that operand format has.", ""]
    for mode in [m for m in ("fp32", "tf32", "bf16") if m in ops.PRECISIONS]:
        SR, L, gm, dLR = pu.native_generator_step(gan, batch, mode, ds)
        errs = pu.grad_errors(gm, g_ref)
        env = None
        if mode != "fp32":
            _, _, g_env, dLR_env = pu.oracle_generator_step(sd, batch, w, ds, rounding=mode)
            env = pu.grad_errors(g_env, g_ref)
            env_rows = pu.summarize(env)
        rows = pu.summarize(errs)
        tol = TOL[mode]
        n_over = sum(1 for e in errs.values() if e > tol)
        n_ref_over = sum(1 for e in ref32.values() if e > tol)
        lines += [f"## {mode.upper()} mode — north-star bar {tol:g}", "",
                  f"* SR: **{rel_l2(SR, SR_ref):.2e}** (ref fp32: {rel_l2(SR32, SR_ref):.2e}); loss: "
                  f"{abs(float(L) - float(L_ref)) / abs(float(L_ref)):.2e} relative; dL/dLR: {rel_l2(dLR, dLR_ref):.2e} "
                  f"(ref fp32: {rel_l2(dLR32, dLR_ref):.2e}"
                  + (f", envelope {rel_l2(dLR_env, dLR_ref):.2e})" if env else ")"),
                  f"* parameter gradients above the flat bar: **{n_over} of {len(errs)}** (ref fp32: {n_ref_over})", "",
                  "| layer family | tensors | median | max | worst tensor | ref fp32 median | ref fp32 max |"
                  + (" envelope median | envelope max |" if env else ""),
                  "|---|---:|---:|---:|---|---:|---:|" + ("---:|---:|" if env else "")]
        for grp, (n, med, mx, worst) in sorted(rows.items()):
            extra = f" {env_rows[grp][1]:.2e} | {env_rows[grp][2]:.2e} |" if env else ""
            lines.append(f"| {grp} | {n} | {med:.2e} | {mx:.2e} | `{worst}` | {ref32_rows[grp][1]:.2e} | "
                         f"{ref32_rows[grp][2]:.2e} |{extra}")
        lines.append("")
        print("\n".join(lines[-(len(rows) + 8):]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", args.out)


if __name__ == "__main__":
    main()
