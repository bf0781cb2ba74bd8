"""Bisect a failing CUDA-graph replay of the tiny GAN's step: each variant runs in its own process.
usage: python scripts/debug/graph_bisect.py [variant] [mode]   (no variant: run them all as subprocesses)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
VARIANTS = ["D_full", "D_no_opt", "D_fwd_only", "G_eval_fwd_only", "D_no_noise", "G_full", "G_no_opt", "G_fwd_only",
            "G_fwd_loss", "G_no_D"]

def main(variant, mode):
    import torch
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.GAN_models import graph_step
    from tests.test_gpu_round2 import _tiny_gan
    gan, (LR, HR, Z) = _tiny_gan(True, noise=variant != "D_no_noise")
    kind = variant[0]
    if variant.endswith("no_opt"):
        (gan.optimizer_D if kind == "D" else gan.optimizer_G).step = lambda *a, **k: None
    if variant == "D_fwd_only":
        def upd(HR_, fake, it, train):
            y, f = gan.D_forward(HR_, fake, it, train_D=True)
            gan.D_loss_dict["train_loss"] = gan._adversarial(y, f, False).detach()
        gan.update_D = upd
    if variant == "G_eval_fwd_only":
        gan.update_D = lambda HR_, fake, it, train: gan.D_loss_dict.__setitem__("train_loss", fake.mean().detach())
    if variant == "G_fwd_only":
        def updG(LR_, HR_, Z_, it, train):
            gan.train_G_loss_dict["total"] = gan.G(LR_, Z_).mean().detach()
        gan.update_G = updG
    if variant == "G_fwd_loss":
        def updG(LR_, HR_, Z_, it, train):
            SR = gan.G(LR_, Z_)
            gan.train_G_loss_dict["total"] = sum(gan.wind_loss_terms(HR_, SR, Z_)).detach()
        gan.update_G = updG
    if variant == "G_no_D":
        gan.cfg.training.adversarial_loss_weight = 0.0
    gan.is_G_iteration = lambda it: kind == "G"
    with ops.precision(mode):
        for it in range(1, 8):
            gan.optimize_parameters(LR, HR, Z, it)
            torch.cuda.synchronize()
    g = [v for v in gan._graphs.values()]
    print(variant, mode, "OK", "graph" if g and g[0] else "NO GRAPH", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "fp32")
    else:
        for mode in (sys.argv[2:] or ["fp32", "bf16"]) if False else ("fp32", "bf16"):
            for v in VARIANTS:
                r = subprocess.run([sys.executable, __file__, v, mode], capture_output=True, text=True, timeout=300)
                tail = (r.stdout.strip().splitlines() or [""])[-1]
                err = [l for l in r.stderr.splitlines() if "Error" in l or "error" in l][:2]
                print(f"{v:18s} {mode}: rc={r.returncode} {tail} {err}", flush=True)
