"""Minimal capture + replay of the tiny GAN's D and G steps (for compute-sanitizer runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from gan_sr_wind_field_b200 import ops
from tests.test_gpu_round2 import _tiny_gan
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
gan, (LR, HR, Z) = _tiny_gan(True)
with ops.precision(mode):
    for it in range(1, 13):
        gan.optimize_parameters(LR, HR, Z, it)
        torch.cuda.synchronize()
        print("it", it, "G" if gan.is_G_iteration(it) else "D", "graphs:", {k[0]: bool(v) for k, v in gan._graphs.items()}, flush=True)
print("ok")
