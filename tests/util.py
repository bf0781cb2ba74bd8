"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north-star tolerances (BASELINE.json): relative L2 against the fp32 reference
TOL = {"fp32": 1e-5, "tf32": 2e-3, "bf16": 2e-2}


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def sd_from(npz, prefix):
    out = {}
    for k in npz.files:
        if k.startswith(prefix):
            a = npz[k]
            t = torch.from_numpy(a.astype(np.float32) if a.dtype == np.float16 else a.copy())
            out[k[len(prefix):]] = t
    return out


def rel_l2(a, b) -> float:
    a = torch.as_tensor(a).detach().double().flatten().cpu()
    b = torch.as_tensor(b).detach().double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def small_generator_kwargs():
    return dict(in_channels=4, out_channels=3, number_of_features=16, number_of_RRDBs=2, upscale=4, hr_kern_size=5,
                number_of_RDB_convs=5, RDB_gc=8, lff_kern_size=1, terrain_number_of_features=8,
                dropout_probability=0.0)


# ---- intrinsic error of bf16 operands -------------------------------------------------------------------
class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class bf16_operand_emulation:
    """Context manager: inside it the oracle's conv3d rounds its operands (and, in backward, the incoming
    gradient) to bf16 and accumulates in fp32 — the *minimal* bf16 mixed-precision scheme, on the CPU.
    Its distance from the fp32 reference is the error any bf16 tensor-core path must incur on a given net; the
    BF16-mode GPU tests bound the CUDA path by that envelope where the north star's flat 2e-2 is not attainable
    by ANY bf16 implementation (deep gradient chains of the deliberately hot golden nets)."""

    def __enter__(self):
        import torch.nn.functional as F
        self._F, self._orig = F, F.conv3d
        orig = self._orig

        def conv(x, w, b=None, stride=1, padding=0):
            return orig(_RoundBF16.apply(x), _RoundBF16.apply(w), b, stride=stride, padding=padding)

        F.conv3d = conv
        return self

    def __exit__(self, *a):
        self._F.conv3d = self._orig
