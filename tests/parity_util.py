"""Full-size parity harness shared by ``tests/test_gpu_fullsize.py`` and ``scripts/parity_table.py``.

The checker is ``oracle/wind_oracle.py`` executed ON THE GPU in strict fp32 (TF32 off): the same restatement that
``tests/test_oracle_pinned.py`` pins bit-exactly to the reference, with torch's own conv kernels doing the arithmetic —
fast enough (≈1 s per step at B = 8) to check the kernels the benchmark actually runs (CTA-pair hr_convs.0, cost-model
split weight gradients) at the shipped upscale8 size, which the CPU oracle cannot do in minutes.
"""
from __future__ import annotations

import contextlib
import os

import torch

from tests.util import ROOT, _RoundBF16, rel_l2

INI8 = os.path.join(ROOT, "configs", "upscale8_pix4_no_adv_no_slicing.ini")
INI16 = os.path.join(ROOT, "configs", "upscale16_pix4_no_adv_no_slicing.ini")


@contextlib.contextmanager
def strict_fp32():
    """torch / cuDNN without TF32 and without reduced-precision reductions: the fp32 oracle arithmetic."""
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@contextlib.contextmanager
def operand_rounding(kind):
    """Inside: the oracle's conv3d rounds its operands (and incoming gradients) to bf16 / tf32 and accumulates in fp32
    — the minimal mixed-precision scheme; its distance from fp32 is the intrinsic error of that operand format."""
    import torch.nn.functional as F
    if kind is None:
        yield
        return
    orig = F.conv3d

    class _RoundTF32(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32) if x.is_contiguous() else \
                (x.contiguous().view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)

        @staticmethod
        def backward(ctx, g):
            g = g.contiguous()
            return (g.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)

    rnd = _RoundBF16.apply if kind == "bf16" else _RoundTF32.apply

    def conv(x, w, b=None, stride=1, padding=0):
        return orig(rnd(x), rnd(w), b, stride=stride, padding=padding)

    F.conv3d = conv
    try:
        yield
    finally:
        F.conv3d = orig


def shipped_gan(ini=INI8, seed=2001, device="cuda:0"):
    """``wind_field_GAN_3D`` of a shipped ini (upscale8: 128 features, 16 RRDBs, 5^3 HR convs, D base width 32),
    seeded initialisation at the shipped init scales (0.1 / 0.2)."""
    from gan_sr_wind_field_b200.config.config import Config
    from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
    cfg = Config(ini)
    cfg.is_train, cfg.gpu_id, cfg.device = True, 0, torch.device(device)
    torch.manual_seed(seed)
    return wind_field_GAN_3D(cfg), cfg


def loss_weights(cfg):
    t = cfg.training
    return dict(pixel=t.pixel_loss_weight, xy=t.gradient_xy_loss_weight, z=t.gradient_z_loss_weight,
                div=t.divergence_loss_weight, dxy=t.xy_divergence_loss_weight, adv=t.adversarial_loss_weight)


def oracle_generator_step(sd, batch, weights, dropout_scale, rounding=None, dtype=torch.float32):
    """Generator forward + the reference's generator loss + backward with the oracle on the GPU (strict fp32, or
    float64 — the arbiter between two fp32 implementations).  Returns (SR, loss, {name: grad}, dL/dLR)."""
    from oracle import wind_oracle as wo
    LR, HR, Z, x, y = (t.to(dtype) for t in batch)
    if dropout_scale is not None:
        dropout_scale = dropout_scale.to(dtype)
    p = {k: v.detach().to(dtype).requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    lr = LR.detach().clone().requires_grad_(True)
    with strict_fp32(), operand_rounding(rounding):
        SR = wo.generator_forward(p, lr, Z, dropout_scale=dropout_scale)
        total, _ = wo.generator_loss(HR, SR, Z, x, y, weights)
        names = [k for k, v in p.items() if v.requires_grad]
        grads = torch.autograd.grad(total, [lr] + [p[k] for k in names])
    return SR.detach(), total.detach(), dict(zip(names, grads[1:])), grads[0]


def native_generator_step(gan, batch, mode, dropout_scale):
    """The same step through the drop-in modules (this package's CUDA path) in precision ``mode``."""
    from gan_sr_wind_field_b200 import ops
    LR, HR, Z, x, y = batch
    gan.feed_xy_niter(x, y, torch.tensor(100000, device=LR.device), 0, 50)
    G = gan.G
    G.train()
    G.hr_convs[1].sample = lambda n, c, dev: dropout_scale.reshape(-1) if dropout_scale is not None else None
    lr = LR.detach().clone().requires_grad_(True)
    G.zero_grad(set_to_none=True)
    t = gan.cfg.training
    with ops.precision(mode):
        SR = G(lr, Z)
        pix, xy, zg, div, dxy = gan.wind_loss_terms(HR, SR, Z)
        total = (pix * t.pixel_loss_weight + xy * t.gradient_xy_loss_weight + zg * t.gradient_z_loss_weight
                 + div * t.divergence_loss_weight + dxy * t.xy_divergence_loss_weight)
        total.backward()
    torch.cuda.synchronize()
    del G.hr_convs[1].sample
    return SR.detach(), total.detach(), {k: p.grad for k, p in G.named_parameters()}, lr.grad


def group_of(name: str) -> str:
    """Coarse layer family of a generator parameter (rows of the parity table)."""
    if name.startswith("hr_convs.0"):
        return "G7 hr_convs.0 (5^3 144->144)"
    if name.startswith("hr_convs.2"):
        return "G8 hr_convs.2 (5^3 144->3)" + (" bias" if name.endswith("bias") else "")
    if name.startswith("terrain_convs.0"):
        return "G6a terrain 1->16"
    if name.startswith("terrain_convs.1"):
        return "G6b terrain 16->16"
    if name.startswith("model.0."):
        return "G1 feature_conv"
    if ".RDBs." in name:
        if "LFF" in name:
            return "G3 LFF " + ("bias" if name.endswith("bias") else "weight")
        return "G2 RDB dense conv"
    if name.startswith("model.1.module."):
        return "G4 lr_conv"
    return "G5 UpConv " + name.split(".")[1]


def summarize(errs: dict):
    """{group: (count, median, max, argmax-name)}"""
    out = {}
    for k, e in errs.items():
        out.setdefault(group_of(k), []).append((e, k))
    rows = {}
    for g, lst in out.items():
        lst.sort()
        rows[g] = (len(lst), lst[len(lst) // 2][0], lst[-1][0], lst[-1][1])
    return rows


def grad_errors(mine: dict, ref: dict):
    return {k: rel_l2(mine[k], ref[k]) for k in ref if mine.get(k) is not None}
