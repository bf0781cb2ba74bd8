"""Synthetic HARMONIE-SIMRA-shaped batches (no network / dataset in this environment).

Shapes and value ranges are what the reference's ``reformat_to_torch`` (process_data.py:420-494) hands the
training loop: HR wind (B,3,X,Y,Z) smooth in [-1,1]; Z (B,1,X,Y,Z) raw altitude in metres = smooth terrain in
[0,480] m + a stretched 2..68 m above-ground ladder; LR (B,4,X/s,Y/s,Z) = strided subsample of
cat(HR, (Z - Z_MIN)/(Z_MAX - Z_MIN)) (process_data.py:457; constants plot_data.py:32-39); slightly
non-uniform x / y coordinate vectors in metres.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

Z_MIN, Z_MAX = -2.71, 550.44


def _smooth(shape, gen, device, passes=2):
    f = torch.randn(shape, generator=gen, device=device)
    for _ in range(passes):
        n, c = shape[:2]
        kz = 3 if shape[4] >= 3 else 1
        f = F.avg_pool3d(f.reshape(n * c, 1, *shape[2:]), (5, 5, kz), stride=1, padding=(2, 2, kz // 2),
                         count_include_pad=False).reshape(shape)
    return f / f.abs().amax().clamp_min(1e-6)


def make_batch(batch: int, hr_xy: int = 128, nz: int = 10, scale: int = 8, seed: int = 2001, device="cpu"):
    """Returns LR, HR, Z, x, y (fp32) on ``device``."""
    gen = torch.Generator(device=device).manual_seed(seed)
    HR = _smooth((batch, 3, hr_xy, hr_xy, nz), gen, device)
    terrain = (_smooth((batch, 1, hr_xy, hr_xy, 1), gen, device, passes=3) * 0.5 + 0.5) * 480.0
    ladder = 2.0 + 66.0 * torch.linspace(0, 1, nz, device=device) ** 1.5
    Z = terrain + ladder.reshape(1, 1, 1, 1, nz)
    LR = torch.cat((HR, (Z - Z_MIN) / (Z_MAX - Z_MIN)), 1)[:, :, ::scale, ::scale, :].contiguous()
    x = torch.cumsum(200.0 * (1.0 + 0.05 * torch.rand(hr_xy, generator=gen, device=device)), 0)
    y = torch.cumsum(200.0 * (1.0 + 0.05 * torch.rand(hr_xy, generator=gen, device=device)), 0)
    return LR, HR.contiguous(), Z.contiguous(), x, y
