"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (see wind_oracle.py).  Restatement of the reference's per-sample input
pipeline and validation metrics in numpy / torch CPU ops:

==============================================  ==========================================================
function here                                   reference lines it follows
==============================================  ==========================================================
``reformat``                                    process_data.py:420-494 (reformat_to_torch)
``augment``                                     process_data.py:198-262 (CustomizedDataset.__getitem__)
``crop``                                        process_data.py:159-176 + download_data.py slice_only_dim_dicts
``psnr`` / ``validation_metrics``               GAN_models/wind_field_GAN_3D.py:730-770, 597-618
==============================================  ==========================================================

Pinned by ``tests/test_oracle_pinned.py`` against the live reference (``reformat_to_torch`` is imported and run on
the same inputs) and by ``tests/golden/prepare_batch.npz`` (written by make_golden.py from the reference).
"""
from __future__ import annotations

import numpy as np
import torch


def crop(fields, x_start, y_start, size):
    """x/y slice [start, start+size), all z (process_data.py:159-176)."""
    return [None if f is None else f[x_start:x_start + size, y_start:y_start + size, :] for f in fields]


def reformat(u, v, w, p, z, z_above_ground, Z_MIN, Z_MAX, Z_ABOVE_GROUND_MAX, UVW_MAX, P_MIN, P_MAX,
             coarseness_factor=4, include_pressure=False, include_z_channel=False,
             include_above_ground_channel=False):
    """float64 fields (X,Y,Z) -> LR (C,x,y,Z), HR (3,X,Y,Z), Z (1,X,Y,Z) float32 tensors."""
    cf = coarseness_factor
    HR = np.stack((u, v, w), 0) / UVW_MAX
    LR = HR
    if include_pressure:
        LR = np.concatenate((LR, ((p - P_MIN) / (P_MAX - P_MIN))[None]), 0)
    LR = LR[:, ::cf, ::cf, :]
    if include_z_channel:
        if include_above_ground_channel:
            LR = np.concatenate((LR, (z_above_ground[None, ::cf, ::cf, :]) / Z_ABOVE_GROUND_MAX,
                                 (z - z_above_ground - Z_MIN)[None, ::cf, ::cf, :]
                                 / (Z_MAX - Z_MIN - Z_ABOVE_GROUND_MAX)), 0)
        else:
            LR = np.concatenate((LR, (z[None, ::cf, ::cf, :] - Z_MIN) / (Z_MAX - Z_MIN)), 0)
    return (torch.from_numpy(np.ascontiguousarray(LR)).float(), torch.from_numpy(np.ascontiguousarray(HR)).float(),
            torch.from_numpy(np.ascontiguousarray(z[None])).float())


def augment(LR, HR, Z, rotations: int, flip_x: bool, flip_y: bool):
    """rot90 in the x/y plane with the horizontal wind components re-expressed in the rotated frame, then the two
    mirror flips with their sign changes (process_data.py:198-262)."""
    k = int(rotations) % 4
    LR, HR, Z = (torch.rot90(t, k, [1, 2]).clone() for t in (LR, HR, Z))
    for t in (HR, LR):
        u, v = t[0].clone(), t[1].clone()
        if k == 1:
            t[0], t[1] = -v, u
        elif k == 2:
            t[0], t[1] = -u, -v
        elif k == 3:
            t[0], t[1] = v, -u
    if flip_x:
        LR, HR, Z = (torch.flip(t, [1]).clone() for t in (LR, HR, Z))
        LR[0], HR[0] = -LR[0], -HR[0]
    if flip_y:
        LR, HR, Z = (torch.flip(t, [2]).clone() for t in (LR, HR, Z))
        LR[1], HR[1] = -LR[1], -HR[1]
    return LR, HR, Z


def psnr(HR, fake, max_diff_squared=4.0, eps=1e-8):
    vox = HR.shape[0] * HR.shape[2] * HR.shape[3] * HR.shape[4]
    mse = torch.sum((HR - fake) ** 2) / vox
    return 10.0 * torch.log10(max_diff_squared / (mse + eps))


def validation_metrics(LR, HR, SR, scale, pixel="l1"):
    """(PSNR of SR, PSNR of the trilinear upsample, trilinear pixel loss) — wind_field_GAN_3D.py:597-618."""
    import torch.nn.functional as F
    tri = F.interpolate(LR[:, :3], scale_factor=(scale, scale, 1), mode="trilinear", align_corners=True)
    pix = F.l1_loss(HR, tri) if pixel != "l2" else F.mse_loss(HR, tri)
    return psnr(HR, SR), psnr(HR, tri), pix
