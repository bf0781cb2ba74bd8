"""GPU input pipeline: what ``CustomizedDataset.__getitem__`` + the DataLoader collate do per
batch on CPU workers (process_data.py:122-262, 420-494), done on the device by ONE gather kernel
(``ws_prepare_batch``, csrc/train_aux.cu) from the float64 per-hour fields of the reference's on-disk format
(download_data.py:456-467: ``[z, z_above_ground, u, v, w, pressure]``, each (X, Y, Z) float64).

The random choices follow the reference's distributions and order per sample — slice offsets
``round(beta(0.25, 0.25) * (size - slice_size))`` for x then y (process_data.py:159-166), ``randint(0, 4)`` quarter
turns (:199), two ``rand() > 0.5`` mirror flips (:246, :252) — drawn from a ``numpy.random.Generator`` on the host
(a few integers per sample) and shipped as one small int32 tensor; the arithmetic (crop, normalisation in float64,
LR subsampling, rot90 / flip index maps and the u/v sign fixes) is bit-exact against the reference
(tests/test_gpu_round2.py::test_prepare_batch_bit_exact_vs_reference_golden).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops


class DeviceWindDataset:
    """Holds the hourly fields of a data split resident in HBM (float64, (T, X, Y, Z) each) and cuts training
    batches out of them on the device.  Constructor arguments mirror ``CustomizedDataset`` (process_data.py:27-51)."""

    def __init__(self, z, z_above_ground, u, v, w, pressure, Z_MIN, Z_MAX, UVW_MAX, P_MIN, P_MAX, Z_ABOVE_GROUND_MAX,
                 include_pressure=False, include_z_channel=False, include_above_ground_channel=False,
                 COARSENESS_FACTOR=4, data_aug_rot=True, data_aug_flip=True, enable_slicing=False, slice_size=64,
                 device="cuda", seed: Optional[int] = None):
        as_dev = lambda a: None if a is None else torch.as_tensor(np.asarray(a), dtype=torch.float64).to(device).contiguous()
        self.z, self.zag, self.u, self.v, self.w, self.p = (as_dev(a) for a in (z, z_above_ground, u, v, w, pressure))
        self.norm = dict(uvw_max=UVW_MAX, p_min=P_MIN, p_max=P_MAX, z_min=Z_MIN, z_max=Z_MAX,
                         z_above_ground_max=Z_ABOVE_GROUND_MAX)
        self.flags = dict(include_pressure=include_pressure, include_z_channel=include_z_channel,
                          include_above_ground_channel=include_above_ground_channel)
        self.coarseness = int(COARSENESS_FACTOR)
        self.rot, self.flip = bool(data_aug_rot), bool(data_aug_flip)
        self.enable_slicing, self.slice_size = bool(enable_slicing), int(slice_size)
        self.rng = np.random.default_rng(seed)

    def __len__(self):
        return self.u.shape[0]

    def sample_augmentation(self, n: int) -> np.ndarray:
        """(n, 5) int32: x_start, y_start, quarter turns, flip_x, flip_y — the reference's draws, in its order."""
        X, Y = self.u.shape[1], self.u.shape[2]
        out = np.zeros((n, 5), dtype=np.int32)
        for i in range(n):
            if self.enable_slicing:
                out[i, 0] = round(self.rng.beta(0.25, 0.25) * (X - self.slice_size))
                out[i, 1] = round(self.rng.beta(0.25, 0.25) * (Y - self.slice_size))
            if self.rot:
                out[i, 2] = self.rng.integers(0, 4)
            if self.flip:
                out[i, 3] = int(self.rng.random() > 0.5)
                out[i, 4] = int(self.rng.random() > 0.5)
        return out

    def batch(self, indices: Sequence[int], aug: Optional[np.ndarray] = None):
        """(LR, HR, Z) float32 device tensors for the given sample indices."""
        idx = torch.as_tensor(list(indices), dtype=torch.long, device=self.u.device)
        if aug is None:
            aug = self.sample_augmentation(len(idx))
        pick = lambda t: None if t is None else t.index_select(0, idx)
        crop = (self.slice_size, self.slice_size) if self.enable_slicing else None
        return ops.prepare_batch(pick(self.u), pick(self.v), pick(self.w), pick(self.z), pressure=pick(self.p),
                                 z_above_ground=pick(self.zag), aug=torch.from_numpy(np.ascontiguousarray(aug)),
                                 crop=crop, coarseness=self.coarseness, **self.flags, **self.norm)
