"""Drop-in ``Discriminator_3D``: VGG-style 3-D discriminator (CNN_models/Discriminator_3D.py:15-193 of the
reference).  Same constructor, ``forward(x) -> (N, 1)``, ``features`` / ``classifier`` / ``dropout``
attributes and ``state_dict`` keys.  The ten convolutions, BatchNorm statistics / normalisation and
LeakyReLUs run on ``libwindsr.so``; the two tiny Linear layers (4.1 MFLOP / sample) stay
``nn.Linear``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..tools import loggingclass as lc
from .torch_blocks import Dropout3d, _CastFn, create_conv_lrelu_layer, create_discriminator_block


class _ToContiguousF32(torch.autograd.Function):
    """channels-last activation -> contiguous NCXYZ fp32 (what ``reshape(N, -1)`` at Discriminator_3D.py:192
    flattens), as one strided-copy launch."""

    @staticmethod
    def forward(ctx, x):
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        ops.copy_(x, out)
        ctx.in_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, d):
        g = ops.empty_cl(*d.shape, ctx.in_dtype, d.device)
        ops.copy_(d.contiguous(), g)
        return g


class Discriminator_3D(nn.Module, lc.GlobalLoggingClass):
    def __init__(self, in_channels: int, base_number_of_features: int, feat_kern_size: int = 3,
                 normalization_type: str = "batch", act_type: str = "leakyrelu", mode="CNA", device="cpu",
                 number_of_z_layers=10, conv_mode: str = "3D", use_mixed_precision: bool = False,
                 enable_slicing: bool = False, dropout_probability: float = 0.0):
        super().__init__()
        self.base_number_of_features = B = base_number_of_features
        if act_type == "leakyrelu":
            slope = 0.2
        elif act_type == "relu":
            slope = 0.0
        else:
            self.status_logs.append(f"Discriminator: warning: activation type {act_type} has not been "
                                    "implemented - defaulting to leaky ReLU (0.2)")
            slope = 0.2
        if conv_mode not in ("3D", None):
            raise NotImplementedError(f"conv_mode {conv_mode!r} is outside the hot path (SURVEY §2)")

        # z extent after each of the five stages (Discriminator_3D.py:56-65): untouched until the last one
        zs = [number_of_z_layers]
        for i in range(5):
            if i == 0 and number_of_z_layers <= 19:
                zs.append(number_of_z_layers)
            elif i in (1, 2, 3):
                zs.append(zs[i])
            else:
                zs.append(zs[i] // 2 + zs[i] % 2)

        def block(cin, cout, first=False, halve=False, nz=10):
            return create_discriminator_block(cin, cout, feat_kern_size=feat_kern_size,
                                              lrelu_negative_slope=slope, normalization_type=normalization_type,
                                              drop_first_norm=first, halve_z_dim=halve, number_of_z_layers=nz,
                                              mode="3D")

        features = [
            block(in_channels, B, first=True, halve=number_of_z_layers > 19, nz=zs[0]),
            block(B, 2 * B, nz=zs[1]),
            block(2 * B, 4 * B, nz=zs[2]),
        ]
        if not enable_slicing:
            features.append(block(4 * B, 8 * B, nz=zs[3]))
            features.append(block(8 * B, 8 * B, halve=True, nz=zs[4]))
        else:
            features.append(block(4 * B, 8 * B, nz=zs[3]))
            features.append(create_conv_lrelu_layer(8 * B, 8 * B, feat_kern_size, normalization_type="batch"))
            features.append(create_conv_lrelu_layer(8 * B, 8 * B, feat_kern_size, stride=(1, 1, 2),
                                                    normalization_type="batch"))

        classifier = [nn.Linear(8 * B * 4 * 4 * zs[5], 100), nn.LeakyReLU(negative_slope=slope),
                      nn.Linear(100, 1)]
        self.dropout = Dropout3d(p=dropout_probability if dropout_probability is not None else 0.0)
        self.features = nn.Sequential(*features)
        self.classifier = nn.Sequential(*classifier)
        self.status_logs.append("Discriminator: finished init")

    def forward(self, x):
        x = x if x.dtype in (torch.float32, torch.bfloat16) else x.float()
        h = self.dropout(self.features(x))
        h = _ToContiguousF32.apply(h)
        return self.classifier(h.reshape(h.shape[0], -1))
