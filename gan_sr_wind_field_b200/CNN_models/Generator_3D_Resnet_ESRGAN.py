"""Drop-in ``Generator_3D``: the 3-D ESRGAN-style RRDB generator with terrain feature extraction.

Constructor, ``forward(x, Z)``, attribute surface (``model``, ``hr_convs``, ``terrain_convs``, ``max_norm``,
``status_logs``) and ``state_dict`` keys follow CNN_models/Generator_3D_Resnet_ESRGAN.py:23-229 of the
reference; the arithmetic runs on the sm_100a kernels of ``libwindsr.so``.

Fused dataflow of ``forward`` (reference lines in brackets):

* ``model[0]`` feature conv [78-85] reads the NCXYZ fp32 LR volume in place and writes the fp32 trunk state;
* 16 RRDBs [183-198]: each RDB is one autograd node whose five convs share one concat buffer; residual
  scalings and skip adds live in the LFF epilogue; ``lr_conv`` [86-94] adds the long skip in its epilogue;
* UpConv stages [201-218]: vectorised nearest upsample + conv/LeakyReLU; the last one writes channels
  [0, F) of the (F+T)-channel HR buffer, ``terrain_convs`` [120-137] write channels [F, F+T): the
  ``torch.cat`` of line 228 costs nothing;
* ``hr_convs`` [95-111]: 5x5x5 conv + LeakyReLU with the Dropout3d channel scale folded into the epilogue,
  then the biased 5x5x5 conv that emits the NCXYZ fp32 result directly.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from .. import ops
from ..tools import loggingclass as lc
from .torch_blocks import (RRDB, ConvBlock, Conv3d, Dropout3d, SkipConnectionBlock, UpConvBlock, _CastFn,
                           create_conv_lrelu_layer, create_UpConv_block)


class Generator_3D(nn.Module, lc.GlobalLoggingClass):
    def __init__(self, in_channels: int, out_channels: int, number_of_features: int, number_of_RRDBs: int,
                 upscale: int = 4, hr_kern_size: int = 3, number_of_RDB_convs: int = 5, RDB_gc: int = 32,
                 lff_kern_size: int = 1, RDB_residual_scaling: float = 0.2, RRDB_residual_scaling: float = 0.2,
                 act_type: str = "leakyrelu", number_of_z_layers: int = 10, conv_mode: str = "3D",
                 use_mixed_precision: bool = False, device="cpu", terrain_number_of_features: int = 16,
                 dropout_probability: float = 0.0, max_norm: float = 1.0):
        super().__init__()
        if act_type == "leakyrelu":
            slope = 0.2
        elif act_type == "relu":
            slope = 0.0
        else:
            self.status_logs.append(f"Generator: warning: activation type {act_type} has not been implemented "
                                    "- defaulting to leaky ReLU (0.2)")
            slope = 0.2
        self.max_norm = max_norm
        if conv_mode in ("2D", "horizontal_3D"):
            raise NotImplementedError(f"conv_mode {conv_mode!r} is outside the hot path (SURVEY §2): the shipped "
                                      "configs all use the 3D mode")
        if conv_mode not in ("3D", None):
            raise ValueError(f"Conv mode {conv_mode} not implemented")
        if dropout_probability is None:
            dropout_probability = 0.0
        F, T = number_of_features, terrain_number_of_features
        hr_pad = (hr_kern_size - 1) // 2

        # construction order == the reference's, so a seeded build draws identical initial parameters
        dropout = Dropout3d(p=dropout_probability)
        feature_conv = create_conv_lrelu_layer(in_channels, F, 3, padding=1, lrelu=False)
        lr_conv = create_conv_lrelu_layer(F, F, 3, padding=1, lrelu_negative_slope=slope, lrelu=False)
        hr_convs = [
            create_conv_lrelu_layer(F + T, F + T, kernel_size=hr_kern_size, padding=hr_pad,
                                    lrelu_negative_slope=slope),
            dropout,
            Conv3d(F + T, out_channels, kernel_size=hr_kern_size, padding=hr_pad),
        ]
        terrain_convs = [
            create_conv_lrelu_layer(1, T, 3, padding=1, lrelu=True),
            create_conv_lrelu_layer(T, T, 3, padding=1, lrelu=False),
        ]
        rrdbs = [RRDB(F, RDB_gc, number_of_RDB_convs, lff_kern_size, lrelu_negative_slope=slope,
                      RDB_residual_scaling=RDB_residual_scaling, RRDB_residual_scaling=RRDB_residual_scaling,
                      mode="3D") for _ in range(number_of_RRDBs)]
        shortcut = SkipConnectionBlock(nn.Sequential(*rrdbs, lr_conv))

        n_up = math.floor(math.log2(upscale))
        if 2 ** n_up != upscale:
            self.status_logs.append(f"ESRDnet: warning: upsampling only supported for factors 2^n. Defaulting "
                                    f"{upscale} to {2 ** n_up}")
        upsampler = [create_UpConv_block(F, F, scale=2, lrelu_negative_slope=slope,
                                         number_of_z_layers=number_of_z_layers, mode="3D") for _ in range(n_up)]

        self.model = nn.Sequential(feature_conv, shortcut, *upsampler)
        self.hr_convs = nn.Sequential(*hr_convs)
        self.terrain_convs = nn.Sequential(*terrain_convs)
        self.status_logs.append("Generator: finished init")

    # -- structure probe: is this still the layout __init__ built? (callers may have sliced things apart) ----
    def _fusable(self) -> bool:
        m = self.model
        if len(m) < 2 or not isinstance(m[0], ConvBlock) or not isinstance(m[1], SkipConnectionBlock):
            return False
        if not all(isinstance(b, UpConvBlock) and len(b) == 2 for b in list(m)[2:]):
            return False
        h = self.hr_convs
        return (len(h) == 3 and isinstance(h[0], ConvBlock) and isinstance(h[1], Dropout3d)
                and isinstance(h[2], Conv3d) and len(self.terrain_convs) == 2
                and all(isinstance(t, ConvBlock) for t in self.terrain_convs))

    def forward(self, x, Z):
        if not self._fusable():
            x = self.model(x)
            Z = self.terrain_convs(Z)
            buf = torch.cat((x.float(), Z.float()), dim=1)
            return self.hr_convs(buf)

        cdt = ops.act_dtype()
        x = x if x.dtype == torch.float32 else x.float()
        Z = Z if Z.dtype == torch.float32 else Z.float()
        n = x.shape[0]
        # low-level features + RRDB trunk in an fp32 residual stream
        h = self.model[0][0].run(x, out_dtype=torch.float32)
        h = self.model[1](h)
        ups = list(self.model)[2:]
        F = self.model[0][0].out_channels
        T = self.terrain_convs[1][0].out_channels
        if h.dtype != cdt:
            h = _CastFn.apply(h, cdt)
        X, Y, Zn = h.shape[2] << len(ups), h.shape[3] << len(ups), h.shape[4]
        buf = ops.empty_cl(n, F + T, X, Y, Zn, cdt, x.device)
        a = h
        for i, blk in enumerate(ups):
            a = blk[0](a)
            last = i == len(ups) - 1
            a = blk[1](a, out=buf[:, :F]) if last else blk[1](a)
        if not ups:
            a = _copy_into(a, buf[:, :F])
        t = self.terrain_convs[0](Z)
        t = self.terrain_convs[1](t, out=buf[:, F:])
        feat = ops.CatFn.apply(a, t, buf)
        scale = self.hr_convs[1].sample(n, F + T, x.device)
        feat = self.hr_convs[0](feat, chan_scale=scale)
        last = self.hr_convs[2]
        if (ops.get_precision() == "bf16" and last.stride == (1, 1, 1) and last.kernel_size[0] * last.out_channels <= 16
                and last.in_channels >= 16 and feat.dtype == torch.bfloat16):
            # very narrow output (144 -> 3): lateral taps folded into the channel dimension (ops.XYFoldConvFn / XFoldConvFn)
            if (last.out_channels <= 8 and last.kernel_size[0] * last.kernel_size[1] * last.out_channels <= 128
                    and os.environ.get("WINDSR_XYFOLD", "1") != "0"):
                return ops.XYFoldConvFn.apply(feat, last.weight, last.bias, last.padding)
            return ops.XFoldConvFn.apply(feat, last.weight, last.bias, last.padding)
        return last.run(feat, out_dtype=torch.float32, out_contig=True)


def _copy_into(src, dst):
    from .torch_blocks import _CopyIntoFn
    return _CopyIntoFn.apply(src, dst)
