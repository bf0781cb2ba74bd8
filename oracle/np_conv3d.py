"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (see wind_oracle.py header).

Independent float64 numpy restatement of the Conv3d forward, data-gradient and weight-gradient that the
reference obtains from ``torch.nn.Conv3d`` + autograd (CNN_models/torch_blocks.py:17,278;
CNN_models/Generator_3D_Resnet_ESRGAN.py:105), i.e. of PyTorch's published cross-correlation definition

    out[n, co, xo, yo, zo] = sum_{ci, i, j, l} w[co, ci, i, j, l] * in[n, ci, xo*sx - px + i, yo*sy - py + j, zo*sz - pz + l]

with zero padding.  Plain loops over the kernel taps, whole-array numpy per tap: small cases only.
It pins the semantics (tap order, stride, padding, layout) of the torch primitives ``wind_oracle`` builds on.
"""
from __future__ import annotations

import numpy as np


def _t3(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def out_size(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def conv3d_fwd(x, w, stride=1, padding=0):
    x, w = np.asarray(x, np.float64), np.asarray(w, np.float64)
    (sx, sy, sz), (px, py, pz) = _t3(stride), _t3(padding)
    N, Ci, X, Y, Z = x.shape
    Co, _, kx, ky, kz = w.shape
    Xo, Yo, Zo = out_size(X, kx, sx, px), out_size(Y, ky, sy, py), out_size(Z, kz, sz, pz)
    xp = np.pad(x, ((0, 0), (0, 0), (px, px), (py, py), (pz, pz)))
    out = np.zeros((N, Co, Xo, Yo, Zo))
    for i in range(kx):
        for j in range(ky):
            for l in range(kz):
                patch = xp[:, :, i:i + (Xo - 1) * sx + 1:sx, j:j + (Yo - 1) * sy + 1:sy, l:l + (Zo - 1) * sz + 1:sz]
                out += np.einsum("ncxyz,oc->noxyz", patch, w[:, :, i, j, l])
    return out


def conv3d_dgrad(dy, w, x_shape, stride=1, padding=0):
    dy, w = np.asarray(dy, np.float64), np.asarray(w, np.float64)
    (sx, sy, sz), (px, py, pz) = _t3(stride), _t3(padding)
    N, Ci, X, Y, Z = x_shape
    Co, _, kx, ky, kz = w.shape
    _, _, Xo, Yo, Zo = dy.shape
    dxp = np.zeros((N, Ci, X + 2 * px, Y + 2 * py, Z + 2 * pz))
    for i in range(kx):
        for j in range(ky):
            for l in range(kz):
                dxp[:, :, i:i + (Xo - 1) * sx + 1:sx, j:j + (Yo - 1) * sy + 1:sy, l:l + (Zo - 1) * sz + 1:sz] += \
                    np.einsum("noxyz,oc->ncxyz", dy, w[:, :, i, j, l])
    return dxp[:, :, px:px + X, py:py + Y, pz:pz + Z]


def conv3d_wgrad(x, dy, w_shape, stride=1, padding=0):
    x, dy = np.asarray(x, np.float64), np.asarray(dy, np.float64)
    (sx, sy, sz), (px, py, pz) = _t3(stride), _t3(padding)
    Co, Ci, kx, ky, kz = w_shape
    _, _, Xo, Yo, Zo = dy.shape
    xp = np.pad(x, ((0, 0), (0, 0), (px, px), (py, py), (pz, pz)))
    dw = np.zeros(w_shape)
    for i in range(kx):
        for j in range(ky):
            for l in range(kz):
                patch = xp[:, :, i:i + (Xo - 1) * sx + 1:sx, j:j + (Yo - 1) * sy + 1:sy, l:l + (Zo - 1) * sz + 1:sz]
                dw[:, :, i, j, l] = np.einsum("noxyz,ncxyz->oc", dy, patch)
    return dw


def upsample_nearest_xy(x):
    return np.repeat(np.repeat(np.asarray(x), 2, axis=2), 2, axis=3)
