"""ctypes binding of the C-ABI in ``include/windsr.h`` (``libwindsr.so``).

The product path has NO fallback: if the shared library is missing (or a call fails) this module raises —
nothing routes to torch/cuDNN or to the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libwindsr.so")

WS_F32, WS_BF16 = 0, 1
MATH_FP32, MATH_TF32, MATH_BF16 = 0, 1, 2
PATH_NONE, PATH_SIMT, PATH_TCGEN05 = 0, 1, 2
PACK_SIMT_FWD, PACK_SIMT_DGRAD, PACK_TC_FWD, PACK_TC_DGRAD, PACK_TC_FWD_TF32, PACK_TC_DGRAD_TF32 = 0, 1, 2, 3, 4, 5
WL_SLOTS, WL_RESULT_FLOATS, WLB_SLOTS = 16, 64, 12


class WsTensor(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dtype", C.c_int32), ("_pad", C.c_int32), ("nstride", C.c_int64),
                ("vstride", C.c_int64), ("cstride", C.c_int64)]


class WsConvShape(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n", "x", "y", "z", "cin", "cout", "kx", "ky", "kz", "sx", "sy", "sz",
                                         "px", "py", "pz")]


class WsEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("oscale", C.c_void_p), ("chan_scale", C.c_void_p),
                ("lrelu_slope", C.c_float), ("alpha", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("res1", WsTensor), ("res2", WsTensor), ("mask", WsTensor),
                ("mask_c0", C.c_int32), ("mask_c1", C.c_int32), ("mask_slope", C.c_float), ("flags", C.c_int32),
                ("out2", WsTensor), ("stat_sum", C.c_void_p), ("stat_sqsum", C.c_void_p),
                ("tail_out", WsTensor), ("tail_mask", WsTensor), ("tail_c0", C.c_int32), ("tail_slope", C.c_float)]


class WsRdbDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n", "x", "y", "z", "f", "gc", "nconv", "k", "k_lff")] + \
               [(k, C.c_float) for k in ("slope", "alpha", "beta1", "beta2")] + \
               [("math", C.c_int32), ("repack", C.c_int32)]


class WsPrepareDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n", "sx", "sy", "sz", "x", "y", "coarseness", "include_pressure",
                                         "include_z_channel", "include_above_ground_channel")] + \
               [("sample_stride", C.c_int64)] + \
               [(k, C.c_double) for k in ("uvw_max", "p_min", "p_max", "z_min", "z_max", "z_above_ground_max")]


class WindSRError(RuntimeError):
    pass


_lib = None

# every symbol include/windsr.h declares: (name, restype, argtypes)
_P, _I, _F, _L, _Z = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t
_TP = C.POINTER(WsTensor)
_SP = C.POINTER(WsConvShape)
_EP = C.POINTER(WsEpilogue)
SYMBOLS = [
    ("ws_version", _I, []),
    ("ws_last_error", C.c_char_p, []),
    ("ws_device_supports_tcgen05", _I, []),
    ("ws_launch_count", _L, []),
    ("ws_packed_weight_bytes", _Z, [_SP, _I]),
    ("ws_pack_weights", _I, [_P, _SP, _I, _P, _P]),
    ("ws_conv3d_fwd_path", _I, [_SP, _TP, _TP, _I]),
    ("ws_conv3d_dgrad_path", _I, [_SP, _TP, _TP, _I]),
    ("ws_conv3d_wgrad_path", _I, [_SP, _TP, _TP, _I]),
    ("ws_conv3d_fwd", _I, [_SP, _TP, _P, _TP, _EP, _I, _P]),
    ("ws_conv3d_dgrad", _I, [_SP, _TP, _P, _TP, _EP, _I, _P]),
    ("ws_conv3d_wgrad_workspace_bytes", _Z, [_SP, _I]),
    ("ws_conv3d_wgrad", _I, [_SP, _TP, _TP, _P, _P, _I, _I, _P, _Z, _P]),
    ("ws_im2col", _I, [C.POINTER(WsConvShape), _TP, _TP, _I, _P]),
    ("ws_rdb_packed_bytes", _Z, [C.POINTER(WsRdbDesc), _I, _I]),
    ("ws_rdb_backward_workspace_bytes", _Z, [C.POINTER(WsRdbDesc)]),
    ("ws_rdb_forward_workspace_bytes", _Z, [C.POINTER(WsRdbDesc)]),
    ("ws_rdb_forward", _I, [C.POINTER(WsRdbDesc), _TP, _TP, _TP, _TP, C.POINTER(_P), C.POINTER(_P), _P, _P, _Z, _P]),
    ("ws_rdb_backward", _I, [C.POINTER(WsRdbDesc), _TP, _TP, _TP, _TP, _TP, _TP, C.POINTER(_P), C.POINTER(_P),
                             C.POINTER(_P), _P, _P, _Z, _P, _P, _P, _Z]),
    ("ws_trunk_wgrad_supported", _I, [C.POINTER(WsRdbDesc)]),
    ("ws_trunk_wgrad_record_floats", _Z, [C.POINTER(WsRdbDesc)]),
    ("ws_trunk_wgrad_workspace_bytes", _Z, [C.POINTER(WsRdbDesc), _I]),
    ("ws_trunk_wgrad", _I, [C.POINTER(WsRdbDesc), _I, _TP, _TP, _TP, _P, _L, _P, _Z, _P]),
    ("ws_trunk_repack_fwd", _I, [C.POINTER(WsRdbDesc), _I, _TP, _TP, _TP, C.POINTER(_P), C.POINTER(_P), _P]),
    ("ws_trunk_repack_bwd", _I, [C.POINTER(WsRdbDesc), _I, _TP, _TP, _TP, _TP, _TP, C.POINTER(_P), C.POINTER(_P), _P]),
    ("ws_upsample_nearest_xy_fwd", _I, [_TP, _TP, _I, _I, _I, _I, _I, _P]),
    ("ws_upsample_nearest_xy_bwd", _I, [_TP, _TP, _I, _I, _I, _I, _I, _P]),
    ("ws_xfold_sum", _I, [_TP, _P, _TP, _I, _I, _I, _I, _I, _I, _I, _P]),
    ("ws_xunfold", _I, [_TP, _TP, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    ("ws_xyfold_sum", _I, [_TP, _P, _TP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    ("ws_xyunfold", _I, [_TP, _TP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    ("ws_copy", _I, [_TP, _TP, _I, _I, _L, _P]),
    ("ws_axpby", _I, [_TP, _F, _TP, _F, _TP, _I, _I, _L, _P]),
    ("ws_lrelu_bwd", _I, [_TP, _TP, _F, _P, _P, _TP, _I, _I, _L, _P]),
    ("ws_bn_finalize", _I, [_P, _P, _L, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    ("ws_scale_shift_lrelu", _I, [_TP, _P, _P, _F, _TP, _I, _I, _L, _P]),
    ("ws_bn_lrelu_bwd_reduce", _I, [_TP, _TP, _TP, _P, _P, _F, _P, _P, _I, _I, _L, _P]),
    ("ws_bn_lrelu_bwd_apply", _I, [_TP, _TP, _TP, _P, _P, _P, _P, _P, _F, _L, _TP, _I, _I, _L, _P]),
    ("ws_axis_coeffs", _I, [_P, _I, _P, _P]),
    ("ws_wind_gradient", _I, [_TP, _TP, _P, _P, _TP, _I, _I, _I, _I, _P]),
    ("ws_windloss_fwd", _I, [_TP, _TP, _TP, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    ("ws_windloss_bwd_workspace_bytes", _Z, [_I, _I, _I, _I]),
    ("ws_windloss_bwd", _I, [_TP, _TP, _TP, _P, _P, _I, _I, _I, _I, _P, _P, _TP, _P, _Z, _P]),
    ("ws_adam_chunk_elems", _I, []),
    ("ws_adam_step", _I, [_P, _P, _I, _I, _P, _F, C.c_double, C.c_double, _F, _F, _F, _P, _P]),
    ("ws_instance_noise", _I, [_P, _P, _L, _F, _P, C.c_uint64, _P, _P]),
    ("ws_validation_metrics", _I, [_TP, _TP, _TP, _I, _I, _I, _I, _I, _I, _P, _P]),
    ("ws_prepare_batch", _I, [C.POINTER(WsPrepareDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
]


def load():
    """Load libwindsr.so (once). Raises WindSRError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WindSRError(
            f"{LIB_PATH} not found: build the CUDA extension first (python -m gan_sr_wind_field_b200.build or "
            "__graft_entry__.build()). There is no CPU / cuDNN fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.ws_version() != 1:
        raise WindSRError(f"libwindsr.so version {lib.ws_version()} != 1")
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().ws_last_error()
        raise WindSRError(f"{what} failed ({status}): {msg.decode() if msg else ''}")


def stream_ptr() -> int:
    """Raw cudaStream_t of torch's current stream on the current device (fast path: no Stream object)."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:  # pragma: no cover - older torch
        return torch.cuda.current_stream().cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return WS_F32
    if t.dtype == torch.bfloat16:
        return WS_BF16
    raise WindSRError(f"unsupported activation dtype {t.dtype}")


def view(t: torch.Tensor) -> WsTensor:
    """ws_tensor view of a logical (N, C, X, Y, Z) tensor of ANY memory format whose voxel index is linear
    (contiguous NCXYZ, channels_last_3d, or a channel slice of either)."""
    if t.dim() != 5:
        raise WindSRError(f"expected a 5-D (N,C,X,Y,Z) tensor, got shape {tuple(t.shape)}")
    n, c, x, y, z = t.shape
    sn, sc, sx, sy, sz = t.stride()
    mult, vs = 1, None
    for d, s in ((z, sz), (y, sy), (x, sx)):
        if d > 1:
            if vs is None:
                if s % mult:
                    raise WindSRError(f"tensor voxels are not linearly strided: {tuple(t.shape)} / {t.stride()}")
                vs = s // mult
            elif s != vs * mult:
                raise WindSRError(f"tensor voxels are not linearly strided: {tuple(t.shape)} / {t.stride()}")
        mult *= d
    if vs is None:
        vs = 1
    return WsTensor(t.data_ptr(), _dtype_code(t), 0, sn, vs, sc)


def null_view() -> WsTensor:
    return WsTensor(None, 0, 0, 0, 0, 0)


def ptr(t):
    return None if t is None else t.data_ptr()
