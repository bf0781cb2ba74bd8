"""Host-side operators over the C-ABI: raw launches + the ``torch.autograd.Function``s the drop-in modules use.

Tensors are *logical* (N, C, X, Y, Z) like the reference's; the internal memory format is
``torch.channels_last_3d`` (N,X,Y,Z,C in memory), bf16 in BF16 mode and fp32 in FP32 mode.  Contiguous NCXYZ
boundary tensors are read/written in place through strided views — no layout-conversion passes.

Every function here launches hand-written CUDA from ``libwindsr.so`` on ``torch.cuda.current_stream()``;
nothing falls back to torch/cuDNN arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (MATH_BF16, MATH_FP32, MATH_TF32, PACK_SIMT_DGRAD, PACK_SIMT_FWD, PACK_TC_DGRAD,
                   PACK_TC_DGRAD_TF32, PACK_TC_FWD, PACK_TC_FWD_TF32, PATH_TCGEN05, WsConvShape, WsEpilogue, check, load,
                   null_view, ptr, stream_ptr, view)

# ------------------------------------------------------------------------------------------------------
# precision
# ------------------------------------------------------------------------------------------------------
_PRECISION = "bf16"
PRECISIONS = ("fp32", "tf32", "bf16")


def set_precision(mode: str) -> None:
    """'fp32' (CUDA-core FFMA, the 1e-5 parity mode), 'tf32' (tcgen05 kind::tf32 on fp32 activations — the arithmetic
    the reference actually ran with on its A100: torch's default ``cudnn.allow_tf32``) or 'bf16' (tcgen05 kind::f16,
    bf16 operands); fp32 accumulation everywhere."""
    global _PRECISION
    if mode not in PRECISIONS:
        raise ValueError(f"precision must be one of {PRECISIONS}, got {mode!r}")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


class precision:
    """Context manager: ``with ops.precision('fp32'): ...``"""

    def __init__(self, mode: str):
        self.mode = mode

    def __enter__(self):
        self.prev = get_precision()
        set_precision(self.mode)

    def __exit__(self, *a):
        set_precision(self.prev)


def math_mode() -> int:
    return {"bf16": MATH_BF16, "tf32": MATH_TF32, "fp32": MATH_FP32}[_PRECISION]


def _tc_pack_kind(dgrad: bool, math: int) -> int:
    if math == MATH_TF32:
        return PACK_TC_DGRAD_TF32 if dgrad else PACK_TC_FWD_TF32
    return PACK_TC_DGRAD if dgrad else PACK_TC_FWD


def act_dtype() -> torch.dtype:
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


# ------------------------------------------------------------------------------------------------------
# CUDA-graph capture support
# ------------------------------------------------------------------------------------------------------
# While torch.cuda.graph() records a training step (GAN_models/graph_step.py) every device buffer a captured kernel
# touches must stay valid for the lifetime of the graph.  Caches that may later drop or replace their tensors
# (packed weights, stencil coefficients, auxiliary workspaces) are therefore bypassed during capture: the operand is
# rebuilt inside the capture — so a replay also re-derives it from the CURRENT weights — and a reference is parked
# in ``_KEEPALIVE`` until the capturing code collects it (``take_keepalive``).
_KEEPALIVE = []


def capturing() -> bool:
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


def keepalive(*tensors):
    _KEEPALIVE.extend(t for t in tensors if t is not None)


def take_keepalive():
    out = list(_KEEPALIVE)
    _KEEPALIVE.clear()
    return out


# ------------------------------------------------------------------------------------------------------
# allocation helpers
# ------------------------------------------------------------------------------------------------------
def empty_cl(n, c, x, y, z, dtype, device) -> torch.Tensor:
    """Logical (n,c,x,y,z) tensor in channels_last_3d memory."""
    return torch.empty((n, x, y, z, c), dtype=dtype, device=device).permute(0, 4, 1, 2, 3)


def zeros_cl(n, c, x, y, z, dtype, device) -> torch.Tensor:
    return torch.zeros((n, x, y, z, c), dtype=dtype, device=device).permute(0, 4, 1, 2, 3)


def _require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.WindSRError("windsr ops need CUDA tensors: the hot path has no CPU fallback")


_triple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def make_shape(x_shape, cout, kernel, stride, padding) -> WsConvShape:
    n, cin, X, Y, Z = x_shape
    kx, ky, kz = _triple(kernel)
    sx, sy, sz = _triple(stride)
    px, py, pz = _triple(padding)
    return WsConvShape(n, X, Y, Z, cin, cout, kx, ky, kz, sx, sy, sz, px, py, pz)


def out_dims(s: WsConvShape) -> Tuple[int, int, int]:
    return ((s.x + 2 * s.px - s.kx) // s.sx + 1, (s.y + 2 * s.py - s.ky) // s.sy + 1,
            (s.z + 2 * s.pz - s.kz) // s.sz + 1)


# ------------------------------------------------------------------------------------------------------
# packed-weight cache (fp32 torch-layout Parameter -> kernel operand layout), keyed by parameter version
# ------------------------------------------------------------------------------------------------------
from torch.optim.optimizer import register_optimizer_step_post_hook as _register_step_hook  # noqa: E402

_WEIGHTS_EPOCH = 0


def invalidate_packed_weights(*_args, **_kwargs) -> None:
    """Force every packed-weight cache to repack on next use.  Called after EVERY optimizer step through torch's
    global step hook: single-kernel ("fused") optimizers update parameters without bumping ``Tensor._version``,
    so the version stamp alone would leave stale operand copies behind.  Call it by hand after writing weights
    through a path torch does not see (raw pointers, ``.data`` aliases)."""
    global _WEIGHTS_EPOCH
    _WEIGHTS_EPOCH += 1


_register_step_hook(invalidate_packed_weights)


class PackedWeights:
    """Caches the packed copies of one weight Parameter; repacks when the parameter changes
    (load_state_dict / in-place ops bump ``_version``; ``.to()`` changes ``data_ptr``; any optimizer step bumps
    the module-wide epoch, see ``invalidate_packed_weights``)."""

    def __init__(self):
        self._cache = {}

    def get(self, w: torch.Tensor, shape: WsConvShape, kind: int, pad_cout: int = 0) -> torch.Tensor:
        """``pad_cout`` > w.shape[0]: pack as if the layer had that many output channels (extra filters zero) —
        lets the 3-channel hr_convs.2 run its dgrad / wgrad on the tensor-core kernels (UMMA needs K, N >= 16)."""
        key = (kind, pad_cout)
        stamp = (w._version, w.data_ptr(), shape.cin, shape.cout, _WEIGHTS_EPOCH)
        hit = self._cache.get(key)
        cap = capturing()
        if hit is not None and hit[0] == stamp and not cap:
            return hit[1]
        src = w
        if pad_cout > w.shape[0]:
            src = torch.cat((w.detach(), w.new_zeros((pad_cout - w.shape[0],) + tuple(w.shape[1:]))), 0)
        packed = pack_weights(src, shape, kind)
        if cap:
            keepalive(packed)  # the captured pack kernel refreshes it on every replay; never served from the cache
        else:
            self._cache[key] = (stamp, packed)
        return packed

    def clear(self):
        self._cache.clear()


def pack_weights(w: torch.Tensor, shape: WsConvShape, kind: int) -> torch.Tensor:
    _require_cuda(w)
    lib = load()
    wd = w.detach()
    if wd.dtype != torch.float32 or not wd.is_contiguous():
        wd = wd.float().contiguous()
    nbytes = lib.ws_packed_weight_bytes(C.byref(shape), kind)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    check(lib.ws_pack_weights(wd.data_ptr(), C.byref(shape), kind, packed.data_ptr(), stream_ptr()),
          "ws_pack_weights")
    return packed


# ------------------------------------------------------------------------------------------------------
# raw launches
# ------------------------------------------------------------------------------------------------------
def _epilogue(bias=None, oscale=None, chan_scale=None, slope=1.0, alpha=1.0, res1=None, beta1=0.0, res2=None,
              beta2=0.0, mask=None, mask_c0=0, mask_c1=0, mask_slope=1.0, out2=None, stat_sum=None,
              stat_sqsum=None) -> WsEpilogue:
    return WsEpilogue(ptr(bias), ptr(oscale), ptr(chan_scale), slope, alpha, beta1, beta2,
                      view(res1) if res1 is not None else null_view(),
                      view(res2) if res2 is not None else null_view(),
                      view(mask) if mask is not None else null_view(), mask_c0, mask_c1, mask_slope,
                      # TF32 mode: LeakyReLU outputs are pure activations (conv operands only) -> store them rounded
                      1 if (_PRECISION == "tf32" and slope != 1.0) else 0,
                      view(out2) if out2 is not None else null_view(), ptr(stat_sum), ptr(stat_sqsum))


def fwd_path(shape: WsConvShape, x: torch.Tensor, math: Optional[int] = None) -> int:
    xv = view(x)
    return load().ws_conv3d_fwd_path(C.byref(shape), C.byref(xv), None, math_mode() if math is None else math)


def dgrad_path(shape: WsConvShape, dy: torch.Tensor, math: Optional[int] = None) -> int:
    dv = view(dy)
    return load().ws_conv3d_dgrad_path(C.byref(shape), C.byref(dv), None, math_mode() if math is None else math)


def wgrad_path(shape: WsConvShape, x: torch.Tensor, dy: torch.Tensor, math: Optional[int] = None) -> int:
    xv, dv = view(x), view(dy)
    return load().ws_conv3d_wgrad_path(C.byref(shape), C.byref(xv), C.byref(dv),
                                       math_mode() if math is None else math)


# Optional in-stream timing of selected launches (bench.py's roofline leg): CUDA events recorded on the
# launching stream right around the kernel, nothing else changes.
_KERNEL_TIMER = None


def set_kernel_timer(predicate) -> None:
    """predicate(kind, shape) -> bool selects launches ('fwd' | 'dgrad' | 'wgrad'); None switches it off."""
    global _KERNEL_TIMER
    _KERNEL_TIMER = (predicate, []) if predicate is not None else None


def kernel_timer_events():
    return _KERNEL_TIMER[1] if _KERNEL_TIMER is not None else []


class _timed:
    def __init__(self, kind, shape):
        self.on = _KERNEL_TIMER is not None and _KERNEL_TIMER[0](kind, shape)

    def __enter__(self):
        if self.on:
            self.t0 = torch.cuda.Event(enable_timing=True)
            self.t1 = torch.cuda.Event(enable_timing=True)
            self.t0.record()

    def __exit__(self, *a):
        if self.on:
            self.t1.record()
            _KERNEL_TIMER[1].append((self.t0, self.t1))


_GRAPH_LAUNCHES = [0]


def count_graph_launches(n: int) -> None:
    """A CUDA-graph replay executed ``n`` libwindsr kernels (counted when the graph was captured)."""
    _GRAPH_LAUNCHES[0] += int(n)


def launch_count() -> int:
    """CUDA kernels of libwindsr.so executed by this process so far: direct launches (counted inside the library,
    which also sees the launches recorded during a graph capture) plus the kernels of every graph replay."""
    return int(load().ws_launch_count()) + _GRAPH_LAUNCHES[0]


def conv_fwd(x: torch.Tensor, w: torch.Tensor, cache: Optional[PackedWeights], shape: WsConvShape,
             out: torch.Tensor, math: Optional[int] = None, **ep) -> torch.Tensor:
    """out = epilogue(conv3d(x, w)); `w` is the fp32 torch-layout weight, packed (and cached) here."""
    _require_cuda(x)
    lib = load()
    math = math_mode() if math is None else math
    xv, ov = view(x), view(out)
    path = lib.ws_conv3d_fwd_path(C.byref(shape), C.byref(xv), C.byref(ov), math)
    kind = _tc_pack_kind(False, math) if path == PATH_TCGEN05 else PACK_SIMT_FWD
    packed = cache.get(w, shape, kind) if cache is not None else pack_weights(w, shape, kind)
    e = _epilogue(**ep)
    with _timed("fwd", shape):
        check(lib.ws_conv3d_fwd(C.byref(shape), C.byref(xv), packed.data_ptr(), C.byref(ov), C.byref(e), math,
                                stream_ptr()), "ws_conv3d_fwd")
    return out


def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, cache: Optional[PackedWeights], shape: WsConvShape,
               dx: torch.Tensor, math: Optional[int] = None, pad_cout: int = 0, **ep) -> torch.Tensor:
    """``shape.cout`` must already be the padded count when ``pad_cout`` is used (dy has that many channels)."""
    _require_cuda(dy)
    lib = load()
    math = math_mode() if math is None else math
    dv, xv = view(dy), view(dx)
    path = lib.ws_conv3d_dgrad_path(C.byref(shape), C.byref(dv), C.byref(xv), math)
    kind = _tc_pack_kind(True, math) if path == PATH_TCGEN05 else PACK_SIMT_DGRAD
    if cache is None:
        cache = PackedWeights()
    packed = cache.get(w, shape, kind, pad_cout=pad_cout)
    e = _epilogue(**ep)
    with _timed("dgrad", shape):
        check(lib.ws_conv3d_dgrad(C.byref(shape), C.byref(dv), packed.data_ptr(), C.byref(xv), C.byref(e), math,
                                  stream_ptr()), "ws_conv3d_dgrad")
    return dx


_WORKSPACES = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand (stream-ordered reuse is safe on one stream)."""
    key = (device, stream_ptr())
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    if capturing():
        keepalive(buf)
    return buf


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, shape: WsConvShape, want_bias: bool = False,
               math: Optional[int] = None, want_weight: bool = True):
    """Returns (dw in torch layout fp32 or None, db or None)."""
    _require_cuda(x)
    lib = load()
    math = math_mode() if math is None else math
    dw = torch.empty((shape.cout, shape.cin, shape.kx, shape.ky, shape.kz), dtype=torch.float32,
                     device=x.device) if want_weight else None
    db = torch.empty((shape.cout,), dtype=torch.float32, device=x.device) if want_bias else None
    nbytes = lib.ws_conv3d_wgrad_workspace_bytes(C.byref(shape), math)
    wsp = _workspace(nbytes, x.device)
    xv, dv = view(x), view(dy)
    with _timed("wgrad", shape):
        check(lib.ws_conv3d_wgrad(C.byref(shape), C.byref(xv), C.byref(dv), ptr(dw), ptr(db), 0, math,
                                  wsp.data_ptr(), wsp.numel(), stream_ptr()), "ws_conv3d_wgrad")
    return dw, db


def copy_(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[...] = src (dtype / layout converting strided copy)."""
    n, c, x, y, z = src.shape
    sv, dv = view(src), view(dst)
    check(load().ws_copy(C.byref(sv), C.byref(dv), n, c, x * y * z, stream_ptr()), "ws_copy")
    return dst


def axpby(x1: torch.Tensor, a: float, x2: Optional[torch.Tensor], b: float, out: torch.Tensor) -> torch.Tensor:
    n, c, x, y, z = x1.shape
    v1, v2, vo = view(x1), (view(x2) if x2 is not None else null_view()), view(out)
    check(load().ws_axpby(C.byref(v1), a, C.byref(v2), b, C.byref(vo), n, c, x * y * z, stream_ptr()), "ws_axpby")
    return out


def lrelu_bwd(dy: torch.Tensor, y: torch.Tensor, slope: float, out: torch.Tensor, chan_scale=None,
              oscale=None) -> torch.Tensor:
    n, c, x, yy, z = dy.shape
    dv, yv, ov = view(dy), view(y), view(out)
    check(load().ws_lrelu_bwd(C.byref(dv), C.byref(yv), slope, ptr(chan_scale), ptr(oscale), C.byref(ov), n, c,
                              x * yy * z, stream_ptr()), "ws_lrelu_bwd")
    return out


def upsample_fwd(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    n, c, X, Y, Z = x.shape
    xv, ov = view(x), view(out)
    check(load().ws_upsample_nearest_xy_fwd(C.byref(xv), C.byref(ov), n, c, X, Y, Z, stream_ptr()),
          "ws_upsample_nearest_xy_fwd")
    return out


def upsample_bwd(dout: torch.Tensor, din: torch.Tensor) -> torch.Tensor:
    n, c, X, Y, Z = din.shape
    dv, iv = view(dout), view(din)
    check(load().ws_upsample_nearest_xy_bwd(C.byref(dv), C.byref(iv), n, c, X, Y, Z, stream_ptr()),
          "ws_upsample_nearest_xy_bwd")
    return din


# ------------------------------------------------------------------------------------------------------
# autograd: generic fused conv block
# ------------------------------------------------------------------------------------------------------
def _as_act(x: torch.Tensor) -> torch.Tensor:
    """Inputs may be fp32 boundary tensors or internal activations; both are consumed in place."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return x


def _im2col_eligible(shape: WsConvShape, x: torch.Tensor) -> bool:
    """Narrow-input conv that should run as im2col + 1x1x1 tensor-core conv: bf16 mode, 2 <= Cin <= 4 (Cin = 1 is the
    terrain conv, whose input is raw altitude in metres and stays on the fp32 CUDA-core path), more than one tap, an
    output width the UMMA N dimension takes, and a volume large enough for the extra pass to pay (D's features.0:
    3 -> 32 at 128x128x10; the generator's 4 -> 128 feature conv on the 16x16x10 LR volume stays direct)."""
    taps = shape.kx * shape.ky * shape.kz
    xo, yo, zo = out_dims(shape)
    return (get_precision() == "bf16" and 2 <= shape.cin <= 4 and taps > 1 and shape.cout % 16 == 0
            and shape.cout <= 256 and shape.n * xo * yo * zo >= 65536 and os.environ.get("WINDSR_IM2COL", "1") != "0")


def _im2col(x: torch.Tensor, shape: WsConvShape):
    """(U, 1x1x1 shape): U[n, v, tap*cin + ci] = x[n, ci, v (+) tap] in channels-last bf16 (ws_im2col)."""
    taps = shape.kx * shape.ky * shape.kz
    cpad = (taps * shape.cin + 15) // 16 * 16
    xo, yo, zo = out_dims(shape)
    u = empty_cl(shape.n, cpad, xo, yo, zo, torch.bfloat16, x.device)
    xv, uv = view(x), view(u)
    check(load().ws_im2col(C.byref(shape), C.byref(xv), C.byref(uv), cpad, stream_ptr()), "ws_im2col")
    return u, make_shape(u.shape, shape.cout, 1, 1, 0)


def _im2col_weight(weight: torch.Tensor, cpad: int) -> torch.Tensor:
    """w'[co][tap*cin + ci] = w[co][ci][tap], zero pad columns; as a (cout, cpad, 1, 1, 1) conv weight."""
    co, ci = weight.shape[:2]
    w2 = weight.detach().reshape(co, ci, -1).permute(0, 2, 1).reshape(co, -1)
    if w2.shape[1] < cpad:
        w2 = torch.cat((w2, w2.new_zeros((co, cpad - w2.shape[1]))), 1)
    return w2.reshape(co, cpad, 1, 1, 1).contiguous()


# ---- deferred ("late") weight gradients ---------------------------------------------------------------------------
_LATE_WGRAD = []
_LATE_WGRAD_QUEUED = [False]
_LATE_WGRAD_MIN_FLOPS = float(os.environ.get("WINDSR_LATE_WGRAD_MIN_TFLOP", "3")) * 1e12


def late_wgrad_enabled() -> bool:
    """WINDSR_LATE_WGRAD=1 (opt-in).  Measured on 2 GPUs it LOSES 1.2 ms (37.7 -> 38.9 ms per step): the weight-gradient
    kernel is one wave of 147 equally loaded CTAs, and while NCCL's all-reduce kernels hold a few SMs part of that wave
    has to run as a second one — hiding the all-reduce behind it costs more than leaving it exposed."""
    return os.environ.get("WINDSR_LATE_WGRAD", "0") == "1"


def capturing_disallows_late() -> bool:
    return False  # (the final callback runs inside the captured backward pass: nothing to exclude)


def _defer_wgrad(weight, bias, compute) -> None:
    _LATE_WGRAD.append((weight, bias, compute))
    if _LATE_WGRAD_QUEUED[0]:
        return
    _LATE_WGRAD_QUEUED[0] = True

    def _run():
        _LATE_WGRAD_QUEUED[0] = False
        items = list(_LATE_WGRAD)
        _LATE_WGRAD.clear()
        for w, b, fn in items:
            with torch.no_grad():
                dw, db = fn()
            for p, g in ((w, dw), (b, db)):
                if p is None or g is None:
                    continue
                p.grad = g if p.grad is None else p.grad + g
                for hook in (getattr(p, "_post_accumulate_grad_hooks", None) or {}).values():
                    hook(p)

    torch.autograd.Variable._execution_engine.queue_callback(_run)


class ConvFn(torch.autograd.Function):
    """y = lrelu(conv(x, w) * oscale + bias) * chan_scale + beta * res

    Covers create_conv_lrelu_layer (torch_blocks.py:5-37) without / with eval-mode BatchNorm, the trunk skip
    (torch_blocks.py:46), Dropout3d folded into hr_convs.0 (Generator…py:95-104) and the biased hr_convs.2.
    """

    @staticmethod
    def forward(ctx, x, weight, bias, res, cfg):
        # cfg: dict(stride, padding, slope, oscale, shift, chan_scale, beta, out, out_dtype, out_contig, cache)
        x = _as_act(x)
        shape = make_shape(x.shape, weight.shape[0], weight.shape[2:], cfg["stride"], cfg["padding"])
        xo, yo, zo = out_dims(shape)
        out = cfg.get("out")
        if out is None:
            dt = cfg.get("out_dtype") or act_dtype()
            if cfg.get("out_contig"):
                out = torch.empty((shape.n, shape.cout, xo, yo, zo), dtype=dt, device=x.device)
            else:
                out = empty_cl(shape.n, shape.cout, xo, yo, zo, dt, x.device)
        oscale, shift = cfg.get("oscale"), cfg.get("shift")
        b = bias if bias is not None else shift
        ctx.im2col = _im2col_eligible(shape, x)
        if ctx.im2col:
            u, shape1 = _im2col(x, shape)
            conv_fwd(u, _im2col_weight(weight, u.shape[1]), None, shape1, out,
                     bias=b.detach() if b is not None else None, oscale=oscale, chan_scale=cfg.get("chan_scale"),
                     slope=cfg.get("slope", 1.0), res1=res, beta1=cfg.get("beta", 1.0) if res is not None else 0.0)
        else:
            conv_fwd(x, weight, cfg.get("cache"), shape, out, bias=b.detach() if b is not None else None,
                     oscale=oscale, chan_scale=cfg.get("chan_scale"), slope=cfg.get("slope", 1.0),
                     res1=res, beta1=cfg.get("beta", 1.0) if res is not None else 0.0)
        ctx.shape = shape
        # NOT the caller's `out` tensor: it is this node's own output, and node -> ctx.cfg -> out -> grad_fn -> node is
        # a reference cycle through C++ that Python's GC cannot see — the node, everything upstream of it and every
        # parameter's AccumulateGrad node would stay alive forever (which, besides the leak, pins those nodes to the
        # stream of their first iteration and invalidates a later CUDA-graph capture)
        ctx.cfg = {k: v for k, v in cfg.items() if k != "out"}
        ctx.has_bias = bias is not None
        ctx.bias_param = bias
        ctx.has_res = res is not None
        ctx.x_dtype = x.dtype
        needs_y = cfg.get("slope", 1.0) != 1.0
        if needs_y and res is not None:
            raise _lib.WindSRError("ConvFn: activation + residual in one epilogue is not differentiable here")
        ctx.save_for_backward(x, weight, out if needs_y else None)
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        cfg, shape = ctx.cfg, ctx.shape
        slope = cfg.get("slope", 1.0)
        chan_scale, oscale = cfg.get("chan_scale"), cfg.get("oscale")
        need_x, need_w, need_b, need_r = ctx.needs_input_grad[:4]
        dres = None
        if ctx.has_res and need_r:
            beta = cfg.get("beta", 1.0)
            dres = dy if beta == 1.0 else dy * beta
        # g = dL/d(conv accumulator) in the compute dtype
        cdt = act_dtype()
        if slope != 1.0 or chan_scale is not None or oscale is not None or dy.dtype != cdt or \
                not _linear_voxels(dy):
            g = empty_cl(*dy.shape, cdt, dy.device)
            src_y = y if y is not None else dy
            lrelu_bwd(dy, src_y, slope if y is not None else 1.0, g, chan_scale=chan_scale, oscale=oscale)
        else:
            g = dy
        dx = dw = db = None
        pad_cout = 0
        if (get_precision() in ("bf16", "tf32") and shape.cout < 16 <= shape.cin and x.dtype == cdt
                and (shape.sx, shape.sy, shape.sz) == (1, 1, 1)):
            # narrow output (hr_convs.2: 144 -> 3): zero-pad the gradient to 16 channels so dgrad and wgrad run on
            # the tensor-core kernels instead of the CUDA-core family
            pad_cout = 16
            gp = zeros_cl(g.shape[0], pad_cout, *g.shape[2:], cdt, g.device)
            copy_(g, gp[:, :shape.cout])
            g = gp
            shape = make_shape(x.shape, pad_cout, (shape.kx, shape.ky, shape.kz), 1, (shape.px, shape.py, shape.pz))
        has_bias, im2col, w_shape = ctx.has_bias, ctx.im2col, weight.shape

        def compute_w():
            dw = db = None
            rem = shape.cout % 128
            # (a narrow layer as a whole — terrain_convs.1, 16 -> 16 — is all "remainder": its M = 16 GEMM would use an
            # eighth of every MMA)
            narrow = 8 <= shape.cout <= 32 and shape.cin >= 16 and os.environ.get("WINDSR_NARROW_XFOLD", "1") != "0"
            if (need_w and not pad_cout and cdt == torch.bfloat16 and ((shape.cout > 128 and 0 < rem <= 32) or narrow)
                    and rem * shape.kx <= 128 and (rem * shape.kx) % 8 == 0 and shape.kx > 1
                    and (shape.sx, shape.sy, shape.sz) == (1, 1, 1) and x.dtype == torch.bfloat16
                    and g.dtype == torch.bfloat16):
                # Cout = 128*q + rem (hr_convs.0: 144 = 128 + 16): a second 128-row M tile for `rem` rows would cost as
                # much as the first.  Instead fold the kx taps of the remainder channels into the channel dimension
                # (U[x', (dx,co)] = g[x' - dx + px, co], windsr.h "x-fold helpers"): one (1,ky,kz) wgrad with kx*rem
                # rows — kx times fewer MMAs for the remainder.
                main = shape.cout - rem
                dw_main = None
                if main > 0:
                    s_main = make_shape(x.shape, main, (shape.kx, shape.ky, shape.kz), 1, (shape.px, shape.py, shape.pz))
                    dw_main, _ = conv_wgrad(x, g[:, :main], s_main)
                cu = rem * shape.kx
                u = empty_cl(g.shape[0], cu, *g.shape[2:], cdt, g.device)
                gv, uv = view(g[:, main:]), view(u)
                check(load().ws_xunfold(C.byref(gv), C.byref(uv), g.shape[0], rem, shape.kx, shape.px, cu, g.shape[2],
                                        g.shape[3], g.shape[4], stream_ptr()), "ws_xunfold")
                s_rem = make_shape(x.shape, cu, (1, shape.ky, shape.kz), 1, (0, shape.py, shape.pz))
                dw_u, _ = conv_wgrad(x, u, s_rem)
                dw_rem = dw_u.reshape(shape.kx, rem, shape.cin, shape.ky, shape.kz).permute(1, 2, 0, 3, 4)
                dw = torch.cat((dw_main, dw_rem), 0) if dw_main is not None else dw_rem.contiguous()
                if need_b and has_bias:
                    _, db = conv_wgrad(x, g, shape, want_bias=True, want_weight=False)  # one pass over g (bias kernel)
            elif need_w and im2col and g.dtype == torch.bfloat16 and not pad_cout:
                # weight gradient of the im2col form: a 1x1x1 wgrad with K = taps*cin on the tensor cores
                u, shape1 = _im2col(x, shape)
                dw2, db = conv_wgrad(u, g, shape1, want_bias=has_bias and need_b)
                taps = shape.kx * shape.ky * shape.kz
                dw = dw2.reshape(shape.cout, -1)[:, :taps * shape.cin].reshape(shape.cout, taps, shape.cin) \
                    .permute(0, 2, 1).reshape(w_shape).contiguous()
            elif need_w or (need_b and has_bias):
                dw, db = conv_wgrad(x, g, shape, want_bias=has_bias and need_b, want_weight=need_w)
                if pad_cout:
                    dw = dw[:w_shape[0]].contiguous() if dw is not None else None
                    db = db[:w_shape[0]].contiguous() if db is not None else None
            return dw, db

        flops = 2.0 * shape.n * shape.cin * shape.cout * shape.kx * shape.ky * shape.kz * dy.shape[2] * dy.shape[3] * dy.shape[4]
        if need_w and late_wgrad_enabled() and flops >= _LATE_WGRAD_MIN_FLOPS and not capturing_disallows_late():
            # the largest weight gradients go LAST (final autograd callback): under data parallelism the all-reduce of
            # everything produced before them — 132 MB of trunk gradients arrive at the very end of backward — then has
            # several milliseconds of compute to hide behind instead of being exposed
            bias_p = ctx.bias_param if need_b and has_bias else None
            _defer_wgrad(weight, bias_p, compute_w)
        else:
            dw, db = compute_w()
        if need_x:
            if cfg.get("dx_contig"):
                dx = torch.empty(x.shape, dtype=torch.float32, device=x.device)
            else:
                dx = empty_cl(*x.shape, torch.float32 if cfg.get("dx_f32") else cdt, x.device)
            conv_dgrad(g, weight, cfg.get("cache"), shape, dx, pad_cout=pad_cout)
        return dx, dw, db, dres, None


class XFoldConvFn(torch.autograd.Function):
    """Stride-1 conv with a very narrow output (kx * cout <= 16: hr_convs.2, 144 -> 3, Generator…py:105-110) with
    the kx taps along x folded into the output-channel dimension (windsr.h "x-fold helpers"): one (1,ky,kz) conv
    with kx*cout (-> 16) channels on the tensor cores + a shifted sum; 5x fewer MMAs than the direct form, whose
    N = 16 MMAs cost as much as N = 144 ones.  Output: contiguous NCXYZ fp32 (the generator's result)."""

    @staticmethod
    def _folded_weight(weight):
        co, ci, kx, ky, kz = weight.shape
        w5 = weight.detach().permute(2, 0, 1, 3, 4).reshape(kx * co, ci, 1, ky, kz)
        if kx * co < 16:
            w5 = torch.cat((w5, w5.new_zeros((16 - kx * co, ci, 1, ky, kz))), 0)
        return w5.contiguous()

    @staticmethod
    def forward(ctx, x, weight, bias, padding):
        lib = load()
        co, ci, kx, ky, kz = weight.shape
        px, py, pz = padding
        n, _, X, Y, Z = x.shape
        w5 = XFoldConvFn._folded_weight(weight)
        shape = make_shape(x.shape, 16, (1, ky, kz), 1, (0, py, pz))
        ybuf = empty_cl(n, 16, X, Y, Z, torch.float32, x.device)
        conv_fwd(x, w5, None, shape, ybuf)
        out = torch.empty((n, co, X, Y, Z), dtype=torch.float32, device=x.device)
        yv, ov = view(ybuf), view(out)
        check(lib.ws_xfold_sum(C.byref(yv), ptr(bias.detach() if bias is not None else None), C.byref(ov), n, co, kx,
                               px, X, Y, Z, stream_ptr()), "ws_xfold_sum")
        ctx.save_for_backward(x, w5)
        ctx.meta = (tuple(weight.shape), shape, px, bias is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = load()
        x, w5 = ctx.saved_tensors
        (co, ci, kx, ky, kz), shape, px, has_bias = ctx.meta
        n, _, X, Y, Z = x.shape
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        dout = dout.contiguous() if not _linear_voxels(dout) else dout
        u = empty_cl(n, 16, X, Y, Z, act_dtype(), x.device)
        dv, uv = view(dout), view(u)
        check(lib.ws_xunfold(C.byref(dv), C.byref(uv), n, co, kx, px, 16, X, Y, Z, stream_ptr()), "ws_xunfold")
        dx = dw = db = None
        if need_w:
            dw5, _ = conv_wgrad(x, u, shape)
            dw = dw5[:kx * co].reshape(kx, co, ci, ky, kz).permute(1, 2, 0, 3, 4).contiguous()
        if need_b and has_bias:
            db = dout.sum((0, 2, 3, 4))
        if need_x:
            dx = empty_cl(*x.shape, act_dtype(), x.device)
            conv_dgrad(u, w5, None, shape, dx)
        return dx, dw, db, None


class XYFoldConvFn(torch.autograd.Function):
    """The same idea with BOTH lateral tap axes folded (windsr.h "xy-fold"): a (1,1,kz) conv with kx*ky*cout (75 -> 80)
    output channels + a 2-D shifted sum.  hr_convs.2 (5x5x5, 144 -> 3) then issues 25x fewer MMAs than its direct form
    (5x fewer than the x-fold) in forward, data-gradient (K = 80) and weight-gradient (N = 144, 5 taps)."""

    @staticmethod
    def _folded_weight(weight, cpad):
        co, ci, kx, ky, kz = weight.shape
        w5 = weight.detach().permute(2, 3, 0, 1, 4).reshape(kx * ky * co, ci, 1, 1, kz)
        if kx * ky * co < cpad:
            w5 = torch.cat((w5, w5.new_zeros((cpad - kx * ky * co, ci, 1, 1, kz))), 0)
        return w5.contiguous()

    @staticmethod
    def forward(ctx, x, weight, bias, padding):
        lib = load()
        co, ci, kx, ky, kz = weight.shape
        px, py, pz = padding
        n, _, X, Y, Z = x.shape
        cpad = (kx * ky * co + 15) // 16 * 16
        w5 = XYFoldConvFn._folded_weight(weight, cpad)
        shape = make_shape(x.shape, cpad, (1, 1, kz), 1, (0, 0, pz))
        ybuf = empty_cl(n, cpad, X, Y, Z, torch.float32, x.device)
        conv_fwd(x, w5, None, shape, ybuf)
        out = torch.empty((n, co, X, Y, Z), dtype=torch.float32, device=x.device)
        yv, ov = view(ybuf), view(out)
        check(lib.ws_xyfold_sum(C.byref(yv), ptr(bias.detach() if bias is not None else None), C.byref(ov), n, co, kx,
                                ky, px, py, X, Y, Z, stream_ptr()), "ws_xyfold_sum")
        ctx.save_for_backward(x, w5)
        ctx.meta = (tuple(weight.shape), shape, px, py, cpad, bias is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = load()
        x, w5 = ctx.saved_tensors
        (co, ci, kx, ky, kz), shape, px, py, cpad, has_bias = ctx.meta
        n, _, X, Y, Z = x.shape
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        dout = dout.contiguous() if not _linear_voxels(dout) else dout
        u = empty_cl(n, cpad, X, Y, Z, act_dtype(), x.device)
        dv, uv = view(dout), view(u)
        check(lib.ws_xyunfold(C.byref(dv), C.byref(uv), n, co, kx, ky, px, py, cpad, X, Y, Z, stream_ptr()),
              "ws_xyunfold")
        dx = dw = db = None
        if need_w:
            dw5, _ = conv_wgrad(x, u, shape)
            dw = dw5[:kx * ky * co].reshape(kx, ky, co, ci, kz).permute(2, 3, 0, 1, 4).contiguous()
        if need_b and has_bias:
            db = dout.sum((0, 2, 3, 4))
        if need_x:
            dx = empty_cl(*x.shape, act_dtype(), x.device)
            conv_dgrad(u, w5, None, shape, dx)
        return dx, dw, db, None


def _linear_voxels(t: torch.Tensor) -> bool:
    try:
        view(t)
        return True
    except _lib.WindSRError:
        return False


def conv_block(x, weight, bias=None, res=None, *, stride=1, padding=0, slope=1.0, oscale=None, shift=None,
               chan_scale=None, beta=1.0, out=None, out_dtype=None, out_contig=False, cache=None,
               dx_f32=False, dx_contig=False):
    cfg = dict(stride=stride, padding=padding, slope=slope, oscale=oscale, shift=shift, chan_scale=chan_scale,
               beta=beta, out=out, out_dtype=out_dtype, out_contig=out_contig, cache=cache, dx_f32=dx_f32,
               dx_contig=dx_contig)
    return ConvFn.apply(x, weight, bias, res, cfg)


# ------------------------------------------------------------------------------------------------------
# autograd: residual dense block (torch_blocks.py:217-290) as ONE node
# ------------------------------------------------------------------------------------------------------
class RDBState:
    """Per-RDB device scratch that outlives a step: the packed copies of its weights (forward and dgrad
    operand layouts) and the parameter versions they were packed from."""

    def __init__(self):
        self.packed = {0: None, 1: None}
        self.stamp = {0: None, 1: None}

    def buffers(self, desc, params, dgrad: int):
        lib = load()
        nconv = desc.nconv
        if self.packed[dgrad] is None or self.packed[dgrad][0].device != params[0].device:
            self.packed[dgrad] = [torch.empty(lib.ws_rdb_packed_bytes(C.byref(desc), i, dgrad), dtype=torch.uint8,
                                              device=params[0].device) for i in range(nconv + 1)]
            self.stamp[dgrad] = None
        # (the geometry decides which packing the executor uses: persistent z-fold, x-fold or direct)
        stamp = tuple((p._version, p.data_ptr()) for p in params[:nconv + 1]) + \
            (desc.math, _WEIGHTS_EPOCH, desc.n, desc.x, desc.y, desc.z)
        repack = stamp != self.stamp[dgrad]
        self.stamp[dgrad] = stamp
        if capturing():  # a replay must repack from the weights of ITS step; eager calls after it must not trust us
            repack = True
            self.stamp[dgrad] = None
        return self.packed[dgrad], int(repack)


# ---- auxiliary stream for the off-critical-path weight gradients of the residual dense blocks -------------------
_AUX = {}           # device -> (torch.cuda.Stream, workspace tensor)
_AUX_CAPTURE = {}   # device -> workspace tensor allocated inside the current graph capture
_AUX_JOIN_QUEUED = [False]


def _aux_enabled() -> bool:
    return os.environ.get("WINDSR_AUX_STREAM", "1") != "0"


def _aux_stream(device, nbytes: int):
    hit = _AUX.get(device)
    if capturing():
        # a dedicated workspace inside the graph's memory pool: the shared one may be re-grown (freed) by later eager
        # calls while the graph still points at it
        stream = hit[0] if hit is not None else torch.cuda.Stream(device=device)
        if hit is None:
            _AUX[device] = (stream, torch.empty(1 << 20, dtype=torch.uint8, device=device))
        cap = _AUX_CAPTURE.get(device)
        if cap is None or cap.numel() < nbytes:
            cap = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            _AUX_CAPTURE[device] = cap
        keepalive(cap)
        return stream, cap
    if hit is None or hit[1].numel() < nbytes:
        stream = hit[0] if hit is not None else torch.cuda.Stream(device=device)
        if hit is not None:
            hit[1].record_stream(stream)  # queued auxiliary-stream kernels may still be using the old workspace
        hit = (stream, torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device))
        _AUX[device] = hit
    return hit


def aux_join(device=None) -> None:
    """Make the current stream wait for the auxiliary stream(s): call before anything reads parameter gradients that
    an RDB backward produced (the autograd final callback below does it for a plain ``.backward()``; GradSync does it
    before it packs a bucket)."""
    for dev, (stream, _) in _AUX.items():
        if device is None or dev == device:
            if capturing():
                # a capturing stream may only wait for streams of the SAME capture: the auxiliary stream joins it when
                # ws_rdb_backward forks weight gradients onto it, which the small-geometry paths never do
                with torch.cuda.stream(stream):
                    forked = torch.cuda.is_current_stream_capturing()
                if not forked:
                    continue
            torch.cuda.current_stream(dev).wait_stream(stream)


def _queue_aux_join(device) -> None:
    if _AUX_JOIN_QUEUED[0]:
        return
    _AUX_JOIN_QUEUED[0] = True

    def _cb():
        _AUX_JOIN_QUEUED[0] = False
        with torch.cuda.device(device):
            aux_join(device)

    torch.autograd.Variable._execution_engine.queue_callback(_cb)


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class RDBFn(torch.autograd.Function):
    """x (fp32 trunk state, N,F,X,Y,Z) -> alpha * (LFF(dense(x)) + b) + beta1 * x + beta2 * outer

    ONE C-ABI call per direction (ws_rdb_forward / ws_rdb_backward): the dense convs write their LeakyReLU'd
    outputs straight into channel slices of one (F+4*gc)-channel concat buffer (no torch.cat,
    torch_blocks.py:214); the LFF epilogue applies bias, the 0.2 residual scale and the skip add(s)
    (torch_blocks.py:285-290, and :328-330 when this is the last RDB of an RRDB).  Backward walks the block in
    reverse: LFF wgrad/dgrad, then per dense conv LeakyReLU-backward of its slice, wgrad, dgrad accumulated into
    the fp32 gradient of the concat buffer.
    """

    @staticmethod
    def forward(ctx, x, outer, cfg, *params):
        # params: w0..w{k-1}, w_lff, b_lff
        _require_cuda(x)
        lib = load()
        nconv = cfg["nconv"]
        ws, w_lff, b_lff = params[:nconv], params[nconv], params[nconv + 1]
        n, f, X, Y, Z = x.shape
        gc = ws[0].shape[0] if nconv else 0
        ctot = f + nconv * gc
        cdt = act_dtype()
        k = ws[0].shape[2] if nconv else 1
        kl = w_lff.shape[2]
        desc = _lib.WsRdbDesc(n, X, Y, Z, f, gc, nconv, k, kl, cfg["slope"], cfg["alpha"], cfg["beta1"],
                              cfg["beta2"] if outer is not None else 0.0, math_mode(), 0)
        state: RDBState = cfg["state"]
        packed, desc.repack = state.buffers(desc, params, 0)
        _AUX_JOIN_QUEUED[0] = False  # a backward pass that died before its final callback must not mute the next one
        buf = empty_cl(n, ctot, X, Y, Z, cdt, x.device)
        out = empty_cl(n, f, X, Y, Z, torch.float32, x.device)
        xv, bv, ov = view(x), view(buf), view(out)
        outer_v = view(outer) if outer is not None else null_view()
        wd = [p.detach() for p in params[:nconv + 1]]
        wsp = _workspace(int(lib.ws_rdb_forward_workspace_bytes(C.byref(desc))), x.device)
        check(lib.ws_rdb_forward(C.byref(desc), C.byref(xv), C.byref(outer_v), C.byref(bv), C.byref(ov),
                                 _ptr_array(wd), _ptr_array(packed), ptr(b_lff), wsp.data_ptr(), wsp.numel(),
                                 stream_ptr()), "ws_rdb_forward")
        ctx.cfg = cfg
        ctx.desc_args = (n, X, Y, Z, f, gc, nconv, k, kl)
        ctx.has_outer = outer is not None
        ctx.save_for_backward(buf, *params)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = load()
        buf, *params = ctx.saved_tensors
        cfg = ctx.cfg
        n, X, Y, Z, f, gc, nconv, k, kl = ctx.desc_args
        ctot = f + nconv * gc
        cdt = act_dtype()
        dev = dy.device
        need = ctx.needs_input_grad
        need_params = any(need[3:])
        desc = _lib.WsRdbDesc(n, X, Y, Z, f, gc, nconv, k, kl, cfg["slope"], cfg["alpha"], cfg["beta1"],
                              cfg["beta2"] if ctx.has_outer else 0.0, math_mode(), 0)
        state: RDBState = cfg["state"]
        packed, desc.repack = state.buffers(desc, params, 1)
        if dy.dtype != torch.float32 or not _linear_voxels(dy):
            dy = dy.float().contiguous(memory_format=torch.channels_last_3d)
        dbuf = empty_cl(n, ctot, X, Y, Z, torch.float32, dev)
        g_lff = empty_cl(n, f, X, Y, Z, cdt, dev)
        g = empty_cl(n, max(nconv * gc, 1), X, Y, Z, cdt, dev)
        dx = empty_cl(n, f, X, Y, Z, torch.float32, dev) if need[0] else None
        grads = [None] * (nconv + 2)
        dw_arr = None
        db = None
        if need_params:
            for i in range(nconv + 1):
                grads[i] = torch.empty_like(params[i], dtype=torch.float32)
            if params[nconv + 1] is not None:
                db = torch.empty_like(params[nconv + 1], dtype=torch.float32)
                grads[nconv + 1] = db
            dw_arr = _ptr_array(grads[:nconv + 1])
        nbytes = int(lib.ws_rdb_backward_workspace_bytes(C.byref(desc)))
        wsp = _workspace(nbytes, dev)
        dyv, bv, dbv, glv, gv = view(dy), view(buf), view(dbuf), view(g_lff), view(g)
        dxv = view(dx) if dx is not None else null_view()
        wd = [p.detach() for p in params[:nconv + 1]]
        # weight gradients on the auxiliary stream (they are off the critical path of the backward chain)
        # (gradient accumulation: AccumulateGrad would add into an existing .grad on the main stream while the
        # auxiliary stream is still writing the new one — keep everything on one stream then)
        use_aux = (need_params and _aux_enabled() and get_precision() == "bf16"
                   and not any(p is not None and p.grad is not None for p in params))
        aux_s, aux_w = _aux_stream(dev, nbytes) if use_aux else (None, None)
        check(lib.ws_rdb_backward(C.byref(desc), C.byref(dyv), C.byref(bv), C.byref(dbv), C.byref(glv), C.byref(gv),
                                  C.byref(dxv), _ptr_array(wd), _ptr_array(packed), dw_arr, ptr(db),
                                  wsp.data_ptr(), wsp.numel(), stream_ptr(),
                                  aux_s.cuda_stream if use_aux else None, aux_w.data_ptr() if use_aux else None,
                                  aux_w.numel() if use_aux else 0), "ws_rdb_backward")
        if use_aux:
            # the operands of the weight gradients must outlive this function on the auxiliary stream, and whoever
            # reads the gradients must wait for it (autograd's final callback; GradSync joins before packing)
            for t in (buf, g, g_lff, *[t for t in grads if t is not None]):
                t.record_stream(aux_s)
            _queue_aux_join(dev)
        douter = None
        if ctx.has_outer and need[1]:
            douter = dy if cfg["beta2"] == 1.0 else dy * cfg["beta2"]
        return (dx, douter, None, *grads)


def trunk_batched_enabled() -> bool:
    return os.environ.get("WINDSR_TRUNK_BATCH", "1") != "0"


def trunk_supported(h: torch.Tensor, sig) -> bool:
    """Does ``ws_trunk_wgrad`` cover blocks of this geometry (sig: torch_blocks._trunk_signature) at h's size?"""
    f, gc, nconv, k = sig[0], sig[1], sig[2], sig[3]
    n, c, X, Y, Z = h.shape
    if c != f:
        return False
    desc = _lib.WsRdbDesc(n, X, Y, Z, f, gc, nconv, k, 1, 0.2, 1.0, 1.0, 0.0, math_mode(), 0)
    return bool(load().ws_trunk_wgrad_supported(C.byref(desc)))


class TrunkFn(torch.autograd.Function):
    """A run of identical residual dense blocks — the RRDB trunk (torch_blocks.py:293-330, 16 x 3 RDBs in the shipped
    generator) — as ONE autograd node.

    Forward is the chain of ``ws_rdb_forward`` calls of ``RDBFn``, with every block's concat buffer living in one slab
    (block index outermost).  Backward runs the blocks' data-gradient chains (``ws_rdb_backward`` with dw = NULL, g and
    g_lff into slabs as well) and then ``ws_trunk_wgrad``: the weight gradients of ALL blocks as one set of batched
    launches (a CTA owns (block, tap group) with the block's full voxel range as its K loop — no split-K, no atomics)
    instead of 48 x (merged GEMM + LFF GEMM + finalize + bias sums).  The parameter gradients are views of one flat
    fp32 tensor.

    cfg: nconv, slope, blocks = [dict(alpha, beta1, beta2, outer = index of the block whose INPUT is the outer skip
    of this block or None, state)].  params: per block w0..w{k-1}, w_lff, b_lff.
    """

    @staticmethod
    def forward(ctx, x, cfg, *params):
        _require_cuda(x)
        lib = load()
        nconv, blocks = cfg["nconv"], cfg["blocks"]
        R, per = len(blocks), nconv + 2
        n, f, X, Y, Z = x.shape
        gc = params[0].shape[0]
        k, kl = params[0].shape[2], params[nconv].shape[2]
        ctot = f + nconv * gc
        cdt = act_dtype()
        dev = x.device
        _AUX_JOIN_QUEUED[0] = False
        slab = torch.empty((R * n, X, Y, Z, ctot), dtype=cdt, device=dev).permute(0, 4, 1, 2, 3)
        inputs = {}
        want_in = {b["outer"] for b in blocks if b["outer"] is not None}
        h = x
        wsp = None
        # weights of ALL blocks packed by one launch (ws_trunk_repack_fwd) when any of them changed
        geo = _lib.WsRdbDesc(n, X, Y, Z, f, gc, nconv, k, kl, cfg["slope"], 1.0, 1.0, 0.0, math_mode(), 0)
        packs = [b["state"].buffers(geo, params[r * per:(r + 1) * per], 0) for r, b in enumerate(blocks)]
        repack = [rp for _, rp in packs]
        out_next = empty_cl(n, f, X, Y, Z, torch.float32, dev)
        if any(repack):
            w_all = [t.detach() for r in range(R) for t in params[r * per:r * per + nconv + 1]]
            p_all = [t for pk, _ in packs for t in pk]
            xv0, bv0, ov0 = view(x), view(slab[:n]), view(out_next)
            rc = lib.ws_trunk_repack_fwd(C.byref(geo), R, C.byref(xv0), C.byref(bv0), C.byref(ov0), _ptr_array(w_all),
                                         _ptr_array(p_all), stream_ptr())
            if rc == 0:
                repack = [0] * R
            elif rc != 1:
                check(rc, "ws_trunk_repack_fwd")
        for r, b in enumerate(blocks):
            if r in want_in:
                inputs[r] = h
            outer = inputs.pop(b["outer"]) if b["outer"] is not None else None
            desc = _lib.WsRdbDesc(n, X, Y, Z, f, gc, nconv, k, kl, cfg["slope"], b["alpha"], b["beta1"],
                                  b["beta2"] if outer is not None else 0.0, math_mode(), 0)
            p = params[r * per:(r + 1) * per]
            packed, desc.repack = packs[r][0], int(repack[r])
            buf = slab[r * n:(r + 1) * n]
            out = out_next if r == 0 else empty_cl(n, f, X, Y, Z, torch.float32, dev)
            if wsp is None:
                wsp = _workspace(int(lib.ws_rdb_forward_workspace_bytes(C.byref(desc))), dev)
            xv, bv, ov = view(h), view(buf), view(out)
            outer_v = view(outer) if outer is not None else null_view()
            wd = [t.detach() for t in p[:nconv + 1]]
            check(lib.ws_rdb_forward(C.byref(desc), C.byref(xv), C.byref(outer_v), C.byref(bv), C.byref(ov),
                                     _ptr_array(wd), _ptr_array(packed), ptr(p[nconv + 1]), wsp.data_ptr(), wsp.numel(),
                                     stream_ptr()), "ws_rdb_forward")
            h = out
        ctx.cfg = cfg
        ctx.geom = (n, X, Y, Z, f, gc, nconv, k, kl)
        ctx.save_for_backward(slab, *params)
        return h

    @staticmethod
    def backward(ctx, dy):
        lib = load()
        slab, *params = ctx.saved_tensors
        cfg = ctx.cfg
        nconv, blocks = cfg["nconv"], cfg["blocks"]
        R, per = len(blocks), nconv + 2
        n, X, Y, Z, f, gc, nconv, k, kl = ctx.geom
        ctot = f + nconv * gc
        cdt = act_dtype()
        dev = dy.device
        need = ctx.needs_input_grad
        need_params = any(need[2:])
        if dy.dtype != torch.float32 or not _linear_voxels(dy):
            dy = dy.float().contiguous(memory_format=torch.channels_last_3d)
        dbuf = empty_cl(n, ctot, X, Y, Z, torch.float32, dev)
        g_lff = torch.empty((R * n, X, Y, Z, f), dtype=cdt, device=dev).permute(0, 4, 1, 2, 3)
        g = torch.empty((R * n, X, Y, Z, nconv * gc), dtype=cdt, device=dev).permute(0, 4, 1, 2, 3)
        wsp = None
        pending = {}  # block index -> gradient that reaches its INPUT through an outer skip
        desc0 = None
        geo = _lib.WsRdbDesc(n, X, Y, Z, f, gc, nconv, k, kl, cfg["slope"], 1.0, 1.0, 0.0, math_mode(), 0)
        packs = [b["state"].buffers(geo, params[r * per:(r + 1) * per], 1) for r, b in enumerate(blocks)]
        repack = [rp for _, rp in packs]
        dx_first = empty_cl(n, f, X, Y, Z, torch.float32, dev)
        if any(repack):
            w_all = [t.detach() for r in range(R) for t in params[r * per:r * per + nconv + 1]]
            p_all = [t for pk, _ in packs for t in pk]
            dyv0, bv0, glv0, gv0, dxv0 = view(dy), view(slab[:n]), view(g_lff[:n]), view(g[:n]), view(dx_first)
            rc = lib.ws_trunk_repack_bwd(C.byref(geo), R, C.byref(dyv0), C.byref(bv0), C.byref(glv0), C.byref(gv0),
                                         C.byref(dxv0), _ptr_array(w_all), _ptr_array(p_all), stream_ptr())
            if rc == 0:
                repack = [0] * R
            elif rc != 1:
                check(rc, "ws_trunk_repack_bwd")
        for r in range(R - 1, -1, -1):
            b = blocks[r]
            has_outer = b["outer"] is not None
            desc = _lib.WsRdbDesc(n, X, Y, Z, f, gc, nconv, k, kl, cfg["slope"], b["alpha"], b["beta1"],
                                  b["beta2"] if has_outer else 0.0, math_mode(), 0)
            desc0 = desc
            p = params[r * per:(r + 1) * per]
            packed, desc.repack = packs[r][0], int(repack[r])
            if wsp is None:
                wsp = _workspace(int(lib.ws_rdb_backward_workspace_bytes(C.byref(desc))), dev)
            if has_outer:
                pending[b["outer"]] = (dy, b["beta2"])
            want_dx = r > 0 or need[0] or r in pending
            dx = None
            if want_dx:
                dx = dx_first if dx_first is not None else empty_cl(n, f, X, Y, Z, torch.float32, dev)
                dx_first = None
            dyv, bv, dbv = view(dy), view(slab[r * n:(r + 1) * n]), view(dbuf)
            glv, gv = view(g_lff[r * n:(r + 1) * n]), view(g[r * n:(r + 1) * n])
            dxv = view(dx) if dx is not None else null_view()
            wd = [t.detach() for t in p[:nconv + 1]]
            check(lib.ws_rdb_backward(C.byref(desc), C.byref(dyv), C.byref(bv), C.byref(dbv), C.byref(glv), C.byref(gv),
                                      C.byref(dxv), _ptr_array(wd), _ptr_array(packed), None, None,
                                      wsp.data_ptr(), wsp.numel(), stream_ptr(), None, None, 0), "ws_rdb_backward")
            if r in pending and dx is not None:
                skip, beta2 = pending.pop(r)
                axpby(dx, 1.0, skip, float(beta2), dx)
            dy = dx
        grads = [None] * (R * per)
        if need_params:
            rec = int(lib.ws_trunk_wgrad_record_floats(C.byref(desc0)))
            flat = torch.empty((R, rec), dtype=torch.float32, device=dev)
            nbytes = int(lib.ws_trunk_wgrad_workspace_bytes(C.byref(desc0), R))
            wws = _workspace(nbytes, dev)
            bv, gv, glv = view(slab[:n]), view(g[:n]), view(g_lff[:n])
            with _timed("trunk_wgrad", None):
                check(lib.ws_trunk_wgrad(C.byref(desc0), R, C.byref(bv), C.byref(gv), C.byref(glv), flat.data_ptr(), rec,
                                         wws.data_ptr(), wws.numel(), stream_ptr()), "ws_trunk_wgrad")
            for r in range(R):
                off = 0
                for i in range(per):
                    p = params[r * per + i]
                    if p is None:
                        continue
                    cnt = p.numel()
                    if need[2 + r * per + i]:
                        grads[r * per + i] = flat[r, off:off + cnt].view(p.shape)
                    off += cnt
        return (dy if need[0] else None, None, *grads)


# ------------------------------------------------------------------------------------------------------
# autograd: nearest upsample, concat, elementwise
# ------------------------------------------------------------------------------------------------------
class UpsampleFn(torch.autograd.Function):
    """nn.Upsample(scale_factor=(2,2,1), mode='nearest') (torch_blocks.py:347); backward = 2x2 block sum."""

    @staticmethod
    def forward(ctx, x):
        n, c, X, Y, Z = x.shape
        contig = x.is_contiguous() and not x.permute(0, 2, 3, 4, 1).is_contiguous()
        out = (torch.empty((n, c, 2 * X, 2 * Y, Z), dtype=x.dtype, device=x.device) if contig
               else empty_cl(n, c, 2 * X, 2 * Y, Z, x.dtype, x.device))
        upsample_fwd(x, out)
        ctx.in_shape = x.shape
        ctx.contig = contig
        return out

    @staticmethod
    def backward(ctx, dy):
        n, c, X, Y, Z = ctx.in_shape
        din = (torch.empty(ctx.in_shape, dtype=dy.dtype, device=dy.device) if ctx.contig
               else empty_cl(n, c, X, Y, Z, dy.dtype, dy.device))
        upsample_bwd(dy, din)
        return din


class CatFn(torch.autograd.Function):
    """torch.cat((a, b), 1) (Generator…py:228) where a and b were already written into slices of `buf`."""

    @staticmethod
    def forward(ctx, a, b, buf):
        ctx.ca = a.shape[1]
        return buf.detach()

    @staticmethod
    def backward(ctx, dy):
        return dy[:, :ctx.ca], dy[:, ctx.ca:], None


class AxpbyFn(torch.autograd.Function):
    """a*x + b*y with one kernel (RRDB / skip adds, torch_blocks.py:46,330)."""

    @staticmethod
    def forward(ctx, x, y, a, b):
        out = empty_cl(*x.shape, x.dtype, x.device)
        axpby(x, a, y, b, out)
        ctx.ab = (a, b)
        return out

    @staticmethod
    def backward(ctx, d):
        a, b = ctx.ab
        return (d if a == 1.0 else d * a), (d if b == 1.0 else d * b), None, None


class ChanScaleFn(torch.autograd.Function):
    """Standalone Dropout3d: x * scale[n, c] (Generator…py:70-74, Discriminator_3D.py:178-182)."""

    @staticmethod
    def forward(ctx, x, scale):
        out = empty_cl(*x.shape, x.dtype, x.device)
        lrelu_bwd(x, x, 1.0, out, chan_scale=scale)
        ctx.save_for_backward(scale)
        return out

    @staticmethod
    def backward(ctx, d):
        (scale,) = ctx.saved_tensors
        g = empty_cl(*d.shape, d.dtype, d.device)
        lrelu_bwd(d, d, 1.0, g, chan_scale=scale)
        return g, None


class LReluFn(torch.autograd.Function):
    """Standalone LeakyReLU (only reached when a caller slices a block apart)."""

    @staticmethod
    def forward(ctx, x, slope):
        out = empty_cl(*x.shape, x.dtype, x.device)
        lrelu_bwd(x, x, slope, out)  # x>0 ? x : slope*x
        ctx.save_for_backward(out)
        ctx.slope = slope
        return out

    @staticmethod
    def backward(ctx, d):
        (y,) = ctx.saved_tensors
        g = empty_cl(*d.shape, d.dtype, d.device)
        lrelu_bwd(d, y, ctx.slope, g)
        return g, None


# ------------------------------------------------------------------------------------------------------
# autograd: conv + BatchNorm3d(train) + LeakyReLU  (discriminator blocks, torch_blocks.py:372-521)
# ------------------------------------------------------------------------------------------------------
class ConvBNLReluFn(torch.autograd.Function):
    """Training-mode BatchNorm3d: the conv epilogue accumulates per-channel sum / sum-of-squares of its raw
    fp32 accumulators, ws_bn_finalize turns them into scale/shift (+ running-stat update), a second
    elementwise pass normalises and applies LeakyReLU."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, cfg):
        lib = load()
        x = _as_act(x)
        shape = make_shape(x.shape, weight.shape[0], weight.shape[2:], cfg["stride"], cfg["padding"])
        xo, yo, zo = out_dims(shape)
        c = shape.cout
        dev = x.device
        cdt = act_dtype()
        raw = empty_cl(shape.n, c, xo, yo, zo, cdt, dev)
        stats = torch.zeros((2, c), dtype=torch.float32, device=dev)
        conv_fwd(x, weight, cfg.get("cache"), shape, raw, stat_sum=stats[0], stat_sqsum=stats[1])
        aux = torch.empty((4, c), dtype=torch.float32, device=dev)  # scale, shift, mean, invstd
        count = shape.n * xo * yo * zo
        check(lib.ws_bn_finalize(stats[0].data_ptr(), stats[1].data_ptr(), count, c, ptr(gamma), ptr(beta),
                                 cfg["eps"], cfg["momentum"], ptr(running_mean), ptr(running_var),
                                 aux[0].data_ptr(), aux[1].data_ptr(), aux[2].data_ptr(), aux[3].data_ptr(),
                                 stream_ptr()), "ws_bn_finalize")
        y = empty_cl(shape.n, c, xo, yo, zo, cdt, dev)
        rv, yv = view(raw), view(y)
        check(lib.ws_scale_shift_lrelu(C.byref(rv), aux[0].data_ptr(), aux[1].data_ptr(), cfg["slope"],
                                       C.byref(yv), shape.n, c, xo * yo * zo, stream_ptr()),
              "ws_scale_shift_lrelu")
        ctx.shape, ctx.cfg, ctx.count = shape, cfg, count
        ctx.save_for_backward(x, weight, gamma, raw, y, aux)
        ctx.mark_non_differentiable()
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = load()
        x, weight, gamma, raw, y, aux = ctx.saved_tensors
        shape, cfg = ctx.shape, ctx.cfg
        c = shape.cout
        dev = dy.device
        cdt = act_dtype()
        n, _, xo, yo, zo = y.shape
        v = xo * yo * zo
        sums = torch.empty((2, c), dtype=torch.float32, device=dev)
        dv, yv, rv = view(dy), view(y), view(raw)
        check(lib.ws_bn_lrelu_bwd_reduce(C.byref(dv), C.byref(yv), C.byref(rv), aux[2].data_ptr(),
                                         aux[3].data_ptr(), cfg["slope"], sums[0].data_ptr(), sums[1].data_ptr(),
                                         n, c, v, stream_ptr()), "ws_bn_lrelu_bwd_reduce")
        g = empty_cl(n, c, xo, yo, zo, cdt, dev)
        gv = view(g)
        check(lib.ws_bn_lrelu_bwd_apply(C.byref(dv), C.byref(yv), C.byref(rv), aux[2].data_ptr(),
                                        aux[3].data_ptr(), ptr(gamma), sums[0].data_ptr(), sums[1].data_ptr(),
                                        cfg["slope"], ctx.count, C.byref(gv), n, c, v, stream_ptr()),
              "ws_bn_lrelu_bwd_apply")
        need_x, need_w, need_g, need_b = ctx.needs_input_grad[:4]
        dx = dw = None
        if need_w:
            dw, _ = conv_wgrad(x, g, shape)
        if need_x:
            dx = empty_cl(*x.shape, cdt, dev)
            conv_dgrad(g, weight, cfg.get("cache"), shape, dx)
        dgamma = sums[1].clone() if need_g else None
        dbeta = sums[0].clone() if need_b else None
        return dx, dw, dgamma, dbeta, None, None, None


# ------------------------------------------------------------------------------------------------------
# wind-field loss stencils
# ------------------------------------------------------------------------------------------------------
_COEF_CACHE = {}


def axis_coeffs(coords: torch.Tensor) -> torch.Tensor:
    """6 floats per grid point: forward-form and linear-form 3-point coefficients of
    torch.gradient(spacing=coords) (process_data.py:303)."""
    key = (coords.data_ptr(), coords._version, coords.numel(), str(coords.device))
    hit = _COEF_CACHE.get(key)
    cap = capturing()
    if hit is not None and not cap:
        return hit
    c32 = coords.detach().to(torch.float32).contiguous()
    coef = torch.empty((coords.numel(), 6), dtype=torch.float32, device=coords.device)
    check(load().ws_axis_coeffs(c32.data_ptr(), c32.numel(), coef.data_ptr(), stream_ptr()), "ws_axis_coeffs")
    if cap:
        keepalive(c32, coef)
        return coef
    if len(_COEF_CACHE) > 64:
        _COEF_CACHE.clear()
    _COEF_CACHE[key] = coef
    return coef


def wind_gradient(field: torch.Tensor, x: torch.Tensor, y: torch.Tensor, Z: torch.Tensor) -> torch.Tensor:
    """calculate_gradient_of_wind_field (process_data.py:301-313): (N,3,X,Y,Z) -> (N,9,X,Y,Z), no autograd."""
    _require_cuda(field)
    n, c, X, Y, Zn = field.shape
    if c != 3:
        raise _lib.WindSRError("wind_gradient expects the 3 wind components")
    out = torch.empty((n, 9, X, Y, Zn), dtype=torch.float32, device=field.device)
    cx, cy = axis_coeffs(x), axis_coeffs(y)
    fv, zv, ov = view(field.float()), view(Z.float()), view(out)
    check(load().ws_wind_gradient(C.byref(fv), C.byref(zv), cx.data_ptr(), cy.data_ptr(), C.byref(ov), n, X, Y,
                                  Zn, stream_ptr()), "ws_wind_gradient")
    return out


class WindLossFn(torch.autograd.Function):
    """One fused pass over HR, SR, Z -> the 16-slot vector of sums and maxes (windsr.h WS_WL_*).

    The scalar loss formula (normalisers, weights, NaN guard — wind_field_GAN_3D.py:388-454) is written with
    ordinary torch scalar ops on these slots; autograd hands dL/d(slot) back to ``backward``, which is exactly
    the coefficient vector ws_windloss_bwd needs, including the path through SR_max/100.
    """

    @staticmethod
    def forward(ctx, hr, sr, Z, x, y):
        _require_cuda(sr)
        n, c, X, Y, Zn = sr.shape
        hr3 = hr[:, :3]
        result = torch.empty((_lib.WL_RESULT_FLOATS,), dtype=torch.float32, device=sr.device)
        argmax = torch.empty((4,), dtype=torch.int64, device=sr.device)
        cx, cy = axis_coeffs(x), axis_coeffs(y)
        hv, sv, zv = view(hr3), view(sr), view(Z)
        check(load().ws_windloss_fwd(C.byref(hv), C.byref(sv), C.byref(zv), cx.data_ptr(), cy.data_ptr(), n, X,
                                     Y, Zn, result.data_ptr(), argmax.data_ptr(), stream_ptr()),
              "ws_windloss_fwd")
        ctx.save_for_backward(hr3, sr, Z, cx, cy, argmax)
        return result[:_lib.WL_SLOTS].clone()

    @staticmethod
    def backward(ctx, dres):
        hr3, sr, Z, cx, cy, argmax = ctx.saved_tensors
        n, c, X, Y, Zn = sr.shape
        coef = torch.zeros((_lib.WLB_SLOTS,), dtype=torch.float32, device=sr.device)
        # cotangents of loss terms the caller dropped with a device-side select arrive as 0 * inf = NaN: they
        # stand for "this term is not in the loss" (wind_field_GAN_3D.py:434-443), i.e. zero
        dres = torch.where(torch.isfinite(dres), dres, torch.zeros_like(dres))
        coef[0:6] = dres[0:6]
        coef[6:10] = dres[7:14:2]  # SR max slots 7, 9, 11, 13
        dsr = torch.empty(sr.shape, dtype=torch.float32, device=sr.device)
        nbytes = load().ws_windloss_bwd_workspace_bytes(n, X, Y, Zn)
        wsp = torch.empty(nbytes, dtype=torch.uint8, device=sr.device)
        hv, sv, zv, dv = view(hr3), view(sr), view(Z), view(dsr)
        check(load().ws_windloss_bwd(C.byref(hv), C.byref(sv), C.byref(zv), cx.data_ptr(), cy.data_ptr(), n, X, Y,
                                     Zn, coef.data_ptr(), argmax.data_ptr(), C.byref(dv), wsp.data_ptr(),
                                     wsp.numel(), stream_ptr()), "ws_windloss_bwd")
        return None, dsr, None, None, None


def windloss_slots(hr, sr, Z, x, y) -> torch.Tensor:
    return WindLossFn.apply(hr, sr.float() if sr.dtype != torch.float32 else sr, Z.float(), x, y)


# ------------------------------------------------------------------------------------------------------
# instance noise, validation metrics, input pipeline (csrc/train_aux.cu)
# ------------------------------------------------------------------------------------------------------
_NOISE_STATE = {}


def _noise_state(device):
    st = _NOISE_STATE.get(device)
    if st is None:
        # seeded from torch's generator so torch.manual_seed() makes runs reproducible; [call counter, block ticket]
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        st = (seed, torch.zeros(2, dtype=torch.int64, device=device))
        _NOISE_STATE[device] = st
    return st


def reseed_instance_noise(seed: int, device=None) -> None:
    for dev in list(_NOISE_STATE) if device is None else [device]:
        _NOISE_STATE[dev] = (int(seed), torch.zeros(2, dtype=torch.int64, device=dev))


def add_instance_noise(x: torch.Tensor, scale: float = 1.0, scale_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x + U[0,1) * scale (* scale_dev[0]) — tools/trainingtricks.py:49-58 fused with the add at
    wind_field_GAN_3D.py:250-299.  Differentiable w.r.t. x (the noise is a constant)."""
    return _InstanceNoiseFn.apply(x, float(scale), scale_dev)


class _InstanceNoiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, scale_dev):
        _require_cuda(x)
        xc = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
        out = torch.empty_like(xc)
        seed, state = _noise_state(x.device)
        check(load().ws_instance_noise(xc.data_ptr(), out.data_ptr(), xc.numel(), scale, ptr(scale_dev), seed,
                                       state.data_ptr(), stream_ptr()), "ws_instance_noise")
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None, None


def validation_metrics(HR: torch.Tensor, SR: Optional[torch.Tensor], LR: torch.Tensor) -> torch.Tensor:
    """One fused pass -> float64[4]: sum (HR-SR)^2, sum (HR-tri)^2, sum |HR-tri|, sum |HR-SR| with tri the
    align-corners trilinear upsample of LR[:, :3] (wind_field_GAN_3D.py:730-770, 597-618); nothing materialised."""
    _require_cuda(HR)
    n, c, X, Y, Zn = HR.shape
    if c != 3 or LR.shape[1] < 3 or LR.shape[4] != Zn:
        raise _lib.WindSRError("validation_metrics expects HR (N,3,X,Y,Z) and LR (N,>=3,x,y,Z)")
    sums = torch.empty(4, dtype=torch.float64, device=HR.device)
    hv, lv = view(HR.float()), view(LR.float()[:, :3])
    sv = view(SR.float()) if SR is not None else null_view()
    check(load().ws_validation_metrics(C.byref(hv), C.byref(sv), C.byref(lv), n, X, Y, Zn, LR.shape[2], LR.shape[3],
                                       sums.data_ptr(), stream_ptr()), "ws_validation_metrics")
    return sums


def prepare_batch(u, v, w, z, *, pressure=None, z_above_ground=None, aug=None, crop=None, coarseness=4,
                  include_pressure=False, include_z_channel=False, include_above_ground_channel=False,
                  uvw_max=1.0, p_min=0.0, p_max=1.0, z_min=0.0, z_max=1.0, z_above_ground_max=1.0):
    """(LR, HR, Z) of a training batch from float64 device fields (N, X, Y, Z) — the crop / normalise / subsample /
    rot90 / flip chain of process_data.py:159-262,420-494 as one gather kernel (``ws_prepare_batch``).
    aug: int32 (N,5) = x_start, y_start, rotations, flip_x, flip_y (zeros when None); crop: (X, Y) slice size."""
    _require_cuda(u)
    n, SX, SY, SZ = u.shape
    X, Y = crop if crop is not None else (SX, SY)
    for t in (u, v, w, z, pressure, z_above_ground):
        if t is not None and (t.dtype != torch.float64 or not t.is_contiguous() or tuple(t.shape) != (n, SX, SY, SZ)):
            raise _lib.WindSRError("prepare_batch: fields must be contiguous float64 (N, X, Y, Z) of one shape")
    if aug is None:
        aug = torch.zeros((n, 5), dtype=torch.int32, device=u.device)
    aug = aug.to(device=u.device, dtype=torch.int32).contiguous()
    if X != Y and bool((aug[:, 2] % 2 == 1).any()):
        raise _lib.WindSRError("prepare_batch: odd rotations need a square crop")
    cf = int(coarseness)
    xl, yl = (X + cf - 1) // cf, (Y + cf - 1) // cf
    lr_c = 3 + int(bool(include_pressure)) + (0 if not include_z_channel else (2 if include_above_ground_channel else 1))
    LR = torch.empty((n, lr_c, xl, yl, SZ), dtype=torch.float32, device=u.device)
    HR = torch.empty((n, 3, X, Y, SZ), dtype=torch.float32, device=u.device)
    Zo = torch.empty((n, 1, X, Y, SZ), dtype=torch.float32, device=u.device)
    d = _lib.WsPrepareDesc(n, SX, SY, SZ, X, Y, cf, int(bool(include_pressure)), int(bool(include_z_channel)),
                           int(bool(include_above_ground_channel)), SX * SY * SZ, float(uvw_max), float(p_min),
                           float(p_max), float(z_min), float(z_max), float(z_above_ground_max))
    check(load().ws_prepare_batch(C.byref(d), u.data_ptr(), v.data_ptr(), w.data_ptr(), ptr(pressure), z.data_ptr(),
                                  ptr(z_above_ground), aug.data_ptr(), LR.data_ptr(), HR.data_ptr(), Zo.data_ptr(),
                                  stream_ptr()), "ws_prepare_batch")
    return LR, HR, Zo
