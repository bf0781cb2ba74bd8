"""Drop-in building blocks for the RRDB generator and the 3-D discriminator.

Same factory / class names, constructor arguments and ``state_dict`` keys as the reference's
``CNN_models/torch_blocks.py`` (3D mode), but every forward runs the hand-written sm_100a
kernels behind ``libwindsr.so`` with the elementwise tail fused into the convolution epilogue:

=========================  ===============================  ==========================================
reference (file:line)      reference arithmetic             here
=========================  ===============================  ==========================================
torch_blocks.py:5-37       Conv3d -> [BN] -> LeakyReLU      ``ConvBlock``: one conv launch, LReLU / eval-BN
                                                            affine in the epilogue; train-BN = stats in the
                                                            epilogue + finalize + one normalise pass
torch_blocks.py:40-46      x + module(x)                    residual add in the last conv's epilogue
torch_blocks.py:192-214    cat((x, lrelu(conv(x))), 1)      epilogue writes the concat-buffer channel slice
torch_blocks.py:217-290    RDB: clone, 4 convs, LFF, .2r+x  ``ops.RDBFn`` (one autograd node)
torch_blocks.py:293-330    RRDB: 3 RDB, .2r+x               folded into the 3rd RDB's LFF epilogue
torch_blocks.py:333-369    Upsample(nearest,(2,2,1)) conv   vectorised upsample kernel + ``ConvBlock``
torch_blocks.py:372-521    discriminator block              two ``ConvBlock``s
=========================  ===============================  ==========================================

The experimental ``Horizontal_Conv_3D`` / ``horizontal_3D`` / ``2D`` modes of the reference are out of scope
(no shipped config uses them) and raise ``NotImplementedError``.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
from torch import nn

from .. import ops


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


class Conv3d(nn.Module):
    """Parameter container + standalone forward of one Conv3d layer.

    Deliberately *named* ``Conv3d``: the reference's ``tools/initialization.py:16`` selects layers by
    ``m.__class__.__name__``.  Parameters have torch's (Cout, Cin, kX, kY, kZ) fp32 layout and nn.Conv3d's
    default initialisation (same RNG consumption), so seeded construction and checkpoints are interchangeable.
    """

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _triple(kernel_size), _triple(stride), _triple(padding)
        self.weight = nn.Parameter(torch.empty((out_channels, in_channels) + self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self._packed = ops.PackedWeights()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def __deepcopy__(self, memo):
        # the packed-weight cache is device scratch, not state (deepcopy: wind_field_GAN_3D.py:581)
        new = type(self)(self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding,
                         bias=self.bias is not None)
        new.weight = nn.Parameter(self.weight.detach().clone(), requires_grad=self.weight.requires_grad)
        if self.bias is not None:
            new.bias = nn.Parameter(self.bias.detach().clone(), requires_grad=self.bias.requires_grad)
        new.train(self.training)
        return new

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, "
                f"stride={self.stride}, padding={self.padding}, bias={self.bias is not None}")

    def run(self, x, **kw):
        """conv + fused epilogue options (see ops.conv_block)."""
        if ops.get_precision() == "bf16" and x.dtype == torch.float32 and self.in_channels >= 16:
            x = _CastFn.apply(x, torch.bfloat16)  # tensor-core operand format; one small cast pass
        return ops.conv_block(x, self.weight, self.bias, stride=self.stride, padding=self.padding,
                              cache=self._packed, **kw)

    def forward(self, x):
        return self.run(x)


class DenseConv3d(Conv3d):
    """The Conv3d inside an ``RDB_Conv``.  The reference wraps every RDB_Conv in ``torch.jit.script``
    (torch_blocks.py:259), which turns its layers into ``RecursiveScriptModule``s — so the class-name match of
    ``init_weights`` (tools/initialization.py:16) silently SKIPS the 4 x 48 dense convs and they keep
    nn.Conv3d's default Kaiming-uniform initialisation.  A distinct class name reproduces that behaviour under
    both the reference's and this package's ``init_weights`` (and keeps the seeded RNG stream aligned)."""


class LeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.01):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, x):
        return ops.LReluFn.apply(x, self.negative_slope)

    def extra_repr(self):
        return f"negative_slope={self.negative_slope}"


class BatchNorm3d(nn.BatchNorm3d):
    """State holder (weight, bias, running_mean, running_var, num_batches_tracked — same keys as
    nn.BatchNorm3d, torch_blocks.py:24-25).  Inside a ``ConvBlock`` its arithmetic is done by the conv
    epilogue / BN kernels; a standalone call is not part of the hot path and is not supported."""

    def forward(self, x):
        raise NotImplementedError("BatchNorm3d is fused into its ConvBlock; call the block, not the layer")


class Upsample(nn.Module):
    """nn.Upsample(scale_factor=(s, s, 1), mode='nearest') for s = 2 (torch_blocks.py:347)."""

    def __init__(self, scale_factor=(2, 2, 1), mode="nearest"):
        super().__init__()
        self.scale_factor, self.mode = tuple(scale_factor), mode
        if self.mode != "nearest" or self.scale_factor != (2, 2, 1):
            raise NotImplementedError("only nearest (2,2,1) upsampling is on the hot path")

    def forward(self, x):
        return ops.UpsampleFn.apply(x)

    def extra_repr(self):
        return f"scale_factor={self.scale_factor}, mode={self.mode}"


class Dropout3d(nn.Module):
    """Channel dropout (Generator…py:70-74, Discriminator_3D.py:178-182): one Bernoulli draw per (n, c),
    kept channels scaled by 1/(1-p).  ``sample`` returns the (N*C,) fp32 scale vector that the preceding
    conv folds into its epilogue; ``forward`` is the standalone (unfused) form."""

    def __init__(self, p=0.5):
        super().__init__()
        self.p = float(p)

    def sample(self, n, c, device):
        if not self.training or self.p == 0.0:
            return None
        keep = 1.0 - self.p
        if keep <= 0.0:
            return torch.zeros(n * c, dtype=torch.float32, device=device)
        return torch.bernoulli(torch.full((n * c,), keep, dtype=torch.float32, device=device)) / keep

    def forward(self, x):
        scale = self.sample(x.shape[0], x.shape[1], x.device)
        return x if scale is None else ops.ChanScaleFn.apply(x, scale)

    def extra_repr(self):
        return f"p={self.p}"


class ConvBlock(nn.Sequential):
    """``Sequential(Conv3d[, BatchNorm3d][, LeakyReLU])`` executed as one fused launch sequence."""

    def _parts(self):
        conv = bn = act = None
        for m in self:
            if isinstance(m, Conv3d):
                conv = m
            elif isinstance(m, nn.BatchNorm3d):
                bn = m
            elif isinstance(m, (LeakyReLU, nn.LeakyReLU)):
                act = m
            else:
                return None
        return conv, bn, act

    def forward(self, x, **kw):
        parts = self._parts()
        if parts is None or parts[0] is None:  # a caller sliced the block apart
            for m in self:
                x = m(x)
            return x
        conv, bn, act = parts
        slope = act.negative_slope if act is not None else 1.0
        if bn is None:
            return conv.run(x, slope=slope, **kw)
        if bn.training:
            if bn.momentum is None:
                raise NotImplementedError("cumulative-average BatchNorm momentum is not used by the reference")
            y = ops.ConvBNLReluFn.apply(x, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                        dict(stride=conv.stride, padding=conv.padding, slope=slope, eps=bn.eps,
                                             momentum=bn.momentum, cache=conv._packed))
            bn.num_batches_tracked += 1
            return y
        # eval: BN is a per-channel affine of the accumulator -> conv epilogue
        with torch.no_grad():
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias - bn.running_mean * scale
        if bn.weight.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("eval-mode BatchNorm with trainable affine is not on the reference's path")
        return ops.conv_block(x, conv.weight, None, stride=conv.stride, padding=conv.padding, slope=slope,
                              oscale=scale.contiguous(), shift=shift.contiguous(), cache=conv._packed, **kw)


def create_conv_lrelu_layer(in_channels, out_channels, kernel_size, stride=1, padding=1,
                            lrelu_negative_slope=0.2, normalization_type="", layer_type=Conv3d, lrelu=True):
    """torch_blocks.py:5-37 (3-D only)."""
    if layer_type not in (Conv3d, DenseConv3d, nn.Conv3d):
        raise NotImplementedError("only 3-D convolutions are on the hot path (SURVEY §2)")
    cls = DenseConv3d if layer_type is DenseConv3d else Conv3d
    layers = [cls(in_channels, out_channels, kernel_size, stride, padding, bias=False)]
    if normalization_type:
        if normalization_type == "batch":
            layers.append(BatchNorm3d(out_channels))
        elif normalization_type == "instance":
            raise NotImplementedError("InstanceNorm3d is not used by any shipped config (SURVEY §2)")
        else:
            raise NotImplementedError(f"Unknown norm type {normalization_type}")
    if lrelu:
        layers.append(LeakyReLU(negative_slope=lrelu_negative_slope))
    return ConvBlock(*layers)


class SkipConnectionBlock(nn.Module):
    """x + module(x) (torch_blocks.py:40-46).  When the wrapped module ends in a plain ConvBlock the add is
    that conv's epilogue residual; otherwise one axpby launch."""

    def __init__(self, submodule):
        super().__init__()
        self.module = submodule

    def forward(self, x):
        mod = self.module
        if isinstance(mod, nn.Sequential) and len(mod) > 0 and isinstance(mod[-1], ConvBlock):
            last = mod[-1]
            parts = last._parts()
            if parts is not None and parts[1] is None and parts[2] is None:
                x32 = _to_f32(x)
                h = run_trunk(list(mod)[:-1], x32)
                return parts[0].run(h, res=x32, beta=1.0, out_dtype=torch.float32)
        return ops.AxpbyFn.apply(_to_f32(x), _to_f32(mod(x)), 1.0, 1.0)


class _CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        out = ops.empty_cl(*x.shape, dtype, x.device)
        ops.copy_(x, out)
        ctx.in_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, d):
        out = ops.empty_cl(*d.shape, ctx.in_dtype, d.device)
        ops.copy_(d, out)
        return out, None


def _to_f32(x):
    return x if x.dtype == torch.float32 else _CastFn.apply(x, torch.float32)


class RDB_Conv(nn.Module):
    """cat((x, lrelu(conv(x))), 1) (torch_blocks.py:192-214).  Inside an ``RDB`` the concat is a slice write;
    standalone it allocates the wider buffer itself."""

    def __init__(self, in_channels, out_channels, kernel_size=3, lrelu_negative_slope=0.2, layer_type=Conv3d):
        super().__init__()
        self.conv = create_conv_lrelu_layer(in_channels, out_channels, kernel_size, stride=1,
                                            padding=(kernel_size - 1) // 2,
                                            lrelu_negative_slope=lrelu_negative_slope, layer_type=DenseConv3d)

    def forward(self, x):
        n, c, X, Y, Z = x.shape
        conv = self.conv[0]
        buf = ops.empty_cl(n, c + conv.out_channels, X, Y, Z, ops.act_dtype(), x.device)
        a = _CopyIntoFn.apply(x, buf[:, :c])
        b = self.conv(a, out=buf[:, c:])
        return ops.CatFn.apply(a, b, buf)


class _CopyIntoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dst):
        ops.copy_(x, dst)
        ctx.in_dtype = x.dtype
        return dst

    @staticmethod
    def backward(ctx, d):
        out = ops.empty_cl(*d.shape, ctx.in_dtype, d.device)
        ops.copy_(d, out)
        return out, None


class RDB(nn.Module):
    """Residual dense block (torch_blocks.py:217-290); children ``conv0..conv{k-2}`` and ``LFF`` keep the
    reference's names so checkpoints load unchanged."""

    def __init__(self, in_channels, growth_channels, number_of_conv_layers, lff_kern_size=1,
                 lrelu_negative_slope=0.2, residual_scaling=0.2, mode="2D"):
        super().__init__()
        if mode != "3D":
            raise NotImplementedError(f"RDB mode {mode!r}: only '3D' is on the hot path (SURVEY §2)")
        self.residual_scaling = residual_scaling
        self.lrelu_negative_slope = lrelu_negative_slope
        for i in range(number_of_conv_layers - 1):
            self.add_module(f"conv{i}", RDB_Conv(in_channels + i * growth_channels, growth_channels,
                                                 lrelu_negative_slope=lrelu_negative_slope))
        if lff_kern_size <= 0 or lff_kern_size % 2 == 0:
            raise ValueError("LFF kernel size (lff_kern_size) must be an odd number > 0")
        self.LFF = Conv3d(in_channels + (number_of_conv_layers - 1) * growth_channels, in_channels,
                          kernel_size=lff_kern_size, padding=(lff_kern_size - 1) // 2)

    def _dense_convs(self):
        return [m.conv[0] for name, m in self._modules.items() if name.startswith("conv")]

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for key, value in self.__dict__.items():
            if key != "_rdb_state":  # device scratch (packed weights), not module state
                setattr(new, key, copy.deepcopy(value, memo))
        return new

    def run(self, x, outer=None, outer_scale=None):
        """residual_scaling * LFF(dense(x)) + x, optionally composed with the enclosing RRDB's
        ``outer_scale * (.) + outer`` (torch_blocks.py:328-330) in the same epilogue."""
        convs = self._dense_convs()
        x = _to_f32(x)
        if outer is None:
            alpha, beta1, beta2 = self.residual_scaling, 1.0, 0.0
        else:
            alpha, beta1, beta2 = self.residual_scaling * outer_scale, outer_scale, 1.0
            outer = _to_f32(outer)
        if not hasattr(self, "_rdb_state"):
            self._rdb_state = ops.RDBState()
        cfg = dict(nconv=len(convs), slope=self.lrelu_negative_slope, alpha=alpha, beta1=beta1, beta2=beta2,
                   state=self._rdb_state)
        params = [c.weight for c in convs] + [self.LFF.weight, self.LFF.bias]
        return ops.RDBFn.apply(x, outer, cfg, *params)

    def forward(self, x):
        return self.run(x)


class RRDB(nn.Module):
    """Residual-in-residual dense block (torch_blocks.py:293-330)."""

    def __init__(self, in_channels, growth_channels, num_convs, lff_kern_size=1, lrelu_negative_slope=0.2,
                 RDB_residual_scaling=0.2, RRDB_residual_scaling=0.2, number_of_RDBs=3, mode="2D"):
        super().__init__()
        self.RRDB_residual_scaling = RRDB_residual_scaling
        self.RDBs = nn.Sequential(*[
            RDB(in_channels, growth_channels, num_convs, lrelu_negative_slope=lrelu_negative_slope,
                residual_scaling=RDB_residual_scaling, lff_kern_size=lff_kern_size, mode=mode)
            for _ in range(number_of_RDBs)])

    def forward(self, x):
        x = _to_f32(x)
        h = x
        blocks = list(self.RDBs)
        for i, rdb in enumerate(blocks):
            if i == len(blocks) - 1:
                return rdb.run(h, outer=x, outer_scale=self.RRDB_residual_scaling)
            h = rdb.run(h)
        return ops.AxpbyFn.apply(h, x, self.RRDB_residual_scaling, 1.0)  # number_of_RDBs == 0


def _trunk_signature(m):
    """Geometry key of an RRDB whose RDBs ops.TrunkFn can batch, else None."""
    if not isinstance(m, RRDB) or len(m.RDBs) == 0:
        return None
    sig = None
    for rdb in m.RDBs:
        if not isinstance(rdb, RDB) or rdb.LFF.bias is None:
            return None
        convs = rdb._dense_convs()
        if len(convs) < 2 or any(c.bias is not None or c.stride != (1, 1, 1) for c in convs):
            return None
        k = convs[0].kernel_size
        if k[0] != k[1] or k[0] != k[2] or rdb.LFF.kernel_size != (1, 1, 1):
            return None
        f, gc = rdb.LFF.out_channels, convs[0].out_channels
        if any(c.in_channels != f + i * gc or c.out_channels != gc for i, c in enumerate(convs)):
            return None
        if rdb.LFF.in_channels != f + len(convs) * gc:
            return None
        s = (rdb.LFF.out_channels, convs[0].out_channels, len(convs), k[0], rdb.lrelu_negative_slope,
             tuple(c.in_channels for c in convs), tuple(c.kernel_size for c in convs))
        if sig is not None and s != sig:
            return None
        sig = s
    return sig


def run_trunk(mods, h):
    """Apply a list of trunk modules; maximal runs of RRDBs with identical RDB geometry go through ``ops.TrunkFn``
    (one autograd node, batched weight gradients) when gradients are being recorded on the BF16 tensor-core path,
    every other case through the modules' own forward.  ``WINDSR_TRUNK_GROUPS`` splits a run into that many nodes
    (their parameter gradients — and with them the gradient all-reduce — become available earlier)."""
    import os
    i = 0
    use = (ops.trunk_batched_enabled() and torch.is_grad_enabled() and h.is_cuda and ops.get_precision() == "bf16")
    # (measured on 2 GPUs: 1 node 37.6 ms, 2 nodes 37.9, 4 nodes 38.6 — the smaller batched launches cost more than the
    # earlier all-reduce gains; ops.ConvFn's late weight gradients hide the trunk's all-reduce instead)
    groups = max(1, int(os.environ.get("WINDSR_TRUNK_GROUPS", "1")))
    while i < len(mods):
        sig = _trunk_signature(mods[i]) if use else None
        j = i
        if sig is not None:
            while j < len(mods) and _trunk_signature(mods[j]) == sig:
                j += 1
        if j > i and ops.trunk_supported(h, sig):
            per = -(-(j - i) // groups)
            for a in range(i, j, per):
                h = _trunk_apply(mods[a:min(j, a + per)], h, sig)
            i = j
        else:
            h = mods[i](h)
            i += 1
    return h


def _trunk_apply(rrdbs, h, sig):
    blocks, params = [], []
    for m in rrdbs:
        first = len(blocks)
        rdbs = list(m.RDBs)
        for t, rdb in enumerate(rdbs):
            if not hasattr(rdb, "_rdb_state"):
                rdb._rdb_state = ops.RDBState()
            last = t == len(rdbs) - 1
            if last:
                blocks.append(dict(alpha=rdb.residual_scaling * m.RRDB_residual_scaling, beta1=m.RRDB_residual_scaling,
                                   beta2=1.0, outer=first, state=rdb._rdb_state))
            else:
                blocks.append(dict(alpha=rdb.residual_scaling, beta1=1.0, beta2=0.0, outer=None, state=rdb._rdb_state))
            params += [c.weight for c in rdb._dense_convs()] + [rdb.LFF.weight, rdb.LFF.bias]
    cfg = dict(nconv=sig[2], slope=sig[4], blocks=blocks)
    return ops.TrunkFn.apply(_to_f32(h), cfg, *params)


class UpConvBlock(nn.Sequential):
    """Sequential(Upsample, ConvBlock) (torch_blocks.py:341-356)."""


def create_UpConv_block(in_channels, out_channels, scale, lrelu_negative_slope=0.2, mode="2D",
                        number_of_z_layers=10):
    if mode != "3D":
        raise NotImplementedError(f"Unknown / unsupported UpConv mode {mode}")
    return UpConvBlock(
        Upsample(scale_factor=(scale, scale, 1), mode="nearest"),
        create_conv_lrelu_layer(in_channels, out_channels, kernel_size=3, padding=1,
                                lrelu_negative_slope=lrelu_negative_slope))


def create_discriminator_block(in_channels, out_channels, feat_kern_size=3, lrelu_negative_slope=0.2,
                               normalization_type="batch", drop_first_norm=False, mode="2D",
                               number_of_z_layers=10, halve_z_dim=True):
    """torch_blocks.py:372-521 (3-D): k3/k5 conv (+BN unless drop_first_norm) + LReLU, then the strided
    (4,4,k) conv + BN + LReLU with stride 2 (halve_z_dim) or (2,2,1)."""
    if feat_kern_size == 5:
        feat_pad = 2
    elif feat_kern_size == 3:
        feat_pad = 1
    else:
        raise NotImplementedError("Only supported kern sizes are 3 and 5")
    if mode != "3D":
        raise NotImplementedError("Only the 3D mode is on the hot path (SURVEY §2)")
    first = create_conv_lrelu_layer(in_channels, out_channels, kernel_size=feat_kern_size,
                                    lrelu_negative_slope=lrelu_negative_slope, padding=feat_pad, stride=1,
                                    normalization_type="" if drop_first_norm else normalization_type)
    second = create_conv_lrelu_layer(out_channels, out_channels, kernel_size=(4, 4, feat_kern_size),
                                     lrelu_negative_slope=lrelu_negative_slope,
                                     padding=1 if halve_z_dim else (1, 1, 1),
                                     stride=2 if halve_z_dim else (2, 2, 1),
                                     normalization_type=normalization_type)
    return nn.Sequential(first, second)
