"""Batch-sharded data parallelism for the GAN step: one process per GPU, gradients only.

The reference has no distributed code; the path shards by sample (independent through every conv
of G and D), so the only exchange is the gradient average of the network being updated: 140.8 MB (G) or
49.2 MB (D) fp32 per step.  ``GradSync`` packs gradients into ~25 MB flat buckets in reverse-registration order
(the order autograd produces them: ``hr_convs`` / the last UpConv first) and launches one asynchronous
``all_reduce`` per bucket as soon as its last gradient has been accumulated, so the NVLink/NVSwitch transfer
overlaps the remaining backward kernels; ``finish`` waits and re-points every ``param.grad`` at its slice of the
reduced bucket (no copy back).  On NCCL the reduction is ``ReduceOp.AVG`` (the division by the world size happens
inside the collective); on gloo (CPU tests) it is SUM followed by one in-place scale per bucket.

``broadcast_module`` / ``allreduce_max_`` are the two other exchanges a replica needs: identical initial weights
and buffers on every rank, and a GLOBAL "loss is not finite" flag so that either every rank skips the optimiser
step or none does (the gradients have already been averaged when the flag is consumed).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class _Bucket:
    __slots__ = ("params", "offsets", "numel", "flat", "pending", "work")

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += p.numel()
        self.numel = off
        self.flat = None
        self.pending = 0
        self.work = None


class GradSync:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 25 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.avg_in_collective = dist.is_initialized() and dist.get_backend(group) == "nccl"
        params = [p for p in params]
        self.buckets: List[_Bucket] = []
        cur, cur_bytes = [], 0
        for p in reversed(params):  # gradients become ready roughly in reverse registration order
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self.buckets.append(_Bucket(cur))
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(_Bucket(cur))
        self._where = {}
        for b in self.buckets:
            for p in b.params:
                self._where[p] = b
        self._active = False
        self._params = params
        self._hooks = {}

    def _ensure_hooks(self):
        # hooks can only be attached while a parameter requires grad (D's are frozen during G steps)
        for p in self._params:
            if p.requires_grad and p not in self._hooks:
                self._hooks[p] = p.register_post_accumulate_grad_hook(self._on_grad)

    def begin(self):
        """Arm the hooks for one backward pass."""
        self._active = self.world > 1
        self._ensure_hooks()
        for b in self.buckets:
            b.pending = sum(1 for p in b.params if p.requires_grad)
            b.work = None

    def _on_grad(self, p: torch.nn.Parameter):
        if not self._active:
            return
        b = self._where[p]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def _launch(self, b: _Bucket):
        live = [p for p in b.params if p.grad is not None]
        if not live:
            return
        from . import ops
        ops.aux_join()  # weight gradients of the residual dense blocks are produced on an auxiliary stream
        dev = live[0].grad.device
        if b.flat is None or b.flat.device != dev:
            b.flat = torch.zeros(b.numel, dtype=torch.float32, device=dev)
        views = []
        for p, off in zip(b.params, b.offsets):
            v = b.flat[off:off + p.numel()]
            if p.grad is None:
                v.zero_()
            else:
                views.append((p, v))
        torch._foreach_copy_([v for _, v in views], [p.grad.reshape(-1) for p, _ in views])
        op = dist.ReduceOp.AVG if self.avg_in_collective else dist.ReduceOp.SUM
        b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def finish(self):
        """Wait for every bucket, write the averaged gradients back."""
        if not self._active:
            return
        self._active = False
        inv = 1.0 / self.world
        for b in self.buckets:
            if b.work is None:
                if b.pending > 0 and any(p.grad is not None for p in b.params):
                    self._launch(b)  # some parameters of the bucket got no gradient this step
                if b.work is None:
                    continue
            b.work.wait()
            if not self.avg_in_collective:
                b.flat.mul_(inv)
            for p, off in zip(b.params, b.offsets):
                if p.grad is not None:
                    p.grad = b.flat[off:off + p.numel()].view_as(p)  # no copy back: the bucket IS the gradient
            b.work = None

    def remove(self):
        for h in self._hooks.values():
            h.remove()
        self._hooks = {}


def broadcast_module(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Every rank takes rank ``src``'s parameters and buffers (BatchNorm running statistics included)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)


def allreduce_max_(flag: torch.Tensor, group=None) -> torch.Tensor:
    """In-place MAX over ranks of a small device tensor (stream-ordered; no host synchronisation on NCCL)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    return flag


class _AllReduceMaxFn(torch.autograd.Function):
    """y = max over ranks of x (elementwise).  Backward: the cotangent reaches the local x only where the local value
    IS the global maximum — the sub-gradient torch.max would give on the concatenated batch.  The loss that consumes y
    is the mean over ranks of per-rank losses, so after the gradient all-reduce(AVG) the winner has to carry the SUM of
    every rank's cotangent: all-reduce(SUM) of the cotangent first."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.detach().clone()
        dist.all_reduce(y, op=dist.ReduceOp.MAX, group=group)
        ctx.save_for_backward(x.detach() == y)
        ctx.group = group
        return y

    @staticmethod
    def backward(ctx, g):
        (winner,) = ctx.saved_tensors
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return torch.where(winner, g, torch.zeros_like(g)), None


def allreduce_max(x: torch.Tensor, group=None) -> torch.Tensor:
    """Differentiable MAX over ranks (identity without a process group)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    return _AllReduceMaxFn.apply(x, group)
