"""CUDA-graph execution of the GAN training step (the reference's step is
GAN_models/wind_field_GAN_3D.py:570-619).

One G step of the upscale8 configuration is ~1 500 kernel launches (a D step ~250): enqueuing them one by one costs
the host about as long as the GPU needs to run them, so any hiccup of the host (GC, a slow core, a busy PCIe root)
lands in the step time.  After ``WARMUP_CALLS`` eager calls with a given (step kind, batch shape, precision) the whole
step — generator / discriminator forward, the fused wind loss, backward, gradient all-reduce, Adam — is captured
once with ``torch.cuda.graph`` and every later call is three device copies into the graph's static inputs, one small
asynchronous copy of the iteration-dependent scalars, and one ``cudaGraphLaunch``.

What makes the step capturable (all in this package):
* no host synchronisation inside the step (device-side NaN guards, ``WindAdam``'s ``found_inf``);
* label values, instance-noise scales, the learning rate: device scalars refreshed before a replay;
* RNG: torch's Philox offsets are graph-aware (dropout masks, label noise); the instance-noise kernel keeps its own
  call counter on the device;
* caches whose tensors could be freed later (packed weights, stencil coefficients, auxiliary workspaces) are
  bypassed during capture and their buffers are owned by the graph (``ops.keepalive``);
* the auxiliary-stream weight gradients fork from / join the capturing stream through events.
"""
from __future__ import annotations

import os
import sys

import torch

from .. import ops

WARMUP_CALLS = int(os.environ.get("WINDSR_GRAPH_WARMUP", "3"))
MAX_GRAPHS = 6


class StepGraph:
    def __init__(self, gan, kind, LR, HR, Z):
        self.kind = kind
        self.static = [torch.empty_like(t) for t in (LR, HR, Z)]
        for s, t in zip(self.static, (LR, HR, Z)):
            s.copy_(t)
        self.graph = torch.cuda.CUDAGraph()
        self.optimizer = gan.optimizer_G if kind == "G" else gan.optimizer_D
        self.optimizer.prepare_capture()
        ops.take_keepalive()
        launches0 = ops.launch_count()
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            gan._train_step_body(kind, *self.static)
        self.keep = ops.take_keepalive()
        self.launches = ops.launch_count() - launches0  # libwindsr kernels one replay launches
        self.optimizer.finish_capture()
        # references to what the captured step publishes (static tensors, refreshed by every replay)
        self.G_losses = dict(gan.train_G_loss_dict) if kind == "G" else None
        self.D_loss = gan.D_loss_dict["train_loss"] if kind == "D" else None
        self.replays = 0

    def replay(self, gan, LR, HR, Z):
        for s, t in zip(self.static, (LR, HR, Z)):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.optimizer.refresh_lr()
        self.graph.replay()
        self.replays += 1
        ops.invalidate_packed_weights()  # the replay moved the weights: eager code must repack
        ops.count_graph_launches(self.launches)
        if self.kind == "G":
            gan.train_G_loss_dict.update(self.G_losses)
        else:
            gan.D_loss_dict["train_loss"] = self.D_loss


def run_captured(gan, kind, LR, HR, Z) -> bool:
    """Replay the captured step for this call if there is one (capturing it when the warm-up count is reached).
    Returns False when the caller should run the step eagerly."""
    key = (kind, tuple(LR.shape), tuple(HR.shape), tuple(Z.shape), ops.get_precision(),
           bool(gan.cfg.training.use_instance_noise), gan._skip_D_in_G_step())
    g = gan._graphs.get(key)
    if g is None:
        n = gan._eager_calls.get(key, 0)
        if n < WARMUP_CALLS or len(gan._graphs) >= MAX_GRAPHS:
            gan._eager_calls[key] = n + 1
            return False
        try:
            g = StepGraph(gan, kind, LR, HR, Z)
        except Exception as exc:  # noqa: BLE001 - a failed capture must not take training down: stay eager
            gan._graphs[key] = False
            if torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
                raise  # the other ranks would wait for collectives this rank never replays
            print(f"[windsr] CUDA-graph capture of the {kind} step failed ({type(exc).__name__}: {exc}); "
                  "continuing eagerly", file=sys.stderr)
            torch.cuda.synchronize()
            return False
        gan._graphs[key] = g
        # the capture pass itself executes nothing: fall through and replay it for this call
    elif g is False:
        return False
    g.replay(gan, LR, HR, Z)
    return True
