"""Checkpoint / device plumbing shared by GAN models (reference: GAN_models/baseGAN.py:19-106).

File names and contents are the reference's: ``G_{it}.pth`` / ``D_{it}.pth`` hold ``state_dict()``s (whose keys
and shapes equal the reference modules'), ``state_{it}.pth`` holds
``{"it", "epoch", "schedulers": [...], "optimizers": [...]}``.  Under data-parallel training only rank 0 writes.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from ..tools import loggingclass as lc


def _is_path(p) -> bool:
    return p is not None and str(p).lower() not in ("null", "none")


class BaseGAN(lc.GlobalLoggingClass):
    G: nn.Module = None
    D: nn.Module = None

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        explicit = getattr(cfg, "device", None)
        if isinstance(explicit, torch.device):
            self.device = explicit
        elif torch.cuda.is_available() and cfg.gpu_id is not None:
            self.device = torch.device(f"cuda:{cfg.gpu_id}")
        else:
            self.device = torch.device("cpu")
        self.is_train = cfg.is_train
        self.schedulers = []
        self.optimizers = []

    def load_model(self, generator_load_path=None, discriminator_load_path=None, state_load_path=None):
        if _is_path(generator_load_path):
            self.G.load_state_dict(torch.load(generator_load_path, map_location="cpu"))
            self.G.eval()
        if _is_path(discriminator_load_path):
            self.D.load_state_dict(torch.load(discriminator_load_path, map_location="cpu"))
            self.G.eval()
        self._after_load()
        if not _is_path(state_load_path):
            return None, None
        state = torch.load(state_load_path, map_location=self.device)
        opts, scheds = state["optimizers"], state["schedulers"]
        assert len(opts) == len(self.optimizers), \
            f"Loaded {len(opts)} optimizers but expected {len(self.optimizers)}"
        assert len(scheds) == len(self.schedulers), \
            f"Loaded {len(scheds)} schedulers but expected {len(self.schedulers)}"
        for mine, saved in zip(self.optimizers, opts):
            mine.load_state_dict(saved)
        for mine, saved in zip(self.schedulers, scheds):
            mine.load_state_dict(saved)
        self._after_load()
        return state["epoch"], state["it"]

    def _after_load(self):
        """Loaded weights / optimiser state invalidate everything derived from the old ones: captured CUDA graphs
        (they hold pointers to the previous optimiser state), packed operand copies; under data parallelism every
        rank takes rank 0's weights and buffers (only rank 0 is guaranteed to have read the files)."""
        from .. import ops
        from ..parallel import broadcast_module
        for attr in ("_graphs", "_eager_calls"):
            if hasattr(self, attr):
                getattr(self, attr).clear()
        ops.invalidate_packed_weights()
        if getattr(self, "world_size", 1) > 1:
            for m in (self.G, self.D):
                if m is not None:
                    broadcast_module(m)

    def save_model(self, save_basepath, epoch, it, save_G=True, save_D=True, save_state=True):
        if getattr(self, "rank", 0) != 0:
            return
        folder = getattr(self.cfg.env, "this_runs_folder", None) or save_basepath
        os.makedirs(folder, exist_ok=True)
        if save_G:
            torch.save(self.G.state_dict(), os.path.join(folder, f"G_{it}.pth"))
        if save_D and self.D is not None:
            torch.save(self.D.state_dict(), os.path.join(folder, f"D_{it}.pth"))
        if save_state:
            torch.save({"it": it, "epoch": epoch,
                        "schedulers": [s.state_dict() for s in self.schedulers],
                        "optimizers": [o.state_dict() for o in self.optimizers]},
                       os.path.join(folder, f"state_{it}.pth"))
