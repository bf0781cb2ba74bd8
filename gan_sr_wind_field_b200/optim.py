"""``WindAdam``: torch.optim.Adam with the update running as ONE hand-written multi-tensor kernel
(``ws_adam_step``, csrc/train_aux.cu).

Reference: ``torch.optim.Adam(params, lr, weight_decay, betas=(beta1, 0.999))`` built at
GAN_models/wind_field_GAN_3D.py:151-162 and stepped at :459 / :566 (the G step only when the loss is finite, :457).

* Same constructor, ``param_groups``, ``state`` (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter) and therefore
  the same ``state_dict()`` layout as torch's Adam: ``state_{it}.pth`` files of the reference load unchanged
  (GAN_models/baseGAN.py:62-80), whatever ``fused`` / ``foreach`` / ``capturable`` flags they were saved with.
* ``found_inf`` (device float, optional): non-zero turns the step into a no-op on the device — the reference's
  "skip the optimiser step when the loss is NaN/Inf" guard without a host synchronisation.
* The learning rate is read from a device scalar refreshed from ``param_groups[...]["lr"]`` before every step, so
  a CUDA-graph replay of the step follows the MultiStepLR schedule (wind_field_GAN_3D.py:163-174).
* CPU parameters: falls back to torch's own implementation (host logic and CPU tests); there is no CPU kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class WindAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.found_inf = None          # optional device float tensor (1 element)
        self.grad_scale = 1.0          # folded into the gradient read (e.g. 1/world_size after a SUM all-reduce)
        self._tables = {}              # group index -> (key, table_dev, chunks_dev, ntensors, nchunks, lr_dev)
        self._staging = {}
        self._pending = []             # (device table, host bytes) of tables referenced by a graph under capture
        self._capture_blobs = {}       # group index -> table buffer pre-allocated for the next capture

    # -- state ----------------------------------------------------------------------------------------------
    def _ensure_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        else:
            s = st["step"]
            if not torch.is_tensor(s):
                st["step"] = torch.tensor(float(s), dtype=torch.float32, device=p.device)
            elif s.device != p.device or s.dtype != torch.float32:
                st["step"] = s.to(device=p.device, dtype=torch.float32)
            for k in ("exp_avg", "exp_avg_sq"):
                if st[k].device != p.device or st[k].dtype != torch.float32 or not st[k].is_contiguous():
                    st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()
        return st

    def _table(self, gi, params):
        """Device pointer table of one param group, rebuilt only when a pointer changed (gradients are re-allocated
        every eager step; under CUDA-graph capture everything is static)."""
        grads = [p.grad for p in params]
        states = [self._ensure_state(p) for p in params]
        key = tuple(g.data_ptr() for g in grads) + tuple(p.data_ptr() for p in params) + \
            tuple(s["exp_avg"].data_ptr() for s in states) + tuple(s["step"].data_ptr() for s in states)
        hit = self._tables.get(gi)
        capturing = torch.cuda.is_current_stream_capturing()
        if hit is not None and hit[0] == key and not capturing:
            return hit
        dev = params[0].device
        chunk = _lib.load().ws_adam_chunk_elems()
        n = len(params)
        tab = np.zeros((n, 6), dtype=np.int64)
        chunks = []
        for i, (p, g, s) in enumerate(zip(params, grads, states)):
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise _lib.WindSRError("WindAdam: gradients must be contiguous fp32")
            tab[i] = (p.data_ptr(), g.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr(),
                      s["step"].data_ptr(), p.numel())
            nck = (p.numel() + chunk - 1) // chunk
            chunks.append(np.stack((np.full(nck, i, dtype=np.int32), np.arange(nck, dtype=np.int32)), 1))
        chunks = np.concatenate(chunks, 0)
        raw = np.concatenate((tab.view(np.uint8).reshape(-1), chunks.view(np.uint8).reshape(-1)))
        if capturing:
            # Nothing may allocate pinned memory or copy from the host inside a capture.  The captured kernel only
            # needs the ADDRESS of its table: it goes into a buffer allocated BEFORE the capture (prepare_capture —
            # memory from the graph's own pool may be shared with tensors that die earlier in the step, whose
            # kernels would overwrite an out-of-band upload on every replay) and is filled right after the capture
            # has ended (finish_capture), before the first replay.
            blob = self._capture_blobs.pop(gi, None)
            if blob is None or blob.numel() < raw.size or blob.device != dev:
                raise _lib.WindSRError("WindAdam.step inside a CUDA-graph capture needs prepare_capture() first")
            self._pending.append((blob, raw.copy()))
        else:
            # a FRESH pinned block per upload (torch's caching host allocator does not hand it out again before the
            # asynchronous copy below has executed).  A reused staging buffer was a race: eager steps re-allocate the
            # gradients, so the table is rebuilt every step, and when the stream lagged the host the next table was
            # written into the staging buffer before the previous copy had read it — the earlier step then updated
            # through the later step's pointers (seen as a 5e-5 mismatch of the Adam test inside the full suite).
            host = torch.from_numpy(raw).pin_memory()
            blob = torch.empty(raw.size, dtype=torch.uint8, device=dev)
            blob.copy_(host, non_blocking=True)
        lr_dev = hit[5] if hit is not None and hit[5].device == dev else \
            torch.zeros((), dtype=torch.float32, device=dev)
        hit = (key, blob, blob[n * 48:], n, int(chunks.shape[0]), lr_dev)
        if not capturing:
            self._tables[gi] = hit
        return hit

    def prepare_capture(self):
        """Call right before capturing a step into a CUDA graph: pre-allocates (outside the graph's memory pool) the
        pointer-table buffer each param group's captured kernel will read."""
        chunk = _lib.load().ws_adam_chunk_elems()
        for gi, group in enumerate(self.param_groups):
            params = list(group["params"])  # upper bound: D's parameters are frozen (requires_grad False) between D steps
            if not params or not params[0].is_cuda:
                continue
            nbytes = 48 * len(params) + 8 * sum((p.numel() + chunk - 1) // chunk for p in params)
            self._capture_blobs[gi] = torch.empty(nbytes, dtype=torch.uint8, device=params[0].device)

    def finish_capture(self):
        """Upload the pointer tables of the kernels captured since the last call (see _table)."""
        for blob, raw in self._pending:
            blob[:raw.size].copy_(torch.from_numpy(raw))
            self._staging[("captured", len(self._staging))] = blob  # owned for the lifetime of the optimizer
        self._pending = []

    # -- device-side learning rate ------------------------------------------------------------------------------
    def lr_tensor(self, gi=0):
        hit = self._tables.get(gi)
        return None if hit is None else hit[5]

    def refresh_lr(self):
        """Copy ``param_groups[i]["lr"]`` into the device scalars the kernels read (call before a graph replay)."""
        for gi, group in enumerate(self.param_groups):
            hit = self._tables.get(gi)
            if hit is not None:
                hit[5].fill_(float(group["lr"]))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = None
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            if not params[0].is_cuda:
                return self._cpu_step(loss)
            if group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("WindAdam: amsgrad / maximize are not used by the reference")
            lib = lib or _lib.load()
            key, blob, chunks, n, nck, lr_dev = self._table(gi, params)
            capturing = torch.cuda.is_current_stream_capturing()
            if not capturing:
                lr_dev.fill_(float(group["lr"]))
            b1, b2 = group["betas"]
            fi = self.found_inf
            _lib.check(lib.ws_adam_step(blob.data_ptr(), chunks.data_ptr(), n, nck, lr_dev.data_ptr(),
                                        float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                        float(group["weight_decay"]), float(self.grad_scale),
                                        fi.data_ptr() if fi is not None else None, _lib.stream_ptr()),
                       "ws_adam_step")
        return loss

    def _cpu_step(self, loss):
        # host-side tests only: torch's reference implementation, honouring found_inf on the host
        if self.found_inf is not None and bool(self.found_inf != 0):
            return loss
        for group in self.param_groups:
            group["fused"] = None
            group["foreach"] = None
            group["capturable"] = False
            for p in group["params"]:
                st = self.state.get(p)
                if st and torch.is_tensor(st.get("step")) and st["step"].dtype != torch.float32:
                    st["step"] = st["step"].float()
        torch.optim.Adam.step(self)
        return loss
