"""Fused optimizer step for the fusion parameters.

`FusedAdamW` is a `torch.optim.Optimizer` with AdamW's hyper-parameters and param-group semantics (the reference builds
`AdamW([{'params': pretrained, 'lr': 0.1*lr}, {'params': new, 'lr': lr}], weight_decay=...)`, training/advanced_trainer.py:91-94,
and drives it with OneCycleLR, :104-110 -- schedulers work unchanged because they only touch `param_groups[i]['lr']`).
`clip_grad_norm_(max_norm)` replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)` (:174-180): the global norm and
the clip coefficient are computed on the device and consumed by the next `step()`; nothing is read back to the host.
Per group, one kernel applies the update to every tensor through a device-side chunk table that is rebuilt only when the set of
(parameter, gradient) addresses changes (it never does under `GraphedTrainStep`, whose gradient buffers are static)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from ._lib import B200FusionError, check, lib, ptr, stream_ptr

CHUNK = 4096


class _Table:
    """device-side chunk table of one param group (see b200f_adamw_step in include/b200_fusion.h)"""

    def __init__(self, params, grads, ms, vs, device):
        T = len(params)
        ptrs = [t.data_ptr() for t in params] + [t.data_ptr() for t in grads] + [t.data_ptr() for t in ms] + [t.data_ptr() for t in vs]
        numel = [t.numel() for t in params]
        chunks = []
        for i, n in enumerate(numel):
            if n >= 2 ** 31:
                raise B200FusionError("FusedAdamW: tensors of 2^31 elements or more are not supported")
            chunks += [(i, s) for s in range(0, n, CHUNK)]
        head = torch.tensor(ptrs + numel, dtype=torch.int64)
        tail = torch.tensor(chunks, dtype=torch.int32).reshape(-1)
        host = torch.cat([head.view(torch.uint8), tail.view(torch.uint8)])
        if torch.cuda.is_available():
            host = host.pin_memory()                     # staged from pinned memory: the copy below does not block the host
        self.buf = host.to(device, non_blocking=True)    # one small H2D copy per (re)build (stream-ordered before the kernels)
        self._host = host                                # keeps the pinned staging buffer alive until the copy has run
        self.n_tensors, self.n_chunks = T, len(chunks)
        self.key = tuple(ptrs)


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}
        self._clip = None          # device tensor [sum of squares, total norm, clip coefficient] of the pending clip, or None

    def _group_tensors(self, gi, group):
        ps = [p for p in group["params"] if p.grad is not None]
        for p in ps:
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_cuda:
                raise B200FusionError("FusedAdamW: parameters and gradients must be CUDA float32 (there is no CPU fallback)")
            if not p.is_contiguous() or not p.grad.is_contiguous():
                raise B200FusionError("FusedAdamW: parameters and gradients must be contiguous")
            st = self.state[p]
            if st and torch.is_tensor(st.get("step")):          # an optimizer_state_dict saved by torch.optim.AdamW (the reference
                st["step"] = int(st["step"].item())             # trainer checkpoints it) keeps `step` as a tensor: normalise once
            if not st:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return ps

    def _table(self, gi, ps) -> Optional[_Table]:
        if not ps:
            return None
        key = tuple([p.data_ptr() for p in ps] + [p.grad.data_ptr() for p in ps] + [self.state[p]["exp_avg"].data_ptr() for p in ps]
                    + [self.state[p]["exp_avg_sq"].data_ptr() for p in ps])
        tab = self._tables.get(gi)
        if tab is None or tab.key != key:
            tab = _Table(ps, [p.grad for p in ps], [self.state[p]["exp_avg"] for p in ps], [self.state[p]["exp_avg_sq"] for p in ps], ps[0].device)
            self._tables[gi] = tab
        return tab

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """Global L2 norm over the gradients of ALL param groups; the next step() scales the gradients by min(1, max_norm / (norm + 1e-6)).
        Returns the total norm as a 0-dim device tensor (no synchronisation)."""
        all_ps: List[torch.Tensor] = []
        for gi, group in enumerate(self.param_groups):
            all_ps += self._group_tensors(gi, group)
        if not all_ps:                                 # no gradients anywhere: norm 0 on the parameters' device, nothing to launch
            dev = next((p.device for g in self.param_groups for p in g["params"]), torch.device("cpu"))
            return torch.zeros((), device=dev)
        tab = self._table("clip", all_ps)
        scratch = torch.empty(3, device=all_ps[0].device, dtype=torch.float32)
        check(lib().b200f_grad_clip_coef(ptr(tab.buf), C.c_int32(tab.n_tensors), C.c_int32(tab.n_chunks), C.c_float(max_norm), ptr(scratch), stream_ptr()),
              "b200f_grad_clip_coef")
        self._clip = scratch
        return scratch[1]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        coef = None if self._clip is None else self._clip[2:]
        for gi, group in enumerate(self.param_groups):
            all_ps = self._group_tensors(gi, group)
            b1, b2 = group["betas"]
            # bias correction depends on the per-parameter step count (a parameter that had no gradient in some iteration lags
            # behind, as in torch): one launch per distinct count -- normally exactly one
            parts = {}
            for p in all_ps:
                parts.setdefault(self.state[p]["step"], []).append(p)
            for si, s_prev in enumerate(sorted(parts)):
                ps = parts[s_prev]
                tab = self._table((gi, si), ps)
                for p in ps:
                    self.state[p]["step"] = s_prev + 1
                    # the kernel writes the parameter through its raw pointer: tell autograd / the operand caches keyed on the
                    # tensor version (mult_engine._Weights) that its contents changed
                    torch.autograd.graph.increment_version(p)
                check(lib().b200f_adamw_step(ptr(tab.buf), C.c_int32(tab.n_tensors), C.c_int32(tab.n_chunks), ptr(coef), C.c_double(group["lr"]),
                                             C.c_double(b1), C.c_double(b2), C.c_double(group["eps"]), C.c_double(group["weight_decay"]),
                                             C.c_int64(s_prev + 1), stream_ptr()), "b200f_adamw_step")
        self._clip = None
        return loss
