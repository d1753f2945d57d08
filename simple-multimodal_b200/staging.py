"""Host -> device staging of per-step encoder features (plumbing around the fusion heads, torch streams only).

`FeaturePrefetcher` double-buffers the pinned-host -> HBM copy of the NEXT step's features on a side stream while the
current step's kernels run, the way a training loop feeds `MultimodalEmotionModel.forward` from a DataLoader with
`pin_memory=True` (reference training/advanced_trainer.py:118-131 moves each batch with `.to(device)` inside the step)."""
from __future__ import annotations

from typing import List, Sequence

import torch


class FeaturePrefetcher:
    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._pending = None

    def submit(self, host_tensors: Sequence[torch.Tensor]) -> None:
        """Start copying one step's (pinned) host tensors; returns immediately."""
        if self._pending is not None:
            raise RuntimeError("FeaturePrefetcher: previous submit() was not consumed by get()")
        with torch.cuda.stream(self.stream):
            dev = [h.to(self.device, non_blocking=True) for h in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (dev, ev)

    def get(self) -> List[torch.Tensor]:
        """Device tensors of the last submit(), ordered after the copy on the CURRENT stream."""
        if self._pending is None:
            raise RuntimeError("FeaturePrefetcher: get() without submit()")
        dev, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)            # the caching allocator must not recycle the block while `cur` still reads it
        return dev
