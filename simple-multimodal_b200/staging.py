"""Host -> device staging of per-step encoder features (plumbing around the fusion heads, torch streams only).

`FeaturePrefetcher` double-buffers the pinned-host -> HBM copy of the NEXT step's features on a side stream while the
current step's kernels run, the way a training loop feeds `MultimodalEmotionModel.forward` from a DataLoader with
`pin_memory=True` (reference training/advanced_trainer.py:118-131 moves each batch with `.to(device)` inside the step)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def bind_host_to_gpu(device) -> List[int]:
    """Pin this PROCESS to the CPU cores next to `device` (NVML's CPU affinity of the GPU), so that pinned staging buffers allocated
    afterwards are first-touched on the GPU's own NUMA node and the per-step H2D copies do not cross the socket interconnect.  One
    process per GPU on an 8-GPU box otherwise floats over both sockets and the eight concurrent 4.4 GB/step feature copies of the
    benchmark become the step's bottleneck.  Returns the cores bound to ([] = left alone: no NVML, no affinity data, not Linux)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(torch.device(device)).uuid)
        uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except TypeError:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:                                # plumbing only: never fail a run over an affinity hint
        return []


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
