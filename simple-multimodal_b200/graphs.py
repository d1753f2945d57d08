"""CUDA-graph capture of a whole fusion-head training step (forward, loss, backward).

A MulT step at B=256 is ~200 kernel launches from this package plus ~100 small torch ops; issuing them from Python costs
~16 ms of host time per step -- close to the 22 ms the B200 needs to execute them, and more than that when several ranks
share a CPU-limited host.  `GraphedTrainStep` runs the step once under `torch.cuda.graph` (every kernel of the library is
launched on torch's current stream and allocates through torch, so it is capturable as is) and replays it with one launch.

Dropout: the seeds are kernel arguments and therefore frozen at capture; the captured step starts with
`kernels.dropout_epoch(1, add=True)`, a one-thread kernel advancing the device-side epoch every mask generator XORs into its
seed, so each replay draws fresh masks while forward and backward of one replay agree (csrc/common.cuh)."""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence

import torch

from . import _lib
from . import kernels as K


class GraphedTrainStep:
    """step = GraphedTrainStep(head, example_inputs, loss_fn, forward_kwargs); loss = step(*inputs)

    After a call, `head.parameters()` hold this step's gradients in `.grad` (static buffers, overwritten by the next call) and
    `step.input_grads` the gradients of the inputs; `step.outputs` is the head's output structure (static tensors)."""

    def __init__(self, head: torch.nn.Module, example_inputs: Sequence[torch.Tensor], loss_fn: Callable, forward_kwargs: Optional[Dict] = None,
                 warmup: int = 2, grad_bucket=None, modality_dropout=None):
        """`grad_bucket` (ops.GradBucket over the head's parameters): gradients accumulate into its flat buffer, which the
        captured step zeroes first -- the buffer can then be all-reduced in place after each replay.
        `modality_dropout` (fusion_layers.ModalityDropout): its keep-mask is sampled INSIDE the captured step and passed to the
        head as `mask=`; the device-side epoch re-draws it on every replay (a mask tensor passed through `forward_kwargs` is a
        static buffer instead: refresh it in place before a replay if it should change)."""
        self.head, self.loss_fn, self.kw = head, loss_fn, dict(forward_kwargs or {})
        self.bucket, self.md = grad_bucket, modality_dropout
        for m in head.modules():                                        # the whole step becomes one graph: per-chunk graphs inside the
            if hasattr(m, "release_graphs"):                            # head (mult_engine.ChunkGraphEngine) would only hold memory
                m.graph_chunks = False
                m.release_graphs()
        self.static_inputs = [x.detach().clone().requires_grad_(x.requires_grad) for x in example_inputs]
        # parameters that already stepped eagerly keep AccumulateGrad nodes tied to the default stream; capture runs on a side
        # stream on purpose, so torch's advisory about that mismatch does not apply here
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                                   # warm-up off the default stream, as torch's capture recipe asks
            for _ in range(warmup):
                self._clear_grads()
                self.loss_fn(self._forward()).backward()
        torch.cuda.current_stream().wait_stream(side)
        self._clear_grads()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            K.dropout_epoch(1, add=True)
            if self.bucket is not None:
                self.bucket.flat.zero_()
            self.outputs = self._forward()
            self.loss = self.loss_fn(self.outputs)
            self.loss.backward()
        self.kernel_launches = _lib.launch_count() - n0          # kernels of this library inside the graph = launched by every replay

    def _forward(self):
        kw = self.kw
        if self.md is not None:
            kw = dict(kw, mask=self.md.sample_mask(self.static_inputs[0].size(0), self.static_inputs[0].device))
        return self.head(*self.static_inputs, **kw)

    def _clear_grads(self):
        if self.bucket is not None:
            self.bucket.zero()
        else:
            for p in self.head.parameters():
                p.grad = None
        for x in self.static_inputs:
            x.grad = None

    @property
    def input_grads(self):
        return [x.grad for x in self.static_inputs]

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        if len(inputs) != len(self.static_inputs):
            raise ValueError("GraphedTrainStep: wrong number of inputs")
        with torch.no_grad():
            for s, x in zip(self.static_inputs, inputs):
                if x is not s:
                    s.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.loss
