"""Heads directly downstream of the fusion output on the same C-ABI kernels.

* `EmotionClassifier`   -- drop-in for reference models/multimodal_model.py:186-219 (same constructor, parameter names and
                           shapes; forward(features) -> main logits; the hierarchical auxiliary classifiers are kept as
                           parameters, and like the reference their outputs are not returned).
* `AuxiliaryHeads`      -- the valence / arousal regressors and the uncertainty head that MultimodalEmotionModel holds as
                           three nn.Linear (multimodal_model.py:55-60) evaluated as ONE packed GEMM over the fused features,
                           plus the class-probability softmaxes of multimodal_model.py:160-164.
* `SmoothedCrossEntropy`-- drop-in for the trainer's nn.CrossEntropyLoss(label_smoothing=0.1) (training/advanced_trainer.py:53,
                           139): one kernel forward (softmax, loss, probabilities saved), one backward.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch
import torch.nn as nn

from . import kernels as K
from . import ops
from ._lib import B200FusionError, check, dtype_code, lib, ptr, stream_ptr
from .fusion_layers import _FusionBase, _mlp_container

Tensor = torch.Tensor


class RowSoftmaxFn(torch.autograd.Function):
    """probs [B,C] fp32 = softmax(logits [B,C]) for small C (thread per sample)."""

    @staticmethod
    def forward(ctx, x):
        x = x if x.stride(-1) == 1 else x.contiguous()
        B, Cn = x.shape
        p = torch.empty((B, Cn), device=x.device, dtype=torch.float32)
        check(lib().b200f_row_softmax_fwd(ptr(x), C.c_int64(x.stride(0)), ptr(p), C.c_int64(B), C.c_int32(Cn), dtype_code(x.dtype), stream_ptr()),
              "b200f_row_softmax_fwd")
        ctx.save_for_backward(p)
        ctx.dt = x.dtype
        return p

    @staticmethod
    def backward(ctx, dp):
        (p,) = ctx.saved_tensors
        B, Cn = p.shape
        dx = torch.empty((B, Cn), device=p.device, dtype=ctx.dt)
        check(lib().b200f_row_softmax_bwd(ptr(p), ptr(dp.float().contiguous()), ptr(dx), C.c_int64(Cn), C.c_int64(B), C.c_int32(Cn),
                                          dtype_code(ctx.dt), stream_ptr()), "b200f_row_softmax_bwd")
        return dx


class SmoothedCrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, eps):
        logits = logits if logits.stride(-1) == 1 else logits.contiguous()
        if target.dtype != torch.int64 or not target.is_cuda:
            raise B200FusionError("cross_entropy: targets must be a CUDA int64 tensor of class indices")
        B, Cn = logits.shape
        probs = torch.empty((B, Cn), device=logits.device, dtype=torch.float32)
        acc = torch.zeros(2, device=logits.device, dtype=torch.float32)           # [loss sum, bad-target flag (as int32 bits)]
        flag = acc[1:].view(torch.int32)
        check(lib().b200f_ce_ls_fwd(ptr(logits), C.c_int64(logits.stride(0)), ptr(target.contiguous()), C.c_float(eps), ptr(probs), ptr(acc),
                                    ptr(flag), C.c_int64(B), C.c_int32(Cn), dtype_code(logits.dtype), stream_ptr()), "b200f_ce_ls_fwd")
        ctx.save_for_backward(probs, target)
        ctx.cfg = (eps, logits.dtype)
        ctx.flag = flag                     # checked lazily (no host sync on the training path); see SmoothedCrossEntropy.check_targets
        return acc[0] / B

    @staticmethod
    def backward(ctx, g):
        probs, target = ctx.saved_tensors
        eps, dt = ctx.cfg
        B, Cn = probs.shape
        dx = torch.empty((B, Cn), device=probs.device, dtype=dt)
        check(lib().b200f_ce_ls_bwd(ptr(probs), ptr(target), C.c_float(eps), ptr(g.float().reshape(1).contiguous()), ptr(dx), C.c_int64(Cn),
                                    C.c_int64(B), C.c_int32(Cn), dtype_code(dt), stream_ptr()), "b200f_ce_ls_bwd")
        return dx, None, None


class SmoothedCrossEntropy(nn.Module):
    """nn.CrossEntropyLoss(label_smoothing=eps), mean reduction, class-index targets (advanced_trainer.py:53,139)."""

    def __init__(self, label_smoothing: float = 0.1):
        super().__init__()
        self.label_smoothing = float(label_smoothing)

    def forward(self, logits: Tensor, target: Tensor) -> Tensor:
        if not logits.is_cuda:
            raise B200FusionError("b200 heads run on CUDA tensors only (there is no CPU fallback)")
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()
        return SmoothedCrossEntropyFn.apply(logits, target, self.label_smoothing)


class EmotionClassifier(_FusionBase):
    """reference models/multimodal_model.py:186-219."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H = config.fusion_hidden_size
        self.classifier = _mlp_container([H, H // 2, config.num_emotions], config.fusion_dropout, final_relu=False)
        self.sentiment_classifier = nn.Linear(H, 3)       # hierarchical auxiliary classifiers: parameters only, exactly like the
        self.positive_classifier = nn.Linear(H, 2)        # reference, whose forward computes and discards their logits
        self.negative_classifier = nn.Linear(H, 4)

    def forward(self, features: Tensor) -> Tensor:
        (x,), _, _ = self._prepare((features,), None)
        l0, l3 = self.classifier[0], self.classifier[3]
        h = ops.dropout(ops.linear(x, l0.weight, l0.bias, relu=True), self._p, self.training)
        return ops.linear(h, l3.weight, l3.bias)


class AuxiliaryHeads(_FusionBase):
    """valence_regressor / arousal_regressor / uncertainty_head of MultimodalEmotionModel (multimodal_model.py:55-60) and the
    probability outputs of its forward (:146-164).  The three nn.Linear keep their names and shapes; they are evaluated as one
    GEMM over a stacked [1 + 1 + E, H] weight."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H = config.fusion_hidden_size
        self.valence_regressor = nn.Linear(H, 1)
        self.arousal_regressor = nn.Linear(H, 1)
        self.uncertainty_head = nn.Linear(H, config.num_emotions)

    def forward(self, fused_features: Tensor, emotion_logits: Tensor = None) -> Dict[str, Tensor]:
        (x,), _, _ = self._prepare((fused_features,), None)
        w = torch.cat([self.valence_regressor.weight, self.arousal_regressor.weight, self.uncertainty_head.weight], 0)   # parameter packing
        b = torch.cat([self.valence_regressor.bias, self.arousal_regressor.bias, self.uncertainty_head.bias], 0)
        y = ops.linear(x, w, b)                                                       # [B, 2 + E]
        out = {"valence": y[:, 0:1], "arousal": y[:, 1:2], "uncertainty": RowSoftmaxFn.apply(y[:, 2:])}
        if emotion_logits is not None:
            out["emotion_probs"] = RowSoftmaxFn.apply(emotion_logits)
        return out
