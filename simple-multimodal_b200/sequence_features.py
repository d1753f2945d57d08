"""Per-token projection of the encoders' sequence outputs -- an EXTENSION, not part of the drop-in contract.

The reference encoders already return `sequence_output` [B,L,D] next to the pooled `features` [B,H]
(models/encoders.py:100-104, 163-167, 247-251), but `MultimodalEmotionModel.forward` hands only the pooled vectors to the fusion
head (models/multimodal_model.py:96-103), so MulT never sees real sequences.  `SequenceProjector` applies each
encoder's own `projection` Linear (encoders.py:35,134,207 -- the SAME parameters, passed in and tied, nothing new to train or to
load) to every token, which gives MulT [B,L,H] sequences, and pools the projected tokens the way that encoder pools before
projecting (CLS token / attention-mask mean / plain mean, encoders.py:86-93,157,241), which gives back the encoder's `features`
for the 2-D heads -- projection and pooling commute because the pooling weights sum to one, so the pooled output equals the
reference's `features` (tests/test_seqproj.py checks it against the executed reference encoders).

Use inside a model's forward (the explicit change INTEGRATION.md describes):

    seq = self.sequence_projector(text_output, audio_output, video_output)          # the encoders' own output dicts
    fusion_output = self.fusion_layer(seq["text_sequence"], seq["audio_sequence"], seq["video_sequence"],
                                      pooled_features=(seq["text_features"], seq["audio_features"], seq["video_features"]),
                                      compute_contrastive_loss=compute_contrastive_loss)

All arithmetic runs in libb200fusion.so: the token GEMM is `b200f_gemm` (tcgen05 for bf16), the pooling `b200f_meanpool_*` /
`b200f_weighted_pool_*`, dropout the library's counter-based kernel.  No fallback."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import B200FusionError
from .fusion_layers import _FusionBase

Tensor = torch.Tensor
POOLINGS = ("cls", "masked_mean", "mean")


def text_pooling_of(backbone_config) -> str:
    """The pooling the reference text encoder applies for this backbone: CLS when 'bert' occurs in `model_type`
    (encoders.py:86 -- true for deberta-v2, the reference default), attention-mask mean otherwise."""
    mt = getattr(backbone_config, "model_type", None)
    return "cls" if isinstance(mt, str) and "bert" in mt else "masked_mean"


class SequenceProjector(_FusionBase):
    """SequenceProjector(config, text_projection, audio_projection, video_projection, text_pooling='cls')

    The three `nn.Linear(D_m, fusion_hidden_size)` are the encoders' `projection` modules; they are registered here as parameter
    containers under the SAME objects (tied), so optimizers and checkpoints see each parameter once, under its encoder name, if
    the model registers the encoders first."""

    def __init__(self, config, text_projection: nn.Linear, audio_projection: nn.Linear, video_projection: nn.Linear, text_pooling: str = "cls"):
        super().__init__()
        if text_pooling not in POOLINGS:
            raise ValueError(f"text_pooling must be one of {POOLINGS}")
        self.config = config
        self.text_projection, self.audio_projection, self.video_projection = text_projection, audio_projection, video_projection
        self.poolings = {"text": text_pooling, "audio": "mean", "video": "mean"}

    def _one(self, seq: Tensor, proj: nn.Linear, pooling: str, attention_mask: Optional[Tensor]):
        if seq.dim() != 3 or seq.size(-1) != proj.in_features:
            raise B200FusionError(f"sequence_output must be [B,L,{proj.in_features}], got {tuple(seq.shape)}")
        tokens = ops.linear(seq, proj.weight, proj.bias)                               # [B,L,H], one GEMM over all tokens
        if pooling == "cls":
            pooled = tokens[:, 0].contiguous()
        elif pooling == "mean":
            pooled = ops.MeanPoolFn.apply(tokens)
        else:
            if attention_mask is None:
                raise B200FusionError("masked_mean pooling needs the encoder's attention_mask")
            m = attention_mask.to(device=seq.device, dtype=torch.float32)
            w = (m / m.sum(dim=1, keepdim=True).clamp(min=1e-9)).contiguous()          # [B,L] plumbing; the pooling itself is a kernel
            pooled = ops.WeightedPoolFn.apply(tokens, w)
            # a row with no valid token: the reference projects the clamped-zero mean, i.e. returns the bias (encoders.py:92-96);
            # the weights of such a row sum to 0 instead of 1, so the missing share of the bias is added back ([B,H], tiny)
            slack = (1.0 - w.sum(dim=1, keepdim=True)).to(pooled.dtype)
            pooled = torch.addcmul(pooled, slack, proj.bias.to(pooled.dtype).unsqueeze(0))
        p = self._p
        return ops.dropout(tokens, p, self.training), ops.dropout(pooled, p, self.training)

    def forward(self, text_output: Dict[str, Tensor], audio_output: Dict[str, Tensor], video_output: Dict[str, Tensor]) -> Dict[str, Tensor]:
        seqs = (text_output["sequence_output"], audio_output["sequence_output"], video_output["sequence_output"])
        seqs, _, _ = self._prepare(seqs, None)
        out = {}
        for name, seq, proj, enc_out in zip(("text", "audio", "video"), seqs, (self.text_projection, self.audio_projection, self.video_projection),
                                            (text_output, audio_output, video_output)):
            tokens, pooled = self._one(seq, proj, self.poolings[name], enc_out.get("attention_mask"))
            out[f"{name}_sequence"], out[f"{name}_features"] = tokens, pooled
        return out
