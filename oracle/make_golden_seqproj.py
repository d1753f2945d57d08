"""Golden vectors for the per-token projection extension from the EXECUTED reference encoders -- TEST
INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    cd /tmp && python /root/repo/oracle/make_golden_seqproj.py

The reference's own `models.encoders.{TextEncoder, AudioEncoder, VideoEncoder}` run unmodified (float64, eval mode), with one
substitution: their `from_pretrained` backbones (no network, encoders.py:20,116,179) are replaced by tiny deterministic stand-ins
that return a fixed random `last_hidden_state`.  Everything after the backbone -- temporal attention / LSTM / facial attention,
the CLS / masked-mean / mean pooling, `projection` -- is the reference's code.  Each record stores what the encoder RETURNS
(`sequence_output`, `attention_mask`, `features`) and its `projection` weights: the extension must reproduce `features` from
`sequence_output` (pooling of the per-token projected features), and its per-token output must equal projection(sequence_output).
Writes tests/golden/seqproj/*.pt."""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "seqproj")


class _Out:
    def __init__(self, h):
        self.last_hidden_state = h


class _Backbone(nn.Module):
    """stand-in for AutoModel / Wav2Vec2Model / ViTModel: config.hidden_size, config.model_type and a fixed output"""

    def __init__(self, hidden, model_type, table):
        super().__init__()
        self.config = types.SimpleNamespace(hidden_size=hidden, model_type=model_type)
        self.table = table            # [N, L, hidden] returned row-aligned with the batch

    def forward(self, *args, input_ids=None, attention_mask=None, pixel_values=None, **kw):
        n = (input_ids if input_ids is not None else pixel_values if pixel_values is not None else args[0]).size(0)
        return _Out(self.table[:n])


class Cfg:
    fusion_hidden_size = 64
    fusion_dropout = 0.0
    text_model_name = audio_model_name = video_model_name = "stub"


def main():
    sys.path.insert(0, "/root/reference")
    enc = importlib.import_module("models.encoders")
    os.makedirs(OUT, exist_ok=True)
    D, B = 128, 5
    g = torch.Generator().manual_seed(99)

    def rnd(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float64)

    cases = []
    # text, model_type containing 'bert' (deberta-v2 is the reference default, config.py) -> CLS pooling; otherwise masked mean
    for name, model_type, pooling in (("text_cls", "deberta-v2", "cls"), ("text_masked_mean", "electra", "masked_mean")):
        L = 12
        table = rnd(B, L, D)
        enc.AutoModel.from_pretrained = staticmethod(lambda _n, t=table, mt=model_type: _Backbone(D, mt, t))
        torch.manual_seed(3)
        m = enc.TextEncoder(Cfg()).double().eval()
        ids = torch.zeros(B, L, dtype=torch.long)
        am = torch.ones(B, L, dtype=torch.long)
        for i, n_valid in enumerate((12, 7, 1, 9, 0)):           # ragged lengths, incl. an all-masked row (the 1e-9 clamp)
            am[i, n_valid:] = 0
        out = m(ids, am)
        cases.append((name, pooling, out["sequence_output"], out["attention_mask"], out["features"], m.projection))
    # audio: wav2vec2 stand-in -> temporal attention -> mean over time
    L = 20
    table = rnd(B, L, D)
    enc.Wav2Vec2Model.from_pretrained = staticmethod(lambda _n, t=table: _Backbone(D, "wav2vec2", t))
    torch.manual_seed(4)
    m = enc.AudioEncoder(Cfg()).double().eval()
    out = m(torch.zeros(B, 100, dtype=torch.float64))
    cases.append(("audio_mean", "mean", out["sequence_output"], None, out["features"], m.projection))
    # video: ViT stand-in (CLS of every frame) -> 2-layer BiLSTM -> facial attention -> mean over frames
    F_ = 6
    table = rnd(B * F_, 3, D)
    enc.ViTModel.from_pretrained = staticmethod(lambda _n, t=table: _Backbone(D, "vit", t))
    torch.manual_seed(5)
    m = enc.VideoEncoder(Cfg()).double().eval()
    out = m(torch.zeros(B, F_, 3, 4, 4, dtype=torch.float64))
    cases.append(("video_mean", "mean", out["sequence_output"], None, out["features"], m.projection))

    for name, pooling, seq, am, feats, proj in cases:
        rec = {"meta": {"pooling": pooling, "B": B, "L": seq.size(1), "D": D, "H": Cfg.fusion_hidden_size,
                        "source": "reference models/encoders.py forward, backbone replaced by a fixed-output stand-in"},
               "sequence_output": seq.detach(), "attention_mask": am, "features": feats.detach(),
               "projection.weight": proj.weight.detach(), "projection.bias": proj.bias.detach()}
        torch.save(rec, os.path.join(OUT, name + ".pt"))
        print(name, tuple(seq.shape), tuple(feats.shape))


if __name__ == "__main__":
    main()
