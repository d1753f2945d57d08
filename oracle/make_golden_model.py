"""Golden vectors of the FULL model (SURVEY 8f rank 3 / BASELINE config 5 at toy size) from the EXECUTED reference -- TEST
INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    cd /tmp && python /root/repo/oracle/make_golden_model.py

The reference's own `models.multimodal_model.MultimodalEmotionModel` (encoders, modality dropout, fusion head, classifier,
auxiliary heads -- all its code) is constructed and run in float64, eval mode, with two substitutions forced by the sandbox:
`from_pretrained` (no network; models/encoders.py:20,116,179) returns random-init backbones of the same architectures at toy
size (deberta-v2 / wav2vec2 / ViT from their configs), and `torch_geometric` is the stand-in of oracle/ref_shim.py (GraphFusion
stays parity-unpinned, as everywhere).  Each record holds the model's `state_dict()` (fp32-representable values), the inputs, every
tensor of the output dictionary, a scalar training loss and the gradients of all non-backbone parameters.
Writes tests/golden/model/model_<fusion_type>.pt."""
import importlib
import os
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.make_golden_heads import load_reference_model_module          # noqa: E402  (installs the torch_geometric stand-in)

OUT = os.path.join(ROOT, "tests", "golden", "model")


class Cfg:
    text_model_name = audio_model_name = video_model_name = "toy"
    fusion_hidden_size = 32
    fusion_dropout = 0.0
    fusion_num_heads = 8
    num_emotions = 7
    graph_hidden_size = 32
    graph_num_layers = 3
    graph_dropout = 0.0
    contrastive_temperature = 0.07
    adapter_size = 8
    prompt_length = 3


def flatten(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(flatten(v, prefix + k + "."))
        elif torch.is_tensor(v):
            out[prefix + k] = v
    return out


def training_loss(out, labels):
    """the trainer's objective (training/advanced_trainer.py:139-166): label-smoothed CE + 0.1 * sum of contrastive losses, plus small
    quadratic terms so that every auxiliary head receives a gradient"""
    loss = nn.CrossEntropyLoss(label_smoothing=0.1)(out["emotion_logits"], labels)
    if out.get("contrastive_losses"):
        loss = loss + 0.1 * sum(out["contrastive_losses"].values())
    return loss + 0.05 * (out["valence"].pow(2).mean() + out["arousal"].pow(2).mean() + out["uncertainty"].pow(2).mean())


def main():
    mm = load_reference_model_module()
    enc = importlib.import_module("models.encoders")
    em = importlib.import_module("simple-multimodal_b200.emotion_model")     # only for build_backbones('tiny'): same toy architectures
    os.makedirs(OUT, exist_ok=True)
    B = 3
    g = torch.Generator().manual_seed(77)
    ids = torch.randint(0, 64, (B, 9), generator=g)
    am = torch.ones(B, 9, dtype=torch.long)
    am[1, 6:] = 0
    am[2, 3:] = 0
    audio = torch.randn(B, 400, generator=g).double()                      # fp32-representable inputs
    video = torch.randn(B, 4, 3, 16, 16, generator=g).double()
    labels = torch.randint(0, Cfg.num_emotions, (B,), generator=g)
    for fusion_type, kw in (("hierarchical", {"compute_contrastive_loss": True}), ("mult", {}), ("late", {}),
                            ("hierarchical_adapter_prompt", {"compute_contrastive_loss": True, "use_adapter": True, "use_prompt": True,
                                                             "missing_modalities": ["audio"]})):
        torch.manual_seed(21)
        bb = em.build_backbones("tiny")
        enc.AutoModel.from_pretrained = staticmethod(lambda _n: bb["text"])
        enc.Wav2Vec2Model.from_pretrained = staticmethod(lambda _n: bb["audio"])
        enc.ViTModel.from_pretrained = staticmethod(lambda _n: bb["video"])
        cfg = Cfg()
        cfg.fusion_type = fusion_type.split("_")[0]
        model = mm.MultimodalEmotionModel(cfg)
        with torch.no_grad():                       # adapters start near zero (std 0.02): make their contribution visible
            for n, p in model.named_parameters():
                if "adapter" in n and n.endswith("weight"):
                    p.mul_(10.0)
        state = {k: v.detach().clone().float() for k, v in model.state_dict().items()}
        model = model.double().eval()
        out = model({"input_ids": ids, "attention_mask": am}, audio, video, **kw)
        loss = training_loss(out, labels)
        loss.backward()
        grads = {n: p.grad.float() for n, p in model.named_parameters()
                 if p.grad is not None and not n.startswith(("text_encoder.model.", "audio_encoder.model.", "video_encoder.vit."))}
        rec = {"meta": {"fusion_type": cfg.fusion_type, "forward_kwargs": kw, "B": B, "mode": "eval", "dtype": "float64 run on fp32-representable weights"},
               "state_dict": state, "input_ids": ids, "attention_mask": am, "audio": audio.float(), "video": video.float(), "labels": labels,
               "outputs": {k: v.detach() for k, v in flatten(out).items()}, "loss": loss.detach(), "grads": grads}
        torch.save(rec, os.path.join(OUT, f"model_{fusion_type}.pt"))
        print(fusion_type, len(state), "state keys;", len(rec["outputs"]), "outputs; loss", float(loss.detach()))


if __name__ == "__main__":
    main()
