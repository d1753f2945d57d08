"""The full model (BASELINE config 5 at toy size) against golden vectors made by the EXECUTED reference
`MultimodalEmotionModel` (oracle/make_golden_model.py): same `state_dict()` keys and shapes (strict load of a reference checkpoint),
same output dictionary, outputs / loss / gradients within fp32 tolerance.  The encoders are stock PyTorch/HF on both sides; the
fusion head, classifier, auxiliary heads and probabilities run in libb200fusion.so."""
import glob
import importlib
import os

import pytest
import torch
import torch.nn as nn

GOLD = os.path.join(os.path.dirname(__file__), "golden", "model")
CASES = ["hierarchical", "mult", "late", "hierarchical_adapter_prompt"]


class Cfg:
    text_model_name = audio_model_name = video_model_name = "toy"
    fusion_hidden_size, fusion_dropout, fusion_num_heads, num_emotions = 32, 0.0, 8, 7
    graph_hidden_size, graph_num_layers, graph_dropout, contrastive_temperature = 32, 3, 0.0, 0.07
    adapter_size, prompt_length = 8, 3


def _load(case):
    return torch.load(os.path.join(GOLD, f"model_{case}.pt"), weights_only=False)


def _model(rec):
    em = importlib.import_module("simple-multimodal_b200.emotion_model")
    cfg = Cfg()
    cfg.fusion_type = rec["meta"]["fusion_type"]
    return em.MultimodalEmotionModel(cfg, em.build_backbones("tiny"))


def _flatten(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flatten(v, prefix + k + "."))
        elif torch.is_tensor(v):
            out[prefix + k] = v
    return out


def _rel(x, ref):
    x, ref = x.detach().double().cpu(), ref.detach().double().cpu()
    return float((x - ref).norm() / ref.norm().clamp_min(1e-30))


def test_golden_files_present():
    assert sorted(os.path.basename(p)[6:-3] for p in glob.glob(os.path.join(GOLD, "model_*.pt"))) == sorted(CASES)


@pytest.mark.parametrize("case", CASES[:3])
def test_state_dict_layout_matches_reference_checkpoint(case):
    rec = _load(case)
    model = _model(rec)
    ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in rec["state_dict"].items()}
    assert ours == ref
    model.load_state_dict(rec["state_dict"], strict=True)
    names = [n for n, _ in model.named_parameters()]
    assert len(names) == len(set(names))                                 # tied helpers (aux heads, sequence projector) add no duplicates


def test_fusion_refuses_cpu_inside_the_model():
    pkg = importlib.import_module("simple-multimodal_b200")
    rec = _load("late")
    model = _model(rec).eval()
    with pytest.raises(pkg.B200FusionError):
        model({"input_ids": rec["input_ids"], "attention_mask": rec["attention_mask"]}, rec["audio"], rec["video"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_full_model_matches_executed_reference(case):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rec = _load(case)
    model = _model(rec)
    model.load_state_dict(rec["state_dict"], strict=True)
    model = model.cuda().eval()
    model.video_encoder.temporal_lstm.train()        # cuDNN refuses an RNN backward in eval mode; its dropout is 0, so the numbers do not change
    kw = dict(rec["meta"]["forward_kwargs"])
    out = model({"input_ids": rec["input_ids"].cuda(), "attention_mask": rec["attention_mask"].cuda()}, rec["audio"].cuda(), rec["video"].cuda(), **kw)
    flat = _flatten(out)
    assert sorted(flat) == sorted(rec["outputs"])                        # same output dictionary, nested keys included
    # the encoders are stock fp32 torch on the GPU vs the float64 golden run: 1e-5 norm-wise holds for the whole pipeline at this depth
    worst = {k: _rel(flat[k], rec["outputs"][k]) for k in flat}
    assert max(worst.values()) < 1e-5, worst
    labels = rec["labels"].cuda()
    loss = nn.CrossEntropyLoss(label_smoothing=0.1)(out["emotion_logits"], labels)
    if out.get("contrastive_losses"):
        loss = loss + 0.1 * sum(out["contrastive_losses"].values())
    loss = loss + 0.05 * (out["valence"].pow(2).mean() + out["arousal"].pow(2).mean() + out["uncertainty"].pow(2).mean())
    assert abs(float(loss.detach()) - float(rec["loss"])) < 1e-5 * max(1.0, abs(float(rec["loss"])))
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    bad = {}
    for n, g_ref in rec["grads"].items():
        if float(g_ref.norm()) < 1e-12:
            continue
        assert n in grads, n
        r = _rel(grads[n], g_ref)
        # 2e-5 norm-wise, with an absolute floor of 1e-6 (the largest gradients here have norm ~2e-2): with a missing modality every
        # sample has the SAME audio embedding, the InfoNCE terms of that pair cancel analytically (logits scaled by 1/tau = 14.3) and
        # what is left of the ~2e-5 projector gradients carries the fp32 rounding of the cancelled terms (measured 2-4e-7 absolute;
        # the same head on its own, identical rows included, agrees with the oracle to 4e-7 relative)
        if r > 2e-5 and float((grads[n].detach().double().cpu() - g_ref.double()).norm()) > 1e-6:
            bad[n] = r
    assert not bad, bad


@pytest.mark.gpu
def test_full_model_with_sequences_trains():
    """use_sequences=True (SURVEY 8f rank 2 inside the model): MulT attends over the encoders' token / frame sequences; with an
    all-ones attention mask and mean-pooled heads the 2-D features equal the reference encoders' `features`."""
    pkg = importlib.import_module("simple-multimodal_b200")
    em = importlib.import_module("simple-multimodal_b200.emotion_model")
    rec = _load("hierarchical")
    cfg = Cfg()
    cfg.fusion_type = "hierarchical"
    model = em.MultimodalEmotionModel(cfg, em.build_backbones("tiny"), use_sequences=True)
    model.load_state_dict(rec["state_dict"], strict=True)
    model = model.cuda().eval()
    out = model({"input_ids": rec["input_ids"].cuda(), "attention_mask": rec["attention_mask"].cuda()}, rec["audio"].cuda(), rec["video"].cuda(),
                compute_contrastive_loss=True)
    for m in ("text", "audio", "video"):                                 # pooled features == what the reference encoders return
        assert _rel(out[f"{m}_features"], rec["outputs"][f"{m}_features"]) < 1e-5, m
    assert _rel(out["early_features"], rec["outputs"]["early_features"]) < 1e-5      # the 2-D heads see the same inputs as the reference
    assert _rel(out["mult_features"], rec["outputs"]["mult_features"]) > 1e-3        # MulT now sees sequences, not pooled vectors
    model.train()
    out = model({"input_ids": rec["input_ids"].cuda(), "attention_mask": rec["attention_mask"].cuda()}, rec["audio"].cuda(), rec["video"].cuda(),
                compute_contrastive_loss=True)
    loss = out["emotion_logits"].float().pow(2).mean() + 0.1 * sum(out["contrastive_losses"].values())
    loss.backward()
    opt = pkg.FusedAdamW([p for p in model.parameters() if p.grad is not None], lr=1e-3)
    opt.clip_grad_norm_(1.0)
    opt.step()
    assert all(torch.isfinite(p).all() for p in model.parameters())
