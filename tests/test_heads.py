"""Heads directly downstream of the fusion output: EmotionClassifier, the valence / arousal /
uncertainty heads, the class-probability softmaxes and the trainer's label-smoothed cross-entropy.
CPU: the oracle restatement against golden vectors produced by the EXECUTED reference (oracle/make_golden_heads.py) and
closed-form known answers.  GPU: the CUDA heads (through the C ABI) against those fixtures and the oracle: fp32 1e-5, bf16 2e-2."""
import glob
import importlib
import math
import os

import pytest
import torch

from oracle import fusion_oracle as fo

pkg = importlib.import_module("simple-multimodal_b200")
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "heads", "*.pt")))
TOL = dict(rtol=1e-9, atol=1e-11)


class Cfg:
    def __init__(self, H, E=7, p=0.0):
        self.fusion_hidden_size, self.num_emotions, self.fusion_dropout = H, E, p


def _oracle(rec):
    m = rec["meta"]
    P = {k: v.double().requires_grad_(True) for k, v in fo.init_head_params(m["H"], m["num_emotions"], m["param_seed"]).items()}
    f = rec["fused"].clone().requires_grad_(True)
    logits = fo.emotion_classifier(f, P)
    aux = fo.auxiliary_heads(f, P)
    probs = fo.softmax_lastdim(logits)
    ce = fo.cross_entropy_label_smoothing(logits, rec["target"], m["label_smoothing"])
    total = ce + 0.05 * ((aux["valence"] ** 2).mean() + (aux["arousal"] ** 2).mean() + (aux["uncertainty"] ** 2).mean() + (probs ** 2).mean())
    total.backward()
    return logits, probs, aux, ce, total, f.grad, {k: v.grad for k, v in P.items()}


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_oracle_heads_match_executed_reference(path):
    rec = torch.load(path, weights_only=True)
    logits, probs, aux, ce, total, df, pg = _oracle(rec)
    torch.testing.assert_close(logits.detach(), rec["logits"], **TOL)
    torch.testing.assert_close(probs.detach(), rec["probs"], **TOL)
    for k in ("valence", "arousal", "uncertainty"):
        torch.testing.assert_close(aux[k].detach(), rec[k], **TOL)
    torch.testing.assert_close(ce.detach(), rec["ce"], **TOL)
    torch.testing.assert_close(df, rec["dfused"], **TOL)
    for k, g in rec["param_grads"].items():
        torch.testing.assert_close(pg[k], g, **TOL)
    # the hierarchical auxiliary classifiers receive no gradient in the reference (their logits are discarded)
    assert not any(k.startswith(("sentiment", "positive", "negative")) for k in rec["param_grads"])


def test_cross_entropy_known_answers():
    C = 7
    uniform = torch.zeros(5, C, dtype=torch.float64)
    t = torch.tensor([0, 1, 2, 3, 6])
    assert abs(float(fo.cross_entropy_label_smoothing(uniform, t, 0.1)) - math.log(C)) < 1e-12        # any smoothing: log C
    sharp = torch.full((3, C), -50.0, dtype=torch.float64)
    sharp[torch.arange(3), torch.tensor([1, 4, 5])] = 50.0
    got = float(fo.cross_entropy_label_smoothing(sharp, torch.tensor([1, 4, 5]), 0.1))
    assert abs(got - 0.1 * 100.0 * (C - 1) / C) < 1e-9                                                  # only the smoothing term is left


def test_module_state_dict_matches_reference_names():
    rec = torch.load(GOLDEN[0], weights_only=True)
    clf = pkg.EmotionClassifier(Cfg(rec["meta"]["H"]))
    aux = pkg.AuxiliaryHeads(Cfg(rec["meta"]["H"]))
    P = fo.init_head_params(rec["meta"]["H"], 7, 1)
    clf.load_state_dict({k: v for k, v in P.items() if "classifier" in k}, strict=True)
    aux.load_state_dict({k: v for k, v in P.items() if "classifier" not in k}, strict=True)
    assert set(rec["param_grads"]) <= set(dict(clf.named_parameters())) | set(dict(aux.named_parameters()))
    with pytest.raises(pkg.B200FusionError):
        clf(torch.randn(2, rec["meta"]["H"]))                                                           # CPU tensors: no fallback


def _rel(x, r):
    x, r = x.detach().double().cpu(), r.detach().double().cpu()
    return float((x - r).norm() / r.norm().clamp_min(1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_cuda_heads_match_reference_fixture(path, dtype):
    rec = torch.load(path, weights_only=True)
    m = rec["meta"]
    P = fo.init_head_params(m["H"], m["num_emotions"], m["param_seed"])
    fused_h = rec["fused"].float()
    if dtype == torch.bfloat16:                       # identical data on both sides: make it bf16-representable, re-run the oracle on it
        P = {k: v.to(torch.bfloat16).float() for k, v in P.items()}
        fused_h = fused_h.to(torch.bfloat16).float()
        P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
        f64 = fused_h.double().requires_grad_(True)
        logits_r = fo.emotion_classifier(f64, P64)
        aux_r = fo.auxiliary_heads(f64, P64)
        probs_r = fo.softmax_lastdim(logits_r)
        ce_r = fo.cross_entropy_label_smoothing(logits_r, rec["target"], 0.1)
        (ce_r + 0.05 * ((aux_r["valence"] ** 2).mean() + (aux_r["arousal"] ** 2).mean() + (aux_r["uncertainty"] ** 2).mean() + (probs_r ** 2).mean())).backward()
        ref = dict(logits=logits_r, probs=probs_r, ce=ce_r, dfused=f64.grad, pg={k: v.grad for k, v in P64.items() if v.grad is not None}, **aux_r)
    else:
        ref = dict(logits=rec["logits"], probs=rec["probs"], ce=rec["ce"], dfused=rec["dfused"], pg=rec["param_grads"], valence=rec["valence"],
                   arousal=rec["arousal"], uncertainty=rec["uncertainty"])
    clf = pkg.EmotionClassifier(Cfg(m["H"])).cuda()
    aux = pkg.AuxiliaryHeads(Cfg(m["H"])).cuda()
    clf.load_state_dict({k: v for k, v in P.items() if "classifier" in k}, strict=True)
    aux.load_state_dict({k: v for k, v in P.items() if "classifier" not in k}, strict=True)
    clf.train(); aux.train()
    f = fused_h.cuda().to(dtype).requires_grad_(True)
    logits = clf(f)
    out = aux(f, logits)
    ce = pkg.SmoothedCrossEntropy(0.1)(logits, rec["target"].cuda())
    total = ce + 0.05 * ((out["valence"].float() ** 2).mean() + (out["arousal"].float() ** 2).mean() + (out["uncertainty"] ** 2).mean()
                         + (out["emotion_probs"] ** 2).mean())
    total.backward()
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert _rel(logits, ref["logits"]) < tol and _rel(out["emotion_probs"], ref["probs"]) < tol
    for k in ("valence", "arousal", "uncertainty"):
        assert _rel(out[k], ref[k]) < tol, k
    assert abs(float(ce.detach()) - float(ref["ce"].detach())) < (1e-5 if dtype == torch.float32 else 1e-3)
    assert _rel(f.grad, ref["dfused"]) < (1e-5 if dtype == torch.float32 else 3e-2)
    grads = {**{k: p.grad for k, p in clf.named_parameters()}, **{k: p.grad for k, p in aux.named_parameters()}}
    for k, g in ref["pg"].items():
        assert _rel(grads[k], g) < (1e-5 if dtype == torch.float32 else 3e-2), k


@pytest.mark.gpu
def test_cross_entropy_kernel_large_batch_and_known_answers():
    B, C = 4096, 7
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = (torch.randn(B, C, device="cuda", generator=g) * 3).requires_grad_(True)
    target = torch.randint(0, C, (B,), device="cuda", generator=g)
    loss = pkg.SmoothedCrossEntropy(0.1)(logits, target)
    loss.backward()
    ref_in = logits.detach().double().cpu().requires_grad_(True)
    ref = fo.cross_entropy_label_smoothing(ref_in, target.cpu(), 0.1)
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) < 1e-5 and _rel(logits.grad, ref_in.grad) < 1e-5
    assert abs(float(logits.grad.sum())) < 1e-5                                     # every row of the gradient sums to zero
    uni = pkg.SmoothedCrossEntropy(0.3)(torch.zeros(64, C, device="cuda"), target[:64])
    assert abs(float(uni) - math.log(C)) < 1e-6
