"""Per-token projection extension: oracle restatement and the CUDA SequenceProjector against golden vectors made
by the EXECUTED reference encoders (oracle/make_golden_seqproj.py: TextEncoder / AudioEncoder / VideoEncoder of
models/encoders.py with stand-in backbones).  fp32 rtol 1e-5, bf16 rtol 2e-2 (norm-wise, BASELINE north_star)."""
import glob
import importlib
import os

import pytest
import torch

from oracle import fusion_oracle as fo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "seqproj")
NAMES = ["text_cls", "text_masked_mean", "audio_mean", "video_mean"]


def _load(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def _rel(x, ref):
    x, ref = x.detach().double().cpu(), ref.detach().double().cpu()
    return float((x - ref).norm() / ref.norm().clamp_min(1e-30))


def test_golden_files_present():
    assert sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLD, "*.pt"))) == sorted(NAMES)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_executed_reference_encoders(name):
    r = _load(name)
    out = fo.sequence_projection(r["sequence_output"], r["projection.weight"], r["projection.bias"], r["meta"]["pooling"], r["attention_mask"])
    assert _rel(out["features"], r["features"]) < 1e-12
    # pooling the projected tokens == projecting the pooled tokens, for every row whose pooling weights sum to one
    pooled_after = fo.encoder_pool(out["sequence_features"], r["meta"]["pooling"], r["attention_mask"])
    rows = torch.ones(r["features"].size(0), dtype=torch.bool) if r["attention_mask"] is None else r["attention_mask"].sum(1) > 0
    assert _rel(pooled_after[rows], r["features"][rows]) < 1e-12
    if r["attention_mask"] is not None and r["meta"]["pooling"] == "masked_mean" and (~rows).any():
        # an all-masked row: the reference returns projection(0) = bias
        assert _rel(r["features"][~rows], r["projection.bias"].expand_as(r["features"][~rows])) < 1e-12


def _projector(pkg, recs, dropout=0.0):
    import torch.nn as nn

    class C:
        fusion_hidden_size, fusion_dropout = recs[0]["meta"]["H"], dropout
    lin = []
    for r in recs:
        m = nn.Linear(r["meta"]["D"], r["meta"]["H"])
        m.load_state_dict({"weight": r["projection.weight"].float(), "bias": r["projection.bias"].float()})
        lin.append(m)
    return pkg.SequenceProjector(C, *lin, text_pooling=recs[0]["meta"]["pooling"]).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("text_case", ["text_cls", "text_masked_mean"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_sequence_projector_matches_reference_features(text_case, dtype):
    pkg = importlib.import_module("simple-multimodal_b200")
    recs = [_load(text_case), _load("audio_mean"), _load("video_mean")]
    sp = _projector(pkg, recs).eval()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    ins = [{"sequence_output": r["sequence_output"].to("cuda", dtype).requires_grad_(True),
            "attention_mask": None if r["attention_mask"] is None else r["attention_mask"].cuda()} for r in recs]
    out = sp(*ins)
    loss = 0
    # oracle in float64 with autograd for the gradients
    P64 = [(r["projection.weight"].clone().requires_grad_(True), r["projection.bias"].clone().requires_grad_(True)) for r in recs]
    S64 = [r["sequence_output"].clone().requires_grad_(True) for r in recs]
    loss64 = 0
    for name, r, x, (w, b) in zip(("text", "audio", "video"), recs, S64, P64):
        ref = fo.sequence_projection(x, w, b, r["meta"]["pooling"], r["attention_mask"])
        assert _rel(out[f"{name}_features"], r["features"]) < tol, name            # == what the reference encoder returned
        assert _rel(out[f"{name}_features"], ref["features"]) < tol, name
        assert _rel(out[f"{name}_sequence"], ref["sequence_features"]) < tol, name
        loss = loss + out[f"{name}_features"].float().pow(2).mean() + out[f"{name}_sequence"].float().pow(2).mean()
        loss64 = loss64 + ref["features"].pow(2).mean() + ref["sequence_features"].pow(2).mean()
    loss.backward()
    loss64.backward()
    assert abs(float(loss.detach()) - float(loss64.detach())) < (1e-5 if dtype == torch.float32 else 1e-3) * max(1.0, abs(float(loss64)))
    for name, i, x, (w, b), lin in zip(("text", "audio", "video"), ins, S64, P64, (sp.text_projection, sp.audio_projection, sp.video_projection)):
        assert _rel(i["sequence_output"].grad, x.grad) < tol, name
        assert _rel(lin.weight.grad, w.grad) < tol, name
        assert _rel(lin.bias.grad, b.grad) < tol, name


@pytest.mark.gpu
def test_projected_sequences_feed_hierarchical_fusion():
    """encoders' sequence outputs -> SequenceProjector -> HierarchicalFusion(3-D for MulT, pooled_features for the 2-D heads) against
    the oracle composition, fp32."""
    pkg = importlib.import_module("simple-multimodal_b200")
    from parity_util import Cfg, rel
    recs = [_load("text_masked_mean"), _load("audio_mean"), _load("video_mean")]
    sp = _projector(pkg, recs).eval()
    cfg = Cfg(H=64, heads=8, graph_hidden=64, graph_layers=3)
    P = fo.init_params("hierarchical", H=64, heads=8, graph_hidden=64, graph_layers=3, seed=11)
    head = pkg.HierarchicalFusion(cfg).cuda()
    head.load_state_dict(P, strict=True)
    head.train()                                       # dropout p = 0
    ins = [{"sequence_output": r["sequence_output"].to("cuda", torch.float32).requires_grad_(True),
            "attention_mask": None if r["attention_mask"] is None else r["attention_mask"].cuda()} for r in recs]
    seq = sp(*ins)
    out = head(seq["text_sequence"], seq["audio_sequence"], seq["video_sequence"], compute_contrastive_loss=True,
               pooled_features=(seq["text_features"], seq["audio_features"], seq["video_features"]))
    fo.objective(out).backward()
    # oracle
    P64 = {k: v.double() for k, v in P.items()}
    S64 = [r["sequence_output"].clone().requires_grad_(True) for r in recs]
    W64 = [(r["projection.weight"].clone().requires_grad_(True), r["projection.bias"].clone().requires_grad_(True)) for r in recs]
    refs = [fo.sequence_projection(x, w, b, r["meta"]["pooling"], r["attention_mask"]) for r, x, (w, b) in zip(recs, S64, W64)]
    ref = fo.hierarchical_fusion(*[q["sequence_features"] for q in refs], P64, heads=8, graph_layers=3, temperature=cfg.contrastive_temperature,
                                 compute_contrastive_loss=True, pooled=tuple(q["features"] for q in refs))
    fo.objective(ref).backward()
    assert rel(out["fused_features"], ref["fused_features"]) < 1e-5
    assert rel(out["mult_features"], ref["mult_features"]) < 1e-5
    for i, x, (w, b), lin in zip(ins, S64, W64, (sp.text_projection, sp.audio_projection, sp.video_projection)):
        assert rel(i["sequence_output"].grad, x.grad) < 2e-5
        assert rel(lin.weight.grad, w.grad) < 2e-5
        assert rel(lin.bias.grad, b.grad) < 2e-5


@pytest.mark.gpu
def test_weighted_pool_kernel_fullsize_property():
    """B=64, L=512, H=512 bf16: weighted pooling with uniform weights equals the mean-pool kernel, and pooling is linear in x."""
    pkg = importlib.import_module("simple-multimodal_b200")
    K = pkg.kernels
    x = torch.randn(64, 512, 512, device="cuda").to(torch.bfloat16)
    w = torch.full((64, 512), 1.0 / 512, device="cuda")
    a, b = K.weighted_pool_fwd(x, w).float(), K.meanpool_fwd(x).float()
    assert float((a - b).abs().max()) <= 1e-2 * float(b.abs().max())
    ref = torch.einsum("bl,blh->bh", w, x.float())
    assert float((a - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    dy = torch.randn(64, 512, device="cuda").to(torch.bfloat16)
    dx = K.weighted_pool_bwd(dy, w)
    assert float((dx.float() - (dy.float()[:, None, :] / 512)).abs().max()) <= 1e-2 * float(dy.float().abs().max()) / 512
