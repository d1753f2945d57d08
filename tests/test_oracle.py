"""Pin the oracle: every golden vector under tests/golden/ was produced by the EXECUTED
reference (oracle/make_golden.py, float64); the oracle restatement must reproduce outputs,
loss, input gradients and parameter gradients.  Also closed-form known answers."""
import glob
import math
import os

import pytest
import torch

from oracle import fusion_oracle as fo

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))
TOL = dict(rtol=1e-9, atol=1e-11)


def run_oracle(meta, dtype=torch.float64):
    P = fo.init_params(meta["kind"], H=meta["H"], heads=meta["heads"], graph_hidden=meta["graph_hidden"],
                       graph_layers=meta["graph_layers"], seed=meta["param_seed"])
    P = {k: v.to(dtype).requires_grad_(True) for k, v in P.items()}
    xs = [x.to(dtype).requires_grad_(True) for x in fo.synthetic_features(meta["B"], meta["lens"], H=meta["H"], seed=meta["feat_seed"])]
    kind = meta["kind"]
    kw = {}
    if kind in ("mult", "adaptive", "hierarchical"):
        kw["heads"] = meta["heads"]
    if kind in ("contrastive", "hierarchical"):
        kw["compute_contrastive_loss"] = meta["flag"]
    if kind == "graph":
        kw["num_layers"] = meta["graph_layers"]
    if kind == "hierarchical":
        kw["graph_layers"] = meta["graph_layers"]
    out = fo.HEADS[kind](*xs, P, **kw)
    loss = fo.objective(out)
    loss.backward()
    return out, loss, xs, P


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_oracle_matches_executed_reference(path):
    rec = torch.load(path, weights_only=True)
    out, loss, xs, P = run_oracle(rec["meta"])
    torch.testing.assert_close(loss.detach(), rec["loss"], **TOL)
    if isinstance(out, torch.Tensor):
        torch.testing.assert_close(out.detach(), rec["outputs"]["__tensor__"], **TOL)
    else:
        assert set(rec["outputs"]) <= set(out)
        for k, ref in rec["outputs"].items():
            torch.testing.assert_close(out[k].detach(), ref, **TOL)
        for k, ref in rec["losses"].items():
            torch.testing.assert_close(out["contrastive_losses"][k].detach(), ref, **TOL)
    for x, g in zip(xs, rec["input_grads"]):
        torch.testing.assert_close(x.grad, g, **TOL)
    assert set(rec["param_grads"]) <= set(P), set(rec["param_grads"]) - set(P)
    for k, ref in rec["param_grads"].items():
        g = P[k].grad
        if rec["meta"]["full"]:
            torch.testing.assert_close(g, ref, **TOL)
        else:
            f = g.flatten()
            torch.testing.assert_close(torch.cat([f.sum()[None], f.norm()[None], f[:4]]), ref, rtol=1e-8, atol=1e-10)


def test_golden_covers_every_head():
    kinds = {torch.load(p, weights_only=True)["meta"]["kind"] for p in GOLDEN}
    assert kinds == set(fo.HEADS)


# ---- closed-form known answers ------------------------------------------------------------
def test_infonce_orthonormal_rows():
    B, tau = 16, 0.07
    z = torch.eye(B, 32, dtype=torch.float64)
    expect = math.log(1 + (B - 1) * math.exp(-1 / tau))
    assert abs(float(fo.info_nce(z, z, tau)) - expect) < 1e-12


def test_infonce_identical_rows():
    B = 37
    z = torch.ones(B, 8, dtype=torch.float64) / math.sqrt(8)
    assert abs(float(fo.info_nce(z, z, 0.07)) - math.log(B)) < 1e-12


def test_infonce_row_permutation_equivariance():
    g = torch.Generator().manual_seed(3)
    z1 = fo.l2_normalize(torch.randn(40, 16, generator=g, dtype=torch.float64))
    z2 = fo.l2_normalize(torch.randn(40, 16, generator=g, dtype=torch.float64))
    perm = torch.randperm(40, generator=g)
    assert abs(float(fo.info_nce(z1, z2, 0.07) - fo.info_nce(z1[perm], z2[perm], 0.07))) < 1e-12


def test_mult_single_key_is_attention_free():
    """L=1: softmax over one key == 1, so attention == out_proj(v_proj(kv))."""
    H, heads = 32, 4
    P = {k: v.double() for k, v in fo.init_params("mult", H=H, heads=heads, seed=1).items()}
    t, a, _ = fo.synthetic_features(3, (None, None, None), H=H, dtype=torch.float64)
    pre = "text_to_audio.attention."
    got, w = fo.multi_head_attention(t[:, None], a[:, None], P, pre, heads)
    v = fo.affine(a[:, None], P[pre + "in_proj_weight"][2 * H:], P[pre + "in_proj_bias"][2 * H:])
    torch.testing.assert_close(got, fo.affine(v, P[pre + "out_proj.weight"], P[pre + "out_proj.bias"]), **TOL)
    torch.testing.assert_close(w, torch.ones_like(w), **TOL)


def test_late_fusion_initial_weights_are_one_third():
    P = fo.init_params("late", H=32, seed=0)
    t, a, v = fo.synthetic_features(4, (None, None, None), H=32)
    out = fo.late_fusion(t, a, v, P)
    torch.testing.assert_close(out["fusion_weights"], torch.full((3,), 1 / 3))


def test_mult_is_per_sample_independent():
    H, heads = 32, 4
    P = {k: v.double() for k, v in fo.init_params("mult", H=H, heads=heads, seed=2).items()}
    xs = fo.synthetic_features(6, (5, 7, 3), H=H, dtype=torch.float64)
    full = fo.mult_fusion(*xs, P, heads=heads)["fused_features"]
    parts = torch.cat([fo.mult_fusion(*[x[i:i + 2] for x in xs], P, heads=heads)["fused_features"] for i in (0, 2, 4)])
    torch.testing.assert_close(full, parts, **TOL)


def test_modality_mask_keeps_at_least_one():
    g = torch.Generator().manual_seed(0)
    m = fo.modality_keep_mask(4096, 0.9, g)          # extreme rate: most rows need the repair
    assert m.shape == (4096, 3) and bool((m.sum(dim=1) >= 1).all())
    assert set(m.unique().tolist()) <= {0.0, 1.0}
    m2 = fo.modality_keep_mask(20000, 0.1, torch.Generator().manual_seed(1))
    assert abs(float(m2.mean()) - 0.9) < 0.01
    t, a, v = fo.synthetic_features(8, (None, 4, None), H=16)
    mt, ma, mv = fo.apply_modality_mask(t, a, v, m[:8])
    torch.testing.assert_close(ma, a * m[:8, 1][:, None, None])       # no 1/(1-p) rescale
    torch.testing.assert_close(mt, t * m[:8, 0][:, None])


def test_pooled_self_attention_identity_fp64():
    """Groundwork for the next attention kernel ("Next kernels"): MulT's self-attention outputs are only used through
    their mean over the sequence (models/fusion_layers.py:161-168), and  mean_q(P V) = (mean_q P) V.  Backward then needs no
    dO·V^T and no P^T·dO product: with g = d(pooled)/L the same row for every query,  dV = (sum_q P)^T (x) g  is rank one and
    dP[q, k] = g · V[k]  does not depend on q.  Checked against autograd through the explicit per-query attention, fp64."""
    torch.manual_seed(5)
    B, h, L, d = 2, 3, 17, 8
    q, k, v = (torch.randn(B, h, L, d, dtype=torch.float64, requires_grad=True) for _ in range(3))
    scale = d ** -0.5
    P = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)                 # [B,h,L,L]
    pooled = (P @ v).mean(dim=2)                                                # what the reference computes, then pools
    pooled_id = P.mean(dim=2, keepdim=True) @ v                                 # column means of P, one vector-matrix product
    assert float((pooled - pooled_id.squeeze(2)).abs().max()) < 1e-14
    up = torch.randn_like(pooled)
    dq, dk, dv = torch.autograd.grad(pooled, (q, k, v), up)
    with torch.no_grad():
        g = up / L                                                              # [B,h,d]: dO is this row for every query
        dv_id = P.sum(dim=2).unsqueeze(-1) * g.unsqueeze(2)                     # (sum_q P)[k] * g   -- rank one per (b, head)
        dp_row = (v * g.unsqueeze(2)).sum(-1)                                   # [B,h,L_k]: dP[q,k] = g . V[k] for every q
        delta = (P * dp_row.unsqueeze(2)).sum(-1, keepdim=True)                 # rowsum(P o dP)
        ds = P * (dp_row.unsqueeze(2) - delta) * scale
        dq_id, dk_id = ds @ k, ds.transpose(-1, -2) @ q
    for a, b in ((dv, dv_id), (dq, dq_id), (dk, dk_id)):
        assert float((a - b).abs().max()) < 1e-13
