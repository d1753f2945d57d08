"""Pin the oracle against the second batch of golden vectors (tests/golden/r2/, produced by the EXECUTED reference through
oracle/make_golden_r2.py): CrossModalTransformer on its own (reference models/fusion_layers.py:182-211), batch sizes 1 and 257,
and MulT at the benchmark sequence lengths (512, 512, 30), H = 512."""
import glob
import os

import pytest
import torch

from oracle import fusion_oracle as fo
from test_oracle import TOL, run_oracle

R2 = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "r2", "*.pt")))
HEAD_FIX = [p for p in R2 if not os.path.basename(p).startswith("cross_")]
CROSS_FIX = [p for p in R2 if os.path.basename(p).startswith("cross_")]


def digest(t):
    f = t.detach().double().flatten()
    return torch.cat([f.sum()[None], f.norm()[None], f[:4] if f.numel() >= 4 else torch.nn.functional.pad(f, (0, 4 - f.numel()))])


def cross_inputs(meta):
    P = fo.init_params("mult", H=meta["H"], heads=meta["heads"], seed=meta["param_seed"])
    P = {k[len("text_to_audio."):]: v for k, v in P.items() if k.startswith("text_to_audio.")}
    g = torch.Generator().manual_seed(meta["feat_seed"])
    q = torch.randn(meta["B"], meta["Lq"], meta["H"], generator=g)
    kv = torch.randn(meta["B"], meta["Lk"], meta["H"], generator=g)
    return P, q, kv


def test_r2_fixture_set_is_complete():
    names = {os.path.basename(p)[:-3] for p in R2}
    assert {"cross_h64", "cross_h64_q1", "cross_h512", "mult3d_h512_bench_b1", "early_h64_b1", "early_h64_b257",
            "contrastive_h64_b257", "hier2d_h32_b1"} <= names


@pytest.mark.parametrize("path", HEAD_FIX, ids=[os.path.basename(p)[:-3] for p in HEAD_FIX])
def test_oracle_matches_executed_reference_r2(path):
    rec = torch.load(path, weights_only=True)
    m = rec["meta"]
    out, loss, xs, P = run_oracle(m)
    torch.testing.assert_close(loss.detach(), rec["loss"], **TOL)
    if isinstance(out, torch.Tensor):
        torch.testing.assert_close(out.detach(), rec["outputs"]["__tensor__"], **TOL)
    else:
        for k, ref in rec["outputs"].items():
            torch.testing.assert_close(out[k].detach(), ref, **TOL)
        for k, ref in rec["losses"].items():
            torch.testing.assert_close(out["contrastive_losses"][k].detach(), ref, **TOL)
    if m["input_digest"]:
        for x, dg, rows in zip(xs, rec["input_grad_digest"], rec["input_grad_rows"]):
            torch.testing.assert_close(digest(x.grad), dg, rtol=1e-8, atol=1e-10)
            torch.testing.assert_close(x.grad[:, ::m["row_stride"]], rows, **TOL)
    else:
        for x, g in zip(xs, rec["input_grads"]):
            torch.testing.assert_close(x.grad, g, **TOL)
    for k, ref in rec["param_grads"].items():
        if m["full"]:
            torch.testing.assert_close(P[k].grad, ref, **TOL)
        else:
            torch.testing.assert_close(digest(P[k].grad), ref, rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("path", CROSS_FIX, ids=[os.path.basename(p)[:-3] for p in CROSS_FIX])
def test_oracle_cross_block_matches_executed_reference(path):
    rec = torch.load(path, weights_only=True)
    m = rec["meta"]
    P, q, kv = cross_inputs(m)
    P = {k: v.double().requires_grad_(True) for k, v in P.items()}
    q, kv = q.double().requires_grad_(True), kv.double().requires_grad_(True)
    out = fo.cross_block(q, kv, P, "", m["heads"])
    loss = fo.objective(out)
    loss.backward()
    torch.testing.assert_close(loss.detach(), rec["loss"], **TOL)
    torch.testing.assert_close(out.detach(), rec["output"], **TOL)
    torch.testing.assert_close(q.grad, rec["input_grads"][0], **TOL)
    torch.testing.assert_close(kv.grad, rec["input_grads"][1], **TOL)
    for k, ref in rec["param_grads"].items():
        if m["full"]:
            torch.testing.assert_close(P[k].grad, ref, **TOL)
        else:
            torch.testing.assert_close(digest(P[k].grad), ref, rtol=1e-8, atol=1e-10)
