"""In-kernel dropout (attention probabilities, FFN hidden layer, GAT / 3-token attention weights): the kernels generate the
mask from a counter-based hash of (seed, row, column) instead of storing it (csrc/common.cuh).  torch's RNG stream cannot be
matched, so parity is checked the other way round: the hash is restated in numpy, the mask is built on the host, and the CUDA
results must equal explicit torch math using THAT mask -- forward and backward (mask placement, 1/(1-p) scaling, the softmax
backward through the dropped weights, torch nn/functional.py:6647-6650).  CPU part: rate and independence of the hash."""
import importlib
import math

import numpy as np
import pytest
import torch

pkg = importlib.import_module("simple-multimodal_b200")
K, ops, FL = pkg.kernels, pkg.ops, pkg.fusion_layers
M32 = np.uint64(0xFFFFFFFF)


# ---- numpy restatement of csrc/common.cuh: hash32 / drop_row_key / drop_keep / drop_threshold -------------------------------
def hash32(x):
    x = np.asarray(x, dtype=np.uint64) & M32
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7feb352d)) & M32
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846ca68b)) & M32
    x ^= x >> np.uint64(16)
    return x


def epoch_mix(epoch):
    """drop_epoch_mix: 0 for the eager default epoch 0, else an avalanche hash of the epoch (added to seed_lo)"""
    return 0 if epoch == 0 else int(hash32((epoch * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF))


def row_key(lo, hi, rows, epoch=0):
    rows = np.asarray(rows, dtype=np.uint64) & M32
    lo = np.uint64((int(lo) + epoch_mix(epoch)) & 0xFFFFFFFF)
    return hash32((lo + hash32(np.uint64(hi) ^ rows)) & M32)


def threshold(p):
    t = float(np.float32(p)) * 4294967296.0
    return 0 if t <= 0 else min(int(t), 4294967295)


def keep_mask(lo, hi, rows, cols, p, epoch=0, attn=False):
    """[len(rows), len(cols)] bool: element kept.  attn: the attention-probability sites' column term (drop_col_attn: XOR-decomposed at
    64-key granularity, equal to the plain col * Mul below 64); otherwise col * Mul (GEMM epilogue, GAT, 3-token attention)."""
    rk = row_key(lo, hi, rows, epoch)[:, None]
    cols = np.asarray(cols, dtype=np.uint64)
    cm = (cols * np.uint64(0x9E3779B1)) & M32
    if attn:
        cm = (((cols >> np.uint64(6)) * np.uint64(0xC2B2AE35)) & M32) ^ (((cols & np.uint64(63)) * np.uint64(0x9E3779B1)) & M32)
    h = ((rk ^ cm[None, :]) * np.uint64(0x85EBCA6B)) & M32
    return h >= np.uint64(threshold(p))


def inv_keep(p):
    return float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))


# ---- CPU: statistics of the mask ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("attn", [False, True], ids=["plain", "attention"])
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_mask_rate_and_independence(p, attn):
    m = keep_mask(0x1234ABCD, 0x9876FEDC, np.arange(4096), np.arange(512), p, attn=attn).astype(np.float64)
    n = m.size
    assert abs(m.mean() - (1 - p)) < 4 * math.sqrt(p * (1 - p) / n)
    assert np.abs(m.mean(0) - (1 - p)).max() < 6 * math.sqrt(p * (1 - p) / m.shape[0])      # every column
    assert np.abs(m.mean(1) - (1 - p)).max() < 6 * math.sqrt(p * (1 - p) / m.shape[1])      # every row
    c = m - m.mean()
    var = (c * c).mean()
    for a, b in ((c[:, :-1], c[:, 1:]), (c[:-1, :], c[1:, :]), (c[:-1, :-1], c[1:, 1:]), (c[:, :-2], c[:, 2:])):
        assert abs((a * b).mean() / var) < 5 / math.sqrt(a.size)                              # neighbours uncorrelated
    for d in (64, 128, 192):                                                                  # across the 64-key groups of the attention term
        assert abs((c[:, :-d] * c[:, d:]).mean() / var) < 5 / math.sqrt(c[:, d:].size)
    other = keep_mask(0x1234ABCE, 0x9876FEDC, np.arange(4096), np.arange(512), p, attn=attn).astype(np.float64)
    assert abs(((other - other.mean()) * c).mean() / var) < 5 / math.sqrt(n)                 # a new seed is a new mask


def test_epochs_are_fresh_masks_not_row_permutations():
    """CUDA-graph replays advance the device-side epoch 1, 2, 3 ...: the masks of two replays must be independent draws.  (The
    first version XORed the epoch next to the row index, so replay e's row r was replay 0's row r ^ e.)"""
    rows, cols, p = np.arange(1024), np.arange(256), 0.3
    base = keep_mask(0xC0FFEE, 0x51DE, rows, cols, p, epoch=0)
    sigs0 = {m.tobytes() for m in base}
    var = p * (1 - p)
    for e in (1, 2, 3, 7):
        m = keep_mask(0xC0FFEE, 0x51DE, rows, cols, p, epoch=e)
        assert abs(m.mean() - (1 - p)) < 4 * math.sqrt(var / m.size)
        assert not any(r.tobytes() in sigs0 for r in m)                                       # no row of epoch e is a row of epoch 0
        c = ((m - m.mean()) * (base - base.mean())).mean() / var
        assert abs(c) < 5 / math.sqrt(m.size)                                                 # and the draws are uncorrelated
        prev = keep_mask(0xC0FFEE, 0x51DE, rows, cols, p, epoch=e + 1)
        assert not any(r.tobytes() in {x.tobytes() for x in m} for r in prev)                 # consecutive epochs too


def test_seed_stream_is_reproducible():
    ops.manual_seed(123)
    a = [ops.next_drop_seed() for _ in range(4)]
    ops.manual_seed(123)
    b = [ops.next_drop_seed() for _ in range(4)]
    assert a == b and len(set(a)) == 4
    assert all(0 <= lo < 2 ** 32 and 0 <= hi < 2 ** 32 for lo, hi in a)


# ---- GPU ---------------------------------------------------------------------------------------------------------------------
def rel(x, r):
    x, r = x.detach().double().cpu(), r.detach().double().cpu()
    return float((x - r).norm() / r.norm().clamp_min(1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,shape", [(torch.bfloat16, (3, 8, 300, 200)), (torch.bfloat16, (2, 8, 512, 512)), (torch.bfloat16, (5, 8, 30, 70)),
                                         (torch.bfloat16, (3, 8, 512, 30)), (torch.bfloat16, (3, 8, 30, 512)), (torch.bfloat16, (2, 8, 100, 17)),
                                         (torch.bfloat16, (2, 8, 17, 100)), (torch.float32, (2, 8, 65, 40))])
@pytest.mark.parametrize("ds_route", [1, 0], ids=["dq-from-stored-dS", "dq-recomputes"])
def test_attention_dropout_matches_explicit_mask(dtype, shape, ds_route):
    B, heads, Lq, Lk = shape
    if ds_route == 0 and not (dtype == torch.bfloat16 and Lq >= 128 and Lk >= 128):
        pytest.skip("only blocks with >= 128 queries and keys have the two dQ routes")
    pkg._lib.lib().b200f_debug_set(15, ds_route)
    try:
        _attention_dropout_case(dtype, shape)
    finally:
        pkg._lib.lib().b200f_debug_set(15, 0)


def _attention_dropout_case(dtype, shape):
    B, heads, Lq, Lk = shape
    W, p, lo, hi = heads * 64, 0.25, 0xDEADBEEF, 0x0BADF00D
    g = torch.Generator(device="cuda").manual_seed(1)
    q, k, v, do = (torch.randn(B, L, W, device="cuda", generator=g).to(dtype) for L in (Lq, Lk, Lk, Lq))
    o, lse = K.attn_fwd(q, k, v, heads, 0.125, dropout=(p, lo, hi))
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    db = torch.zeros(3, W, device="cuda")
    K.attn_bwd(do, q, k, v, o, lse, heads, 0.125, dq, dk, dv, dbq=db[0], dbk=db[1], dbv=db[2], dropout=(p, lo, hi))
    torch.cuda.synchronize()
    mask = torch.from_numpy(keep_mask(lo, hi, np.arange(B * heads * Lq), np.arange(Lk), p, attn=True)).view(B, heads, Lq, Lk).double()
    qd, kd, vd = (t.detach().double().cpu().requires_grad_(True) for t in (q, k, v))
    split = lambda t, L: t.view(B, L, heads, 64).transpose(1, 2)
    s = split(qd, Lq) @ split(kd, Lk).transpose(-1, -2) * 0.125
    pd = torch.softmax(s, -1) * mask * inv_keep(p)
    o_ref = (pd @ split(vd, Lk)).transpose(1, 2).reshape(B, Lq, W)
    (o_ref * do.double().cpu()).sum().backward()
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
    assert rel(o, o_ref) < tol
    assert float((lse.double().cpu() - torch.logsumexp(s, -1)).abs().max()) < (2e-2 if dtype == torch.bfloat16 else 1e-4)   # LSE of the UNdropped scores
    assert rel(dq, qd.grad) < tol and rel(dk, kd.grad) < tol and rel(dv, vd.grad) < tol
    assert rel(db[0], dq.double().sum((0, 1))) < 1e-3 and rel(db[2], dv.double().sum((0, 1))) < 1e-3
    # p = 0 is the plain kernel
    o0, _ = K.attn_fwd(q, k, v, heads, 0.125, dropout=(0.0, lo, hi))
    o1, _ = K.attn_fwd(q, k, v, heads, 0.125)
    assert torch.equal(o0, o1)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,shape", [(torch.bfloat16, (1000, 512, 256)), (torch.bfloat16, (300, 264, 192)), (torch.float32, (130, 96, 64))])
def test_gemm_epilogue_dropout_matches_explicit_mask(dtype, shape):
    """tcgen05 epilogue (N % 64 == 0, bf16) and the GEMM + in-place kernel route (ragged N, fp32) generate the same mask"""
    M, N, Kd = shape
    p, lo, hi = 0.1, 77, 0xABCDEF01
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(M, Kd, device="cuda", generator=g).to(dtype)
    w = (torch.randn(N, Kd, device="cuda", generator=g) * 0.1).to(dtype)
    b = torch.randn(N, device="cuda", generator=g)
    y = K.linear_fwd(x, w, b, relu=True, dropout=(p, lo, hi))
    torch.cuda.synchronize()
    mask = torch.from_numpy(keep_mask(lo, hi, np.arange(M), np.arange(N), p)).double()
    ref = torch.relu(x.double().cpu() @ w.double().cpu().T + b.double().cpu()) * mask * inv_keep(p)
    assert rel(y, ref) < (5e-3 if dtype == torch.bfloat16 else 1e-5)
    assert bool(((y.cpu() == 0) | (mask > 0)).all())                          # dropped entries are exact zeros


def _tok3_reference(qkv, heads, mask, ik):
    B, _, H3 = qkv.shape
    H, D = H3 // 3, H3 // 3 // heads
    q, k, v = (qkv[:, :, i * H:(i + 1) * H].reshape(B, 3, heads, D).transpose(1, 2) for i in range(3))
    pr = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(D), -1) * mask * ik
    return (pr @ v).transpose(1, 2).reshape(B, 3, H), pr.mean(1)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_tok3_attention_dropout(dtype):
    B, heads, H, p, lo, hi = 37, 8, 512, 0.3, 5, 6
    qkv = torch.randn(B, 3, 3 * H, device="cuda").to(dtype).requires_grad_(True)
    ctx, avgw = ops.Tok3AttnFn.apply(qkv, heads, 1 / math.sqrt(H // heads), (p, lo, hi))
    gc, gw = torch.randn_like(ctx), torch.randn_like(avgw)
    ((ctx * gc).sum() + (avgw * gw).sum()).backward()
    mask = torch.from_numpy(keep_mask(lo, hi, np.arange(B * heads * 3), np.arange(3), p)).view(B, heads, 3, 3).double()
    ref_in = qkv.detach().double().cpu().requires_grad_(True)
    rc, rw = _tok3_reference(ref_in, heads, mask, inv_keep(p))
    ((rc * gc.double().cpu()).sum() + (rw * gw.double().cpu()).sum()).backward()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel(ctx, rc) < tol and rel(avgw, rw) < tol and rel(qkv.grad, ref_in.grad) < tol


@pytest.mark.gpu
def test_gat_dropout():
    B, heads, Cc, p, lo, hi, slope = 21, 4, 64, 0.4, 9, 10, 0.2
    xp = torch.randn(B, 3, heads * Cc, device="cuda", requires_grad=True)
    a_s = torch.randn(1, heads, Cc, device="cuda", requires_grad=True)
    a_d = torch.randn(1, heads, Cc, device="cuda", requires_grad=True)
    bias = torch.randn(Cc, device="cuda", requires_grad=True)
    out = ops.GatFn.apply(xp, a_s, a_d, bias, heads, slope, (p, lo, hi))
    go = torch.randn_like(out)
    (out * go).sum().backward()
    # explicit math: alpha[b,h,i,j] over sources j of target i, dropout on alpha, mean over heads, + bias, relu
    mask = torch.from_numpy(keep_mask(lo, hi, np.arange(B * heads * 3), np.arange(3), p)).view(B, heads, 3, 3).double()
    X, S, D_, Bi = (t.detach().double().cpu().requires_grad_(True) for t in (xp, a_s, a_d, bias))
    x4 = X.view(B, 3, heads, Cc)
    es = (x4 * S.view(1, 1, heads, Cc)).sum(-1)                    # [B, node, head]
    ed = (x4 * D_.view(1, 1, heads, Cc)).sum(-1)
    e = torch.nn.functional.leaky_relu(es.permute(0, 2, 1)[:, :, None, :] + ed.permute(0, 2, 1)[:, :, :, None], slope)   # [B,h,i,j]
    al = torch.softmax(e, -1) * mask * inv_keep(p)
    agg = torch.einsum("bhij,bjhc->bihc", al, x4).mean(2) + Bi
    ref = torch.relu(agg)
    (ref * go.double().cpu()).sum().backward()
    assert rel(out, ref) < 1e-5
    for got, want in ((xp.grad, X.grad), (a_s.grad, S.grad), (a_d.grad, D_.grad), (bias.grad, Bi.grad)):
        assert rel(got, want) < 1e-5


class _Cfg:
    fusion_hidden_size, fusion_num_heads, fusion_dropout = 512, 8, 0.1
    num_emotions, graph_hidden_size, graph_num_layers, graph_dropout = 7, 512, 3, 0.1
    contrastive_temperature = 0.07


def _run_mult(head, xs, seed):
    ops.manual_seed(seed)
    for p_ in head.parameters():
        p_.grad = None
    ins = [x.clone().requires_grad_(True) for x in xs]
    out = head(*ins)
    (out["fused_features"].float() ** 2).mean().backward()
    torch.cuda.synchronize()
    return out["fused_features"].detach(), [i.grad for i in ins], {k: v.grad.clone() for k, v in head.named_parameters()}


@pytest.mark.gpu
def test_mult_trains_with_reference_default_dropout():
    """fusion_dropout = 0.1 (reference config.py:30) in training mode: runs, is reproducible under manual_seed, differs between
    seeds and from eval mode, and chunks recomputed in backward regenerate the masks of the forward pass (same gradients as
    resident chunks)."""
    torch.manual_seed(0)
    head = FL.MultimodalTransformer(_Cfg).cuda()
    head.train()
    xs = [torch.randn(6, L, 512, device="cuda").to(torch.bfloat16) for L in (96, 64, 30)]
    head.chunk_size = 2
    o1, dx1, pg1 = _run_mult(head, xs, 11)
    o2, dx2, pg2 = _run_mult(head, xs, 11)
    assert torch.equal(o1, o2) and all(torch.equal(a, b) for a, b in zip(dx1, dx2))
    o3, _, _ = _run_mult(head, xs, 12)
    assert not torch.equal(o1, o3)
    head.stash_fraction = 1e-9                      # one resident chunk, two recomputed in backward
    try:
        o4, dx4, pg4 = _run_mult(head, xs, 11)
    finally:
        head.stash_fraction = FL.MultimodalTransformer.stash_fraction
    assert torch.equal(o1, o4) and all(torch.equal(a, b) for a, b in zip(dx1, dx4))
    for k in pg1:
        assert rel(pg4[k], pg1[k]) < 1e-3 or float(pg1[k].norm()) < 1e-6, k
    assert all(torch.isfinite(g).all() for g in pg1.values())
    head.eval()
    with torch.no_grad():
        oe = head(*xs)["fused_features"]
    assert not torch.equal(oe, o1)
    assert 0.02 < rel(o1, oe) < 1.0                 # dropout perturbs, it does not destroy


@pytest.mark.gpu
def test_every_head_trains_with_reference_default_dropout():
    torch.manual_seed(0)
    ops.manual_seed(3)
    head = FL.HierarchicalFusion(_Cfg).cuda()
    head.train()
    xs = [torch.randn(8, L, 512, device="cuda").to(torch.bfloat16).requires_grad_(True) for L in (48, 32, 30)]
    out = head(*xs, compute_contrastive_loss=True)
    loss = (out["fused_features"].float() ** 2).mean() + 0.1 * sum(out["contrastive_losses"].values())
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss) and all(torch.isfinite(x.grad).all() for x in xs)
    assert all(p_.grad is not None and torch.isfinite(p_.grad).all() for p_ in head.parameters())
    # the dropped adaptive attention weights still average to rows of mean ~1 (inverted dropout is unbiased)
    assert abs(float(out["attention_weights"].sum(-1).mean()) - 1.0) < 0.2


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 8, 300, 200), (3, 8, 512, 30), (3, 8, 30, 512)])
def test_attention_dropout_mask_at_nonzero_epoch(shape):
    """the kernels fold the device-side epoch exactly as `row_key(..., epoch)` restates it (forward and backward agree)"""
    B, heads, Lq, Lk = shape
    W, p, lo, hi = heads * 64, 0.25, 0x13579BDF, 0x2468ACE0
    g = torch.Generator(device="cuda").manual_seed(3)
    q, k, v, do = (torch.randn(B, L, W, device="cuda", generator=g).to(torch.bfloat16) for L in (Lq, Lk, Lk, Lq))
    try:
        K.dropout_epoch(5)
        o, lse = K.attn_fwd(q, k, v, heads, 0.125, dropout=(p, lo, hi))
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        K.attn_bwd(do, q, k, v, o, lse, heads, 0.125, dq, dk, dv, dropout=(p, lo, hi))
        torch.cuda.synchronize()
    finally:
        K.dropout_epoch(0)
    mask = torch.from_numpy(keep_mask(lo, hi, np.arange(B * heads * Lq), np.arange(Lk), p, epoch=5, attn=True)).view(B, heads, Lq, Lk).double()
    qd, kd, vd = (t.detach().double().cpu().requires_grad_(True) for t in (q, k, v))
    split = lambda t, L: t.view(B, L, heads, 64).transpose(1, 2)
    s_ = split(qd, Lq) @ split(kd, Lk).transpose(-1, -2) * 0.125
    pd = torch.softmax(s_, -1) * mask * inv_keep(p)
    o_ref = (pd @ split(vd, Lk)).transpose(1, 2).reshape(B, Lq, W)
    (o_ref * do.double().cpu()).sum().backward()
    assert rel(o, o_ref) < 2e-2
    assert rel(dq, qd.grad) < 2e-2 and rel(dk, kd.grad) < 2e-2 and rel(dv, vd.grad) < 2e-2


@pytest.mark.gpu
def test_modality_mask_follows_the_epoch():
    """ModalityDropout.sample_mask inside a captured step: the host offset is frozen, the device epoch re-draws the mask"""
    try:
        K.dropout_epoch(0)
        a = K.modality_mask(4096, 0.3, 99, 0, "cuda")
        K.dropout_epoch(1)
        b = K.modality_mask(4096, 0.3, 99, 0, "cuda")
        K.dropout_epoch(0)
        c = K.modality_mask(4096, 0.3, 99, 0, "cuda")
    finally:
        K.dropout_epoch(0)
    assert torch.equal(a, c) and not torch.equal(a, b)
    assert float(b.sum(1).min()) >= 1.0 and abs(float(b.mean()) - 0.7) < 0.05
