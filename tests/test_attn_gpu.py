"""Attention kernels (tcgen05 bf16, CUDA-core fp32) against explicit torch math on the same packed, strided
projection tensors: forward output + LSE, backward dQ/dK/dV."""
import importlib
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("simple-multimodal_b200")
K = pkg.kernels

SHAPES = [(2, 8, 512, 512), (3, 8, 512, 30), (3, 8, 30, 512), (2, 8, 100, 200), (1, 8, 128, 128), (2, 8, 1, 1), (1, 8, 257, 129),
          (40, 8, 512, 512), (150, 8, 30, 30), (2, 8, 640, 384)]      # > 148 work items per kernel: several items per persistent CTA


def reference(q, k, v, heads, scale):
    B, Lq, W = q.shape
    Lk, D = k.size(1), W // heads
    qh = q.double().view(B, Lq, heads, D).transpose(1, 2)
    kh = k.double().view(B, Lk, heads, D).transpose(1, 2)
    vh = v.double().view(B, Lk, heads, D).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    lse = torch.logsumexp(s, dim=-1)
    o = torch.softmax(s, dim=-1) @ vh
    return o.transpose(1, 2).reshape(B, Lq, W), lse


def rel(x, r):
    """norm-wise relative error; an exactly-zero reference (softmax over a single key has no dQ/dK) is compared absolutely."""
    x, r = x.detach().double(), r.detach().double()
    if float(r.norm()) == 0.0:
        return float(x.abs().max())
    return float((x - r).norm() / r.norm())


def packed(B, L, width, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(B, L, width, device="cuda", generator=g).to(dtype)


@pytest.fixture(params=[0, 1, 2], ids=["persistent", "tile-per-cta", "persistent-recompute-dq"])
def tc_variant(request):
    """both generations of the tcgen05 kernels: b200f_debug_set(4 / 5, v) selects the forward / backward variant; the persistent
    backward in its two forms: dS stored by the dK/dV kernel + dQ = dS K (default with >= 128 queries and keys), and the dQ kernel that
    recomputes S / P / dP (b200f_debug_set(15, 0); also what runs when the caller gives no scratch)"""
    lib = pkg._lib.lib()
    lib.b200f_debug_set(4, 1 if request.param == 1 else 2)
    lib.b200f_debug_set(5, 1 if request.param == 1 else 0)
    lib.b200f_debug_set(15, 0 if request.param == 2 else 1)
    lib.b200f_debug_set(10, 0)          # shapes with a narrow side stay on the tcgen05 tiles here (test_narrow_attention covers attn_narrow.cu)
    yield request.param
    lib.b200f_debug_set(4, 0)
    lib.b200f_debug_set(5, 0)
    lib.b200f_debug_set(15, 0)          # the library's default: dQ recomputes (see csrc/attn_tc.cu for the measurement)
    lib.b200f_debug_set(10, 1)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "fp32"])
@pytest.mark.parametrize("shape", SHAPES)
def test_attention_forward_backward(shape, dtype, tc_variant):
    B, heads, Lq, Lk = shape
    if dtype == torch.float32 and (tc_variant >= 1 or B > 3):
        pytest.skip("the CUDA-core fp32 path has one variant; large batches are covered in bf16")
    W = heads * 64
    scale = 1 / math.sqrt(64)
    pq = packed(B, Lq, 2 * W, dtype, 1)                     # Q lives in columns [W, 2W) of a wider projection
    pkv = packed(B, Lk, 3 * W, dtype, 2)                    # K in [0,W), V in [2W,3W)
    q, k, v = pq[:, :, W:], pkv[:, :, :W], pkv[:, :, 2 * W:]
    o, lse = K.attn_fwd(q, k, v, heads, scale)
    torch.cuda.synchronize()
    qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    o_ref, lse_ref = reference(qd, kd, vd, heads, scale)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel(o, o_ref) < tol, f"O rel err {rel(o, o_ref)}"
    assert float((lse.double() - lse_ref).abs().max()) < (1e-4 if dtype == torch.float32 else 2e-2)
    do = packed(B, Lq, W, dtype, 3)
    (o_ref * do.double()).sum().backward()
    dpq, dpkv = torch.zeros_like(pq), torch.zeros_like(pkv)
    db = torch.ones(3, W, device="cuda", dtype=torch.float32)          # accumulators: the kernels ADD the column sums
    K.attn_bwd(do, q, k, v, o, lse, heads, scale, dpq[:, :, W:], dpkv[:, :, :W], dpkv[:, :, 2 * W:], dbq=db[0], dbk=db[1], dbv=db[2])
    torch.cuda.synchronize()
    gt = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel(dpq[:, :, W:], qd.grad) < gt, f"dQ {rel(dpq[:, :, W:], qd.grad)}"
    assert rel(dpkv[:, :, :W], kd.grad) < gt, f"dK {rel(dpkv[:, :, :W], kd.grad)}"
    assert rel(dpkv[:, :, 2 * W:], vd.grad) < gt, f"dV {rel(dpkv[:, :, 2 * W:], vd.grad)}"
    assert float(dpq[:, :, :W].abs().max()) == 0 and float(dpkv[:, :, W:2 * W].abs().max()) == 0     # untouched slices
    # fused bias gradients == column sums of the stored gradients (fp32 accumulation of the rounded values)
    for acc, grad in ((db[0], dpq[:, :, W:]), (db[1], dpkv[:, :, :W]), (db[2], dpkv[:, :, 2 * W:])):
        want = grad.double().sum(dim=(0, 1)) + 1.0
        assert float((acc.double() - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max())) + 1e-3 * float(grad.float().abs().max())


NARROW_SHAPES = [(3, 8, 512, 30), (3, 8, 30, 512), (2, 8, 1, 1), (150, 8, 30, 30), (2, 8, 200, 7), (2, 8, 7, 200), (2, 8, 33, 32), (2, 8, 32, 33),
                 (4, 8, 130, 30), (4, 8, 30, 130), (2, 8, 64, 1), (2, 8, 1, 64), (2, 8, 17, 1000), (2, 8, 1000, 17), (40, 8, 512, 30), (40, 8, 30, 512),
                 (1, 1, 5, 3), (2, 3, 30, 70)]


@pytest.mark.parametrize("shape", NARROW_SHAPES)
def test_narrow_attention(shape):
    """attn_narrow.cu (one side <= 32 rows: the video blocks of MulT, reference models/fusion_layers.py:146-153): forward output, LSE,
    dQ / dK / dV and the fused bias gradients against fp64 torch math on strided packed projections, ragged tiles on both sides."""
    B, heads, Lq, Lk = shape
    W = heads * 64
    scale = 1 / math.sqrt(64)
    dtype = torch.bfloat16
    pq = packed(B, Lq, 2 * W, dtype, 11)
    pkv = packed(B, Lk, 3 * W, dtype, 12)
    q, k, v = pq[:, :, W:], pkv[:, :, :W], pkv[:, :, 2 * W:]
    n0 = pkg._lib.launch_count()
    o, lse = K.attn_fwd(q, k, v, heads, scale)
    assert pkg._lib.launch_count() - n0 == 1
    torch.cuda.synchronize()
    qd, kd, vd = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    o_ref, lse_ref = reference(qd, kd, vd, heads, scale)
    assert rel(o, o_ref) < 1e-2, f"O rel err {rel(o, o_ref)}"
    assert float((lse.double() - lse_ref).abs().max()) < 2e-2
    do = packed(B, Lq, W, dtype, 13)
    (o_ref * do.double()).sum().backward()
    dpq, dpkv = torch.zeros_like(pq), torch.zeros_like(pkv)
    db = torch.ones(3, W, device="cuda", dtype=torch.float32)
    n0 = pkg._lib.launch_count()
    K.attn_bwd(do, q, k, v, o, lse, heads, scale, dpq[:, :, W:], dpkv[:, :, :W], dpkv[:, :, 2 * W:], dbq=db[0], dbk=db[1], dbv=db[2])
    assert pkg._lib.launch_count() - n0 == 1                 # one kernel: no separate delta / dKdV / column-sum pass
    torch.cuda.synchronize()
    assert rel(dpq[:, :, W:], qd.grad) < 2e-2, f"dQ {rel(dpq[:, :, W:], qd.grad)}"
    assert rel(dpkv[:, :, :W], kd.grad) < 2e-2, f"dK {rel(dpkv[:, :, :W], kd.grad)}"
    assert rel(dpkv[:, :, 2 * W:], vd.grad) < 2e-2, f"dV {rel(dpkv[:, :, 2 * W:], vd.grad)}"
    assert float(dpq[:, :, :W].abs().max()) == 0 and float(dpkv[:, :, W:2 * W].abs().max()) == 0
    for acc, grad in ((db[0], dpq[:, :, W:]), (db[1], dpkv[:, :, :W]), (db[2], dpkv[:, :, 2 * W:])):
        want = grad.double().sum(dim=(0, 1)) + 1.0
        assert float((acc.double() - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max())) + 1e-3 * float(grad.float().abs().max())
    # the 128-wide tcgen05 tiles on the same operands agree (A/B switch b200f_debug_set(10, 0))
    lib = pkg._lib.lib()
    lib.b200f_debug_set(10, 0)
    try:
        o2, lse2 = K.attn_fwd(q, k, v, heads, scale)
        torch.cuda.synchronize()
    finally:
        lib.b200f_debug_set(10, 1)
    assert rel(o, o2) < 1e-2 and float((lse - lse2).abs().max()) < 2e-2
    # and so does the first-generation narrow-key backward (b200f_debug_set(10, 2)); both are bit-reproducible run to run
    lib.b200f_debug_set(10, 2)
    try:
        dpq2, dpkv2 = torch.zeros_like(pq), torch.zeros_like(pkv)
        K.attn_bwd(do, q, k, v, o, lse, heads, scale, dpq2[:, :, W:], dpkv2[:, :, :W], dpkv2[:, :, 2 * W:])
        torch.cuda.synchronize()
    finally:
        lib.b200f_debug_set(10, 1)
    assert rel(dpq2, dpq) < 1e-2 and rel(dpkv2, dpkv) < 1e-2
    dpq3, dpkv3 = torch.zeros_like(pq), torch.zeros_like(pkv)
    K.attn_bwd(do, q, k, v, o, lse, heads, scale, dpq3[:, :, W:], dpkv3[:, :, :W], dpkv3[:, :, 2 * W:])
    torch.cuda.synchronize()
    assert torch.equal(dpq3, dpq) and torch.equal(dpkv3, dpkv)


def test_attention_tc_matches_cuda_core_path():
    B, heads, Lq, Lk = 2, 8, 300, 77
    W = heads * 64
    q, k, v = (packed(B, L, W, torch.bfloat16, s) for L, s in ((Lq, 5), (Lk, 6), (Lk, 7)))
    o_tc, lse_tc = K.attn_fwd(q, k, v, heads, 0.125)
    pkg._lib.lib().b200f_debug_force_simt_attention(1)
    try:
        o_cc, lse_cc = K.attn_fwd(q, k, v, heads, 0.125)
    finally:
        pkg._lib.lib().b200f_debug_force_simt_attention(0)
    torch.cuda.synchronize()
    assert rel(o_tc, o_cc) < 1e-2 and float((lse_tc - lse_cc).abs().max()) < 1e-2


@pytest.mark.parametrize("shape", [(3, 8, 512, 512), (2, 8, 300, 77), (5, 8, 30, 30), (3, 8, 30, 512), (2, 8, 100, 17), (2, 8, 128, 128), (40, 8, 512, 512)])
@pytest.mark.parametrize("p_drop", [0.0, 0.2])
def test_attention_pooled_output(shape, p_drop):
    """pooled=: mean over the queries of the stored O from the forward epilogue's partial column sums (every kernel family that has it),
    bit-reproducible from run to run (no atomics)"""
    B, heads, Lq, Lk = shape
    W = heads * 64
    q, k, v = packed(B, Lq, W, torch.bfloat16, 21), packed(B, Lk, W, torch.bfloat16, 22), packed(B, Lk, W, torch.bfloat16, 23)
    drop = (p_drop, 5, 6) if p_drop else None
    pooled = torch.empty(B, W, device="cuda")
    o, _ = K.attn_fwd(q, k, v, heads, 0.125, dropout=drop, pooled=pooled)
    o2, _ = K.attn_fwd(q, k, v, heads, 0.125, dropout=drop)
    torch.cuda.synchronize()
    assert torch.equal(o, o2)
    want = o.double().mean(dim=1)
    assert float((pooled.double() - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-7
    pb = torch.empty(B, W, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        p2 = torch.empty(B, W, device="cuda")
        K.attn_fwd(q, k, v, heads, 0.125, dropout=drop, pooled=p2)
        K.attn_fwd(q, k, v, heads, 0.125, dropout=drop, pooled=pb)
        assert torch.equal(p2, pooled) and torch.equal(pb, pooled.to(torch.bfloat16))
    assert K.attn_pool_supported(q, heads) and not K.attn_pool_supported(q.float(), heads)
    with pytest.raises(pkg.B200FusionError):
        K.attn_fwd(q.float(), k.float(), v.float(), heads, 0.125, pooled=pooled)
