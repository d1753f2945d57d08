"""tcgen05 / SIMT GEMM against torch.matmul in fp32 (through the C ABI)."""
import ctypes as C
import importlib
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
L = importlib.import_module("simple-multimodal_b200._lib")


def run_gemm(A, B, M, N, K, a_layout, b_layout, dtype, bias=None, residual=None, mask=None, relu=False,
             accum_into=None, split_k=1, alpha=1.0, out_f32=False, sync=True):
    dev = A.device
    flags = 0
    if relu:
        flags |= L.EPI_RELU
    if accum_into is not None:
        Cbuf = accum_into
        flags |= L.EPI_ACCUM
    elif out_f32 or dtype == torch.float32:
        Cbuf = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
        flags |= L.EPI_OUT_F32 if dtype != torch.float32 else 0
    else:
        Cbuf = torch.full((M, N), float("nan"), device=dev, dtype=dtype)
    args = L.GemmArgs(M=M, N=N, K=K, a_layout=a_layout, b_layout=b_layout,
                      A=A.data_ptr(), lda=A.stride(0), B=B.data_ptr(), ldb=B.stride(0),
                      C=Cbuf.data_ptr(), ldc=Cbuf.stride(0),
                      bias=None if bias is None else bias.data_ptr(),
                      residual=None if residual is None else residual.data_ptr(),
                      ldr=0 if residual is None else residual.stride(0),
                      relu_mask=None if mask is None else mask.data_ptr(), ldm=0 if mask is None else mask.stride(0),
                      alpha=alpha, flags=flags, dtype=L.dtype_code(dtype), split_k=split_k)
    L.check(L.lib().b200f_gemm(C.byref(args), L.stream_ptr()), "b200f_gemm")
    if sync:
        torch.cuda.synchronize()
    return Cbuf


def make_operands(M, N, K, a_layout, b_layout, dtype, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((M, K) if a_layout == 0 else (K, M), device="cuda", generator=g).to(dtype)
    B = torch.randn((N, K) if b_layout == 0 else (K, N), device="cuda", generator=g).to(dtype)
    Am = A.float() if a_layout == 0 else A.float().t()
    Bm = B.float() if b_layout == 0 else B.float().t()
    return A, B, Am @ Bm.t()


def rel_err(x, ref):
    return float((x.float() - ref).norm() / ref.norm().clamp_min(1e-30))


SHAPES = [(128, 128, 64), (128, 256, 64), (256, 256, 128), (384, 512, 512), (1000, 520, 200), (4096, 1536, 512),
          (130, 24, 72), (64, 2048, 512)]


@pytest.mark.parametrize("layouts", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_bf16_tc(shape, layouts):
    M, N, K = shape
    a_l, b_l = layouts
    if (a_l == 1 and M % 8) or (b_l == 1 and N % 8) or (a_l == 0 and K % 8) or (b_l == 0 and K % 8):
        pytest.skip("leading dimension not a multiple of 8")
    A, B, ref = make_operands(M, N, K, a_l, b_l, torch.bfloat16)
    out = run_gemm(A, B, M, N, K, a_l, b_l, torch.bfloat16, out_f32=True)
    assert rel_err(out, ref) < 1e-5, f"fp32-out rel err {rel_err(out, ref)}"
    out16 = run_gemm(A, B, M, N, K, a_l, b_l, torch.bfloat16)
    assert rel_err(out16, ref) < 5e-3


@pytest.mark.parametrize("layouts", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("shape", [(128, 128, 64), (200, 136, 100), (513, 70, 33)])
def test_gemm_f32_simt(shape, layouts):
    M, N, K = shape
    A, B, ref = make_operands(M, N, K, *layouts, torch.float32)
    out = run_gemm(A, B, M, N, K, *layouts, torch.float32)
    assert rel_err(out, ref) < 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_epilogue(dtype):
    M, N, K = 300, 264, 192
    A, B, ref = make_operands(M, N, K, 0, 0, dtype, seed=1)
    g = torch.Generator(device="cuda").manual_seed(5)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(dtype)
    mask = torch.randn(M, N, device="cuda", generator=g).to(dtype)
    want = torch.relu(0.5 * ref + bias + res.float()) * (mask.float() > 0)
    got = run_gemm(A, B, M, N, K, 0, 0, dtype, bias=bias, residual=res, mask=mask, relu=True, alpha=0.5)
    assert rel_err(got, want) < (2e-6 if dtype == torch.float32 else 5e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_weight_grad_accumulate_split_k(dtype):
    M, N, K = 512, 136, 4096          # dW[N_out=512, K_in=136] over 4096 tokens, both operands MN-major
    A, B, ref = make_operands(M, N, K, 1, 1, dtype, seed=2)
    acc = torch.ones(M, N, device="cuda", dtype=torch.float32)
    run_gemm(A, B, M, N, K, 1, 1, dtype, accum_into=acc, split_k=8)
    assert rel_err(acc - 1.0, ref) < 1e-5


@pytest.mark.parametrize("dtype,shape", [(torch.bfloat16, (1000, 512, 256)), (torch.bfloat16, (4133, 2048, 512)), (torch.bfloat16, (300, 264, 192)),
                                         (torch.bfloat16, (77, 128, 64)), (torch.float32, (300, 264, 192))])
def test_gemm_fused_column_sums(dtype, shape):
    """colsum[n] += sum_m C[m,n] of the STORED output (bias gradient fused in the epilogue; N % 64 != 0 and fp32 take
    the GEMM + column-sum-kernel route inside the library).  With the ReLU-mask epilogue, as the FFN2 input-gradient GEMM uses it."""
    K = importlib.import_module("simple-multimodal_b200.kernels")
    M, N, Kd = shape
    A, B, ref = make_operands(M, N, Kd, 0, 1, dtype, seed=6)
    mask = torch.randn(M, N, device="cuda").to(dtype)
    acc = torch.full((N,), 2.0, device="cuda")
    out = K.gemm(A, B, M=M, N=N, K=Kd, a_layout=0, b_layout=1, relu_mask=mask, colsum=acc)
    torch.cuda.synchronize()
    want = ref * (mask.float() > 0)
    assert rel_err(out, want) < (2e-6 if dtype == torch.float32 else 5e-3)
    sums = out.double().sum(0) + 2.0
    assert float((acc.double() - sums).abs().max()) <= 1e-5 * float(out.float().abs().sum(0).max()) + 1e-4


@pytest.mark.parametrize("shape", [(1000, 512, 256), (4133, 2048, 512), (77, 128, 64), (300, 192, 64), (130, 64, 64), (8192, 2048, 512)])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_gemm_sign_bits_replace_the_relu_mask(shape, p_drop):
    """The one-bit ReLU' mask: the Linear+ReLU(+Dropout) forward writes bit (m, n) = [stored y > 0]; the input-gradient GEMM that reads
    the bits gives exactly what it gives when it reads the stored activation as `relu_mask` (FFN1 -> FFN2-dgrad of a MulT block)."""
    K = importlib.import_module("simple-multimodal_b200.kernels")
    M, N, Kd = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(M, Kd, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, Kd, device="cuda", generator=g) * Kd ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    bits = K.sign_bits_for(x, N)
    assert bits is not None and bits.shape == (M, N // 32)
    bits.fill_(-1)
    drop = (p_drop, 123, 456) if p_drop else None
    y = K.linear_fwd(x, w, bias, relu=True, dropout=drop, sign_bits_out=bits)
    y_plain = K.linear_fwd(x, w, bias, relu=True, dropout=drop)
    assert torch.equal(y, y_plain)                                                    # writing the bits does not touch the output
    # decode: element e of a 32-column word sits at bit 8 * (e % 4) + e / 4 (the library's own order, include/b200_fusion.h)
    e = torch.arange(32, device="cuda")
    pos = 8 * (e % 4) + e // 4
    dec = ((bits.to(torch.int64).unsqueeze(-1) >> pos) & 1).reshape(M, N).bool()
    assert torch.equal(dec, y > 0)
    frac = float(dec.float().mean())
    assert 0.3 < frac < 0.6                                                          # about half pass the ReLU, 10 % of those dropped
    dy = torch.randn(M, 64, device="cuda", generator=g).bfloat16()                   # next Linear: [64, N] weight, dgrad K = 64 -> N
    w2 = torch.randn(64, N, device="cuda", generator=g).bfloat16()
    c1, c2 = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    d_mask = K.linear_dgrad(dy, w2, relu_mask=y, colsum=c1, alpha=1.25)
    d_bits = K.linear_dgrad(dy, w2, sign_bits=bits, colsum=c2, alpha=1.25)
    assert torch.equal(d_mask, d_bits)
    assert torch.allclose(c1, c2, rtol=1e-5, atol=1e-3)
    assert float(d_bits[~dec].abs().max()) == 0.0 and float(d_bits[dec].abs().min()) >= 0.0


def test_gemm_sign_bits_bad_args():
    K = importlib.import_module("simple-multimodal_b200.kernels")
    x = torch.randn(256, 64, device="cuda").bfloat16()
    w = torch.randn(128, 64, device="cuda").bfloat16()
    bits = K.sign_bits_for(x, 128)
    with pytest.raises(L.B200FusionError, match="ReLU"):
        K.linear_fwd(x, w, None, sign_bits_out=bits)                                  # no ReLU: the bit trick assumes non-negative outputs
    assert K.sign_bits_for(x.float(), 128) is None and K.sign_bits_for(x, 120) is None
    with pytest.raises(L.B200FusionError, match="tcgen05 path only"):
        K.gemm(x.float(), w.float(), M=256, N=128, K=64, relu=True, sign_bits_out=bits)
    with pytest.raises(L.B200FusionError, match="int32"):
        K.linear_fwd(x, w, None, relu=True, sign_bits_out=bits.float())
    with pytest.raises(L.B200FusionError, match="pass one"):
        K.linear_dgrad(x, torch.randn(64, 128, device="cuda").bfloat16(), relu_mask=torch.ones(256, 128, device="cuda").bfloat16(), sign_bits=bits)


def test_gemm_ragged_bf16_uses_cuda_cores():
    """Leading dimensions that are not 16-byte multiples (7-class logits, 3 gate logits) cannot go through TMA."""
    M, N, K = 130, 7, 512
    A, B, ref = make_operands(M, N, K, 0, 0, torch.bfloat16, seed=3)
    out = run_gemm(A, B, M, N, K, 0, 0, torch.bfloat16)
    assert out.stride(0) == 7 and rel_err(out, ref) < 5e-3
    A2, B2, ref2 = make_operands(M, 512, 7, 0, 1, torch.bfloat16, seed=4)      # dgrad through a [7,512] weight
    assert rel_err(run_gemm(A2, B2, M, 512, 7, 0, 1, torch.bfloat16), ref2) < 5e-3


def test_gemm_bad_args_report_errors():
    A = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(L.B200FusionError, match="empty shape"):
        run_gemm(A, B, 128, 128, 0, 0, 0, torch.bfloat16)
