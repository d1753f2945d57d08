"""Tensor-level wrappers over the C ABI (no autograd here): each function checks its tensors,
marshals pointers / leading dimensions / the current CUDA stream and calls libb200fusion.so.
PyTorch only provides device memory and streams; every FLOP runs in the library's kernels."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from ._lib import B200FusionError, check, dtype_code, lib, ptr, require_cuda, stream_ptr

Tensor = torch.Tensor


def _rows(x: Tensor):
    """View a [..., K] tensor with contiguous last dim as (rows, K, ld)."""
    if x.dim() == 2:
        if x.stride(1) != 1 and x.size(1) != 1:
            raise B200FusionError("matrix rows must be contiguous in the last dim")
        return x.size(0), x.size(1), x.stride(0)
    if not x.is_contiguous():
        raise B200FusionError("N-d operand must be contiguous")
    k = x.size(-1)
    return x.numel() // k, k, k


_EVENT_POOL = []


def _event():
    return _EVENT_POOL.pop() if _EVENT_POOL else torch.cuda.Event(enable_timing=True)


def prealloc_profile_events(n: int) -> None:
    """Create the CUDA events the live GEMM profile will need up front, outside any timed region."""
    while len(_EVENT_POOL) < n:
        _EVENT_POOL.append(torch.cuda.Event(enable_timing=True))


# Optional live profile of the dominant kernel (bench.py): CUDA-event pairs around every GEMM launch.
GEMM_PROFILE = None          # None, or a list collecting (start_event, stop_event, flops, is_tensor_core)


def gemm(a: Tensor, b: Tensor, *, M: int, N: int, K: int, a_layout: int = 0, b_layout: int = 0,
         out: Optional[Tensor] = None, bias: Optional[Tensor] = None, residual: Optional[Tensor] = None,
         relu_mask: Optional[Tensor] = None, relu: bool = False, alpha: float = 1.0, accumulate: bool = False,
         out_f32: bool = False, split_k: int = 1, colsum: Optional[Tensor] = None, dropout=None,
         sign_bits_out: Optional[Tensor] = None, sign_bits: Optional[Tensor] = None) -> Tensor:
    """C[M,N] = epi(alpha * sum_k A(m,k) B(n,k)); see b200f_gemm in include/b200_fusion.h.
    `a`/`b` are 2-D views (row stride = leading dimension).  `sign_bits_out` / `sign_bits`: int32 [M, >= N/32], the one-bit form of
    `relu_mask` written by the ReLU forward GEMM and read by the next Linear's input-gradient GEMM (`sign_bits_for`)."""
    require_cuda(a, b, out, bias, residual, relu_mask, sign_bits_out, sign_bits)
    for name, t in (("sign_bits_out", sign_bits_out), ("sign_bits", sign_bits)):
        if t is not None and (t.dtype != torch.int32 or t.dim() != 2 or t.size(0) != M or t.stride(1) != 1 or t.size(1) * 32 < N):
            raise B200FusionError(f"gemm: {name} must be an int32 [M, >= N/32] tensor with contiguous rows")
    dt = a.dtype
    if b.dtype != dt:
        raise B200FusionError(f"gemm operand dtypes differ: {dt} vs {b.dtype}")
    flags = 0
    if relu:
        flags |= L.EPI_RELU
    want_f32 = accumulate or out_f32 or dt == torch.float32
    if out is None:
        if accumulate:
            raise B200FusionError("gemm(accumulate=True) needs an `out` buffer")
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if want_f32 else dt)
    if want_f32 and out.dtype != torch.float32:
        raise B200FusionError("gemm: fp32 output requested but `out` is not float32")
    if not want_f32 and out.dtype != dt:
        raise B200FusionError("gemm: `out` dtype must match the operands")
    if accumulate:
        flags |= L.EPI_ACCUM
    elif want_f32 and dt != torch.float32:
        flags |= L.EPI_OUT_F32
    if bias is not None and bias.dtype != torch.float32:
        raise B200FusionError("gemm: bias must be float32")
    if colsum is not None and (colsum.dtype != torch.float32 or not colsum.is_contiguous() or colsum.numel() != N or want_f32 and dt != torch.float32):
        raise B200FusionError("gemm: colsum must be a contiguous float32 [N] accumulator and the output must be in the operand dtype")
    for name, t in (("residual", residual), ("relu_mask", relu_mask)):
        if t is not None and t.dtype != dt:
            raise B200FusionError(f"gemm: {name} dtype must match the operands")
    args = L.GemmArgs(M=M, N=N, K=K, a_layout=a_layout, b_layout=b_layout,
                      A=a.data_ptr(), lda=a.stride(0), B=b.data_ptr(), ldb=b.stride(0),
                      C=out.data_ptr(), ldc=out.stride(0),
                      bias=None if bias is None else bias.data_ptr(),
                      residual=None if residual is None else residual.data_ptr(),
                      ldr=0 if residual is None else residual.stride(0),
                      relu_mask=None if relu_mask is None else relu_mask.data_ptr(),
                      ldm=0 if relu_mask is None else relu_mask.stride(0),
                      alpha=alpha, flags=flags, dtype=dtype_code(dt), split_k=split_k,
                      colsum=None if colsum is None else colsum.data_ptr(),
                      sign_bits_out=None if sign_bits_out is None else sign_bits_out.data_ptr(),
                      sign_bits=None if sign_bits is None else sign_bits.data_ptr(),
                      ldsb=(sign_bits_out if sign_bits_out is not None else sign_bits).stride(0) if (sign_bits_out is not None or sign_bits is not None) else 0,
                      **_drop_fields(dropout))
    if GEMM_PROFILE is None:
        check(lib().b200f_gemm(C.byref(args), stream_ptr()), "b200f_gemm")
        return out
    e0, e1 = _event(), _event()
    e0.record()
    check(lib().b200f_gemm(C.byref(args), stream_ptr()), "b200f_gemm")
    e1.record()
    tc = dt == torch.bfloat16 and a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0 and N >= 8 and K >= 8
    GEMM_PROFILE.append((e0, e1, 2.0 * M * N * K, tc))
    return out


def _drop_fields(dropout):
    """dropout = None or (p, seed_lo, seed_hi): in-kernel inverted dropout with a counter-based mask (csrc/common.cuh)."""
    if dropout is None or dropout[0] <= 0.0:
        return dict(dropout_p=0.0, drop_seed_lo=0, drop_seed_hi=0)
    p, lo, hi = dropout
    if not 0.0 <= p < 1.0:
        raise B200FusionError(f"dropout probability {p} not in [0, 1)")
    return dict(dropout_p=float(p), drop_seed_lo=int(lo) & 0xFFFFFFFF, drop_seed_hi=int(hi) & 0xFFFFFFFF)


def _wgrad_split(tokens: int, n_out: int, k_in: int) -> int:
    tiles = ((n_out + 127) // 128) * ((k_in + 255) // 256)
    want = max(1, (2 * 148) // max(tiles, 1))
    return max(1, min(want, tokens // 512 if tokens >= 512 else 1))


def sign_bits_for(x2: Tensor, n_out: int) -> Optional[Tensor]:
    """An int32 [M, n_out/32] buffer for the one-bit ReLU' mask of `linear_fwd(x2, w[n_out, K], relu=True)`, or None where the
    tcgen05 epilogue that writes it does not apply (fp32 parity mode, ragged widths): the caller then keeps `relu_mask=`."""
    if x2.dtype != torch.bfloat16 or n_out % 64 or x2.size(1) % 8 or x2.stride(0) % 8 or x2.data_ptr() % 16:
        return None
    return torch.empty((x2.size(0), n_out // 32), device=x2.device, dtype=torch.int32)


def linear_fwd(x2: Tensor, w: Tensor, bias: Optional[Tensor], *, relu=False, residual=None, out=None, dropout=None, sign_bits_out=None) -> Tensor:
    """y[M,N] = dropout(relu(x2[M,K] w[N,K]^T + bias (+ residual)))  (relu / dropout optional).
    `sign_bits_out` (from `sign_bits_for`): also write bit (m, n) = [y > 0] for the backward of the ReLU (and dropout)."""
    M, K = x2.shape
    return gemm(x2, w, M=M, N=w.size(0), K=K, out=out, bias=bias, residual=residual, relu=relu, dropout=dropout, sign_bits_out=sign_bits_out)


def linear_dgrad(dy2: Tensor, w: Tensor, *, relu_mask=None, sign_bits=None, residual=None, out=None, colsum=None, alpha: float = 1.0) -> Tensor:
    """dx[M,K] = dy2[M,N] w[N,K]  (B operand N-contiguous: no transposed weight copy), zeroed where `relu_mask` <= 0 / where its
    one-bit form `sign_bits` is clear.
    `colsum` [K] fp32 += column sums of dx: the bias gradient of the Linear that produced this GEMM's input activation."""
    M, N = dy2.shape
    return gemm(dy2, w, M=M, N=w.size(1), K=N, a_layout=0, b_layout=1, out=out, relu_mask=relu_mask, sign_bits=sign_bits, residual=residual,
                colsum=colsum, alpha=alpha)


def linear_wgrad(dy2: Tensor, x2: Tensor, dw: Tensor) -> Tensor:
    """dw[N,K] (fp32) += dy2[M,N]^T x2[M,K]; both operands are read in place, reduction over tokens."""
    M, N = dy2.shape
    K = x2.size(1)
    return gemm(dy2, x2, M=N, N=K, K=M, a_layout=1, b_layout=1, out=dw, accumulate=True,
                split_k=_wgrad_split(M, N, K))


def colsum_accum(x2: Tensor, out: Tensor) -> None:
    M, N = x2.shape
    check(lib().b200f_colsum_accum(ptr(x2), C.c_int64(x2.stride(0)), ptr(out), C.c_int64(M), C.c_int64(N),
                                   dtype_code(x2.dtype), stream_ptr()), "b200f_colsum_accum")


def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, post1=None, post2=None):
    rows, H, _ = _rows(x)
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    check(lib().b200f_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(post1), ptr(post2), ptr(y), ptr(mean), ptr(rstd),
                                    C.c_int64(rows), C.c_int32(H), C.c_float(eps), dtype_code(x.dtype), stream_ptr()),
          "b200f_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, gamma: Tensor, dgamma: Tensor, dbeta: Tensor,
                  dres: Optional[Tensor] = None, dxsum: Optional[Tensor] = None) -> Tensor:
    rows, H, _ = _rows(x)
    dx = torch.empty_like(x)
    check(lib().b200f_layernorm_bwd(ptr(dy), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dres), ptr(dx), ptr(dgamma),
                                    ptr(dbeta), ptr(dxsum), C.c_int64(rows), C.c_int32(H), dtype_code(x.dtype), stream_ptr()),
          "b200f_layernorm_bwd")
    return dx


def add(a: Tensor, b: Tensor, c: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    out = torch.empty_like(a) if out is None else out
    check(lib().b200f_add(ptr(a), ptr(b), ptr(c), ptr(out), C.c_int64(a.numel()), dtype_code(a.dtype), stream_ptr()), "b200f_add")
    return out


def relu_bwd(dy: Tensor, ref: Tensor) -> Tensor:
    dx = torch.empty_like(dy)
    check(lib().b200f_relu_bwd(ptr(dy), ptr(ref), ptr(dx), C.c_int64(dy.numel()), dtype_code(dy.dtype), stream_ptr()), "b200f_relu_bwd")
    return dx


def zeros_f32(shape, device) -> Tensor:
    return torch.zeros(shape, device=device, dtype=torch.float32)


def cast_to_bf16(src: Tensor, scale: float = 1.0, out: Optional[Tensor] = None) -> Tensor:
    src = src.contiguous()
    dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16) if out is None else out
    if dst.dtype != torch.bfloat16 or dst.numel() != src.numel() or not dst.is_contiguous():
        raise B200FusionError("cast_to_bf16: `out` must be a contiguous bfloat16 tensor of the source's size")
    check(lib().b200f_cast_f32_to_bf16(ptr(src), ptr(dst), C.c_int64(src.numel()), C.c_float(scale), stream_ptr()), "b200f_cast_f32_to_bf16")
    return dst


def cast_to_f32(src: Tensor) -> Tensor:
    src = src.contiguous()
    dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    check(lib().b200f_cast_bf16_to_f32(ptr(src), ptr(dst), C.c_int64(src.numel()), stream_ptr()), "b200f_cast_bf16_to_f32")
    return dst


def meanpool_fwd(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    B, Lx, H = x.shape
    if out is None:
        out = torch.empty((B, H), device=x.device, dtype=x.dtype)
    check(lib().b200f_meanpool_fwd(ptr(x), ptr(out), C.c_int64(out.stride(0)), C.c_int32(B), C.c_int32(Lx), C.c_int32(H),
                                   dtype_code(x.dtype), stream_ptr()), "b200f_meanpool_fwd")
    return out


def stage_pool(x: Tensor, copy: Tensor, mean: Optional[Tensor], mask: Optional[Tensor], col: int) -> None:
    """copy[b,l,:] = keep * x[b,l,:]; mean[b,:] = keep * mean_l x[b,l,:] (mean may be a column slice of a wider [B, *] tensor, or None);
    keep = mask[b, col] (mask None: 1).  One read of x."""
    require_cuda(x, copy, mean, mask)
    B, Lx, H = x.shape
    if copy.shape != x.shape or copy.dtype != x.dtype or not x.is_contiguous() or not copy.is_contiguous():
        raise B200FusionError("stage_pool: x / copy must be contiguous [B,L,H] tensors of one dtype")
    if mean is not None and (mean.shape != (B, H) or mean.dtype != x.dtype or mean.stride(1) != 1):
        raise B200FusionError("stage_pool: mean must be [B,H] rows of the input dtype")
    check(lib().b200f_stage_pool(ptr(x), ptr(copy), ptr(mean), C.c_int64(0 if mean is None else mean.stride(0)), ptr(mask), C.c_int32(col),
                                 C.c_int32(B), C.c_int32(Lx), C.c_int32(H), dtype_code(x.dtype), stream_ptr()), "b200f_stage_pool")


def add_rowbcast(a: Tensor, b: Tensor, c: Tensor, v: Tensor, scale: float, Lx: int, out: Optional[Tensor] = None) -> Tensor:
    """out[r,:] = a[r,:] + b[r,:] + c[r,:] + scale * v[r // Lx, :]; a/b/c contiguous [rows, H], v [rows // Lx, H] rows (may be a column slice)."""
    rows, H = a.shape
    out = torch.empty_like(a) if out is None else out
    if v.stride(1) != 1 or v.size(0) * Lx != rows or v.size(1) != H:
        raise B200FusionError("add_rowbcast: v must hold one contiguous row of H per sample")
    check(lib().b200f_add_rowbcast(ptr(a), ptr(b), ptr(c), ptr(v), C.c_int64(v.stride(0)), C.c_float(scale), ptr(out), C.c_int64(rows), C.c_int32(Lx),
                                   C.c_int32(H), dtype_code(a.dtype), stream_ptr()), "b200f_add_rowbcast")
    return out


def meanpool_bwd(dy: Tensor, Lx: int) -> Tensor:
    B, H = dy.shape
    dx = torch.empty((B, Lx, H), device=dy.device, dtype=dy.dtype)
    check(lib().b200f_meanpool_bwd(ptr(dy), C.c_int64(dy.stride(0)), ptr(dx), C.c_int32(B), C.c_int32(Lx), C.c_int32(H),
                                   dtype_code(dy.dtype), stream_ptr()), "b200f_meanpool_bwd")
    return dx


def weighted_pool_fwd(x: Tensor, w: Tensor) -> Tensor:
    """y[b,:] = sum_l w[b,l] x[b,l,:]  (w fp32 [B,L], contiguous)"""
    B, Lx, H = x.shape
    out = torch.empty((B, H), device=x.device, dtype=x.dtype)
    check(lib().b200f_weighted_pool_fwd(ptr(x), ptr(w), ptr(out), C.c_int64(out.stride(0)), C.c_int32(B), C.c_int32(Lx), C.c_int32(H),
                                        dtype_code(x.dtype), stream_ptr()), "b200f_weighted_pool_fwd")
    return out


def weighted_pool_bwd(dy: Tensor, w: Tensor) -> Tensor:
    B, H = dy.shape
    Lx = w.size(1)
    dx = torch.empty((B, Lx, H), device=dy.device, dtype=dy.dtype)
    check(lib().b200f_weighted_pool_bwd(ptr(dy), C.c_int64(dy.stride(0)), ptr(w), ptr(dx), C.c_int32(B), C.c_int32(Lx), C.c_int32(H),
                                        dtype_code(dy.dtype), stream_ptr()), "b200f_weighted_pool_bwd")
    return dx


def concat3_fwd(t: Tensor, a: Tensor, v: Tensor, mask: Optional[Tensor]) -> Tensor:
    B, H = t.shape
    cat = torch.empty((B, 3 * H), device=t.device, dtype=t.dtype)
    check(lib().b200f_concat3_fwd(ptr(t), ptr(a), ptr(v), ptr(mask), ptr(cat), C.c_int64(B), C.c_int32(H), dtype_code(t.dtype), stream_ptr()),
          "b200f_concat3_fwd")
    return cat


def concat3_bwd(dcat: Tensor, mask: Optional[Tensor], H: int):
    B = dcat.size(0)
    outs = [torch.empty((B, H), device=dcat.device, dtype=dcat.dtype) for _ in range(3)]
    check(lib().b200f_concat3_bwd(ptr(dcat), ptr(mask), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), C.c_int32(0), C.c_int64(B), C.c_int32(H),
                                  dtype_code(dcat.dtype), stream_ptr()), "b200f_concat3_bwd")
    return outs


def rowmask_apply_(x: Tensor, mask: Tensor, col: int) -> Tensor:
    B = x.size(0)
    Lx = 1 if x.dim() == 2 else x.size(1)
    check(lib().b200f_rowmask_apply(ptr(x), ptr(mask), C.c_int32(col), C.c_int64(B), C.c_int64(Lx), C.c_int32(x.size(-1)),
                                    dtype_code(x.dtype), stream_ptr()), "b200f_rowmask_apply")
    return x


def rowmask_copy(src: Tensor, dst: Tensor, mask: Optional[Tensor], col: int) -> Tensor:
    """dst[b, l, :] = src[b, l, :] * mask[b, col] (mask None: plain copy); src/dst contiguous [B,L,H] (or [B,H]) of one dtype."""
    require_cuda(src, dst, mask)
    if src.shape != dst.shape or src.dtype != dst.dtype or not src.is_contiguous() or not dst.is_contiguous():
        raise B200FusionError("rowmask_copy: src/dst must be contiguous tensors of one shape and dtype")
    B = src.size(0)
    Lx = 1 if src.dim() == 2 else src.size(1)
    check(lib().b200f_rowmask_copy(ptr(src), ptr(dst), ptr(mask), C.c_int32(col), C.c_int64(B), C.c_int64(Lx), C.c_int32(src.size(-1)),
                                   dtype_code(src.dtype), stream_ptr()), "b200f_rowmask_copy")
    return dst


def l2norm_fwd(y: Tensor, eps: float):
    rows, D = y.shape
    z = torch.empty_like(y)
    norm = torch.empty(rows, device=y.device, dtype=torch.float32)
    check(lib().b200f_l2norm_fwd(ptr(y), ptr(z), ptr(norm), C.c_int64(rows), C.c_int32(D), C.c_float(eps), dtype_code(y.dtype), stream_ptr()),
          "b200f_l2norm_fwd")
    return z, norm


def l2norm_bwd(dz: Tensor, z: Tensor, norm: Tensor, eps: float) -> Tensor:
    rows, D = z.shape
    dy = torch.empty_like(z)
    check(lib().b200f_l2norm_bwd(ptr(dz), ptr(z), ptr(norm), ptr(dy), C.c_int64(rows), C.c_int32(D), C.c_float(eps), dtype_code(z.dtype),
                                 stream_ptr()), "b200f_l2norm_bwd")
    return dy


def _tokens(x: Tensor):
    """[B,L,W] view with contiguous last dim and stride(0) == L*stride(1) -> (ptr, ld)."""
    if x.dim() != 3 or x.stride(2) != 1 or x.stride(0) != x.size(1) * x.stride(1):
        raise B200FusionError("attention operands must be [B,L,H*D] token-major views")
    return x.data_ptr(), x.stride(1)


def attn_pool_supported(q: Tensor, heads: int) -> bool:
    """whether attn_fwd can produce the pooled output (`pooled=`) for this operand: the bf16 / head-dim-64 kernels"""
    return q.dtype == torch.bfloat16 and q.size(-1) // heads == 64


def attn_fwd(q: Tensor, k: Tensor, v: Tensor, heads: int, scale: float, out: Optional[Tensor] = None, dropout=None,
             pooled: Optional[Tensor] = None):
    """`pooled` (optional, [B, H*D] in q's dtype or fp32): receives mean over the Lq rows of the stored output, computed from partial
    column sums the attention epilogue writes (one slot per row group, no atomics: bit-reproducible) + one small finishing kernel."""
    require_cuda(q, k, v, pooled)
    B, Lq, W = q.shape
    Lk = k.size(1)
    D = W // heads
    if pooled is not None and (pooled.shape != (B, W) or pooled.stride(1) != 1 or not attn_pool_supported(q, heads)):
        raise B200FusionError("attn_fwd: pooled must be [B, H*D] rows (bf16, head dim 64 only)")
    if out is None:
        out = torch.empty((B, Lq, W), device=q.device, dtype=q.dtype)
    lse = torch.empty((B, heads, Lq), device=q.device, dtype=torch.float32)
    (qp, ldq), (kp, ldk), (vp, ldv), (op, ldo) = _tokens(q), _tokens(k), _tokens(v), _tokens(out)
    args = L.AttnArgs(B=B, H=heads, Lq=Lq, Lk=Lk, D=D, Q=qp, ldq=ldq, K=kp, ldk=ldk, V=vp, ldv=ldv, O=op, ldo=ldo,
                      LSE=lse.data_ptr(), scale=scale, dtype=dtype_code(q.dtype), **_drop_fields(dropout))
    partial = None
    if pooled is not None:
        parts = int(lib().b200f_attn_pool_parts(C.byref(args)))
        if parts <= 0:
            raise B200FusionError("attn_fwd: no pooled output for this shape / dtype")
        partial = torch.empty((B, parts, W), device=q.device, dtype=torch.float32)
        args.pool_sum = partial.data_ptr()
    check(lib().b200f_attn_fwd(C.byref(args), stream_ptr()), "b200f_attn_fwd")
    if pooled is not None:
        check(lib().b200f_pool_finish(ptr(partial), ptr(pooled), C.c_int64(pooled.stride(0)), C.c_int64(B), C.c_int32(parts), C.c_int32(W),
                                      C.c_float(1.0 / Lq), dtype_code(pooled.dtype), stream_ptr()), "b200f_pool_finish")
    return out, lse


ATTN_BWD_SCRATCH_MAX = 6 << 30      # bytes; one MulT chunk of 1024 samples at 512 x 512, 8 heads needs 4 GiB


def attn_bwd(do: Tensor, q: Tensor, k: Tensor, v: Tensor, o: Tensor, lse: Tensor, heads: int, scale: float,
             dq: Tensor, dk: Tensor, dv: Tensor, dbq: Optional[Tensor] = None, dbk: Optional[Tensor] = None,
             dbv: Optional[Tensor] = None, dropout=None) -> None:
    """dq/dk/dv <- attention backward.  dbq/dbk/dbv (optional fp32 [H*D] views, contiguous) are INCREMENTED by the column sums
    of dq/dk/dv over all tokens: the bias gradients of the producing projections, summed in the kernels' epilogue."""
    for t in (dbq, dbk, dbv):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != q.size(2)):
            raise B200FusionError("attn_bwd: bias-gradient accumulators must be contiguous float32 [H*D]")
    B, Lq, W = q.shape
    Lk = k.size(1)
    D = W // heads
    delta = torch.empty_like(lse)
    (qp, ldq), (kp, ldk), (vp, ldv), (op, ldo) = _tokens(q), _tokens(k), _tokens(v), _tokens(o)
    (gp, ldg), (dqp, lddq), (dkp, lddk), (dvp, lddv) = _tokens(do), _tokens(dq), _tokens(dk), _tokens(dv)
    args = L.AttnArgs(B=B, H=heads, Lq=Lq, Lk=Lk, D=D, Q=qp, ldq=ldq, K=kp, ldk=ldk, V=vp, ldv=ldv, O=op, ldo=ldo,
                      LSE=lse.data_ptr(), scale=scale, dtype=dtype_code(q.dtype), dO=gp, lddo=ldg, dQ=dqp, lddq=lddq,
                      dK=dkp, lddk=lddk, dV=dvp, lddv=lddv, delta=delta.data_ptr(),
                      dbq=None if dbq is None else dbq.data_ptr(), dbk=None if dbk is None else dbk.data_ptr(),
                      dbv=None if dbv is None else dbv.data_ptr(), **_drop_fields(dropout))
    # scratch for the score gradient (bf16 [B, H, Lk, Lq]): with it the tcgen05 backward computes dS once and dQ = dS K is a
    # memory-bound GEMM; without it (shape not served, or the scratch would exceed ATTN_BWD_SCRATCH_MAX) dQ recomputes S / P / dP
    need = int(lib().b200f_attn_bwd_ws_bytes(C.byref(args)))
    if 0 < need <= ATTN_BWD_SCRATCH_MAX:
        ws = torch.empty(need, device=q.device, dtype=torch.uint8)
        args.bwd_ws, args.bwd_ws_bytes = ws.data_ptr(), need
    check(lib().b200f_attn_bwd(C.byref(args), stream_ptr()), "b200f_attn_bwd")


def _infonce_ws(Bl: int, Bg: int, dt, for_grad: bool, device) -> Tensor:
    fn = lib().b200f_infonce_workspace_bytes
    n = fn(C.c_int64(Bl), C.c_int64(Bg), dtype_code(dt), C.c_int32(1 if for_grad else 0))
    return torch.empty(n, device=device, dtype=torch.uint8)


def infonce_lse(x: Tensor, y: Tensor, diag_off: int, inv_tau: float, want_diag: bool = True):
    Bl, D = x.shape
    Bg = y.size(0)
    lse = torch.empty(Bl, device=x.device, dtype=torch.float32)
    diag = torch.empty(Bl, device=x.device, dtype=torch.float32) if want_diag else None
    ws = _infonce_ws(Bl, Bg, x.dtype, False, x.device)
    check(lib().b200f_infonce_lse(ptr(x), ptr(y), ptr(lse), ptr(diag), C.c_int64(Bl), C.c_int64(Bg), C.c_int32(D), C.c_int64(diag_off),
                                  C.c_float(inv_tau), dtype_code(x.dtype), ptr(ws), C.c_size_t(ws.numel()), stream_ptr()), "b200f_infonce_lse")
    return lse, diag


def infonce_grad(x: Tensor, y: Tensor, lse_x: Tensor, lse_y: Tensor, coef: float, gscale: Optional[Tensor], dx: Tensor,
                 accumulate: bool, diag_off: int, inv_tau: float) -> None:
    Bl, D = x.shape
    Bg = y.size(0)
    ws = _infonce_ws(Bl, Bg, x.dtype, True, x.device)
    check(lib().b200f_infonce_grad(ptr(x), ptr(y), ptr(lse_x), ptr(lse_y), C.c_float(coef), ptr(gscale), ptr(dx), C.c_int32(int(accumulate)),
                                   C.c_int64(Bl), C.c_int64(Bg), C.c_int32(D), C.c_int64(diag_off), C.c_float(inv_tau), dtype_code(x.dtype),
                                   ptr(ws), C.c_size_t(ws.numel()), stream_ptr()), "b200f_infonce_grad")


def _drop_args(dropout):
    d = _drop_fields(dropout)
    return C.c_float(d["dropout_p"]), C.c_uint32(d["drop_seed_lo"]), C.c_uint32(d["drop_seed_hi"])


def gat_fwd(xp: Tensor, att_src: Tensor, att_dst: Tensor, bias: Tensor, heads: int, slope: float, dropout=None):
    B = xp.size(0)
    Cc = bias.numel()
    out = torch.empty((B, 3, Cc), device=xp.device, dtype=xp.dtype)
    alpha = torch.empty((B, 3, heads, 3), device=xp.device, dtype=torch.float32)
    check(lib().b200f_gat_fwd(ptr(xp), ptr(att_src), ptr(att_dst), ptr(bias), ptr(out), ptr(alpha), C.c_int64(B), C.c_int32(heads), C.c_int32(Cc),
                              C.c_float(slope), *_drop_args(dropout), dtype_code(xp.dtype), stream_ptr()), "b200f_gat_fwd")
    return out, alpha


def gat_bwd(dout: Tensor, out: Tensor, xp: Tensor, alpha: Tensor, att_src: Tensor, att_dst: Tensor, heads: int, slope: float, dropout=None):
    B = xp.size(0)
    Cc = out.size(-1)
    dxp = torch.empty_like(xp)
    datt_src = torch.zeros(heads * Cc, device=xp.device, dtype=torch.float32)
    datt_dst = torch.zeros_like(datt_src)
    dbias = torch.zeros(Cc, device=xp.device, dtype=torch.float32)
    check(lib().b200f_gat_bwd(ptr(dout), ptr(out), ptr(xp), ptr(alpha), ptr(att_src), ptr(att_dst), ptr(dxp), ptr(datt_src), ptr(datt_dst),
                              ptr(dbias), C.c_int64(B), C.c_int32(heads), C.c_int32(Cc), C.c_float(slope), *_drop_args(dropout),
                              dtype_code(xp.dtype), stream_ptr()),
          "b200f_gat_bwd")
    return dxp, datt_src, datt_dst, dbias


def tok3_attn_fwd(qkv: Tensor, heads: int, scale: float, dropout=None):
    B = qkv.size(0)
    H = qkv.size(-1) // 3
    ctx = torch.empty((B, 3, H), device=qkv.device, dtype=qkv.dtype)
    probs = torch.empty((B, heads, 3, 3), device=qkv.device, dtype=torch.float32)
    avgw = torch.empty((B, 3, 3), device=qkv.device, dtype=torch.float32)
    check(lib().b200f_tok3_attn_fwd(ptr(qkv), ptr(ctx), ptr(probs), ptr(avgw), C.c_int64(B), C.c_int32(heads), C.c_int32(H), C.c_float(scale),
                                    *_drop_args(dropout), dtype_code(qkv.dtype), stream_ptr()), "b200f_tok3_attn_fwd")
    return ctx, probs, avgw


def tok3_attn_bwd(dctx: Tensor, davgw: Optional[Tensor], qkv: Tensor, probs: Tensor, heads: int, scale: float, dropout=None) -> Tensor:
    B = qkv.size(0)
    H = qkv.size(-1) // 3
    dqkv = torch.empty_like(qkv)
    check(lib().b200f_tok3_attn_bwd(ptr(dctx), ptr(davgw), ptr(qkv), ptr(probs), ptr(dqkv), C.c_int64(B), C.c_int32(heads), C.c_int32(H),
                                    C.c_float(scale), *_drop_args(dropout), dtype_code(qkv.dtype), stream_ptr()), "b200f_tok3_attn_bwd")
    return dqkv


def gate_mix_fwd(att: Tensor, logits: Tensor):
    B, _, H = att.shape
    gate = torch.empty((B, 3), device=att.device, dtype=torch.float32)
    mixed = torch.empty((B, H), device=att.device, dtype=att.dtype)
    check(lib().b200f_gate_mix_fwd(ptr(att), ptr(logits), ptr(gate), ptr(mixed), C.c_int64(B), C.c_int32(H), dtype_code(att.dtype), stream_ptr()),
          "b200f_gate_mix_fwd")
    return gate, mixed


def gate_mix_bwd(dmixed: Tensor, dgate_ext: Optional[Tensor], att: Tensor, gate: Tensor, logits_like: Tensor):
    B, _, H = att.shape
    datt = torch.empty_like(att)
    dlogits = torch.empty_like(logits_like)
    check(lib().b200f_gate_mix_bwd(ptr(dmixed), ptr(dgate_ext), ptr(att), ptr(gate), ptr(datt), ptr(dlogits), C.c_int64(B), C.c_int32(H),
                                   dtype_code(att.dtype), stream_ptr()), "b200f_gate_mix_bwd")
    return datt, dlogits


def late_combine_fwd(lt: Tensor, la: Tensor, lv: Tensor, w3: Tensor):
    B, E = lt.shape
    wsoft = torch.empty(3, device=lt.device, dtype=torch.float32)
    fused = torch.empty_like(lt)
    check(lib().b200f_late_combine_fwd(ptr(lt), ptr(la), ptr(lv), ptr(w3), ptr(wsoft), ptr(fused), C.c_int64(B), C.c_int32(E),
                                       dtype_code(lt.dtype), stream_ptr()), "b200f_late_combine_fwd")
    return fused, wsoft


def late_combine_bwd(dfused: Tensor, lt: Tensor, la: Tensor, lv: Tensor, wsoft: Tensor, dwsoft_ext: Optional[Tensor]):
    B, E = lt.shape
    dl = [torch.empty_like(lt) for _ in range(3)]
    dw3 = torch.zeros(3, device=lt.device, dtype=torch.float32)
    check(lib().b200f_late_combine_bwd(ptr(dfused), ptr(lt), ptr(la), ptr(lv), ptr(wsoft), ptr(dwsoft_ext), ptr(dl[0]), ptr(dl[1]), ptr(dl[2]),
                                       ptr(dw3), C.c_int64(B), C.c_int32(E), dtype_code(lt.dtype), stream_ptr()), "b200f_late_combine_bwd")
    return dl, dw3


def modality_mask(B: int, rate: float, seed: int, offset: int, device) -> Tensor:
    mask = torch.empty((B, 3), device=device, dtype=torch.float32)
    check(lib().b200f_modality_mask(ptr(mask), C.c_int64(B), C.c_float(rate), C.c_uint64(seed), C.c_uint64(offset), stream_ptr()),
          "b200f_modality_mask")
    return mask


def dropout_epoch(value: int, add: bool = False) -> None:
    """Device-side dropout epoch (csrc/common.cuh): epoch += value (add) or epoch = value, ordered on the current stream."""
    check(lib().b200f_dropout_epoch(C.c_uint32(int(value) & 0xFFFFFFFF), C.c_int32(1 if add else 0), stream_ptr()), "b200f_dropout_epoch")


def dropout(x: Tensor, p: float, seed: int, offset: int) -> Tensor:
    y = torch.empty_like(x)
    check(lib().b200f_dropout(ptr(x), ptr(y), C.c_int64(x.numel()), C.c_float(p), C.c_uint64(seed), C.c_uint64(offset), dtype_code(x.dtype),
                              stream_ptr()), "b200f_dropout")
    return y
