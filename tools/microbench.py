"""GPU microbenchmark (not a pytest): every MulT GEMM shape (fwd / dgrad / wgrad) and the attention kernels through the C ABI,
next to cuBLAS / torch SDPA on the same operands.  Writes gpurun_out/microbench.json.

    python tools/microbench.py [--chunk 256] [--only gemm|attn] [--iters 10]
"""
import argparse
import importlib
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simple-multimodal_b200")
K = pkg.kernels

ap = argparse.ArgumentParser()
ap.add_argument("--chunk", type=int, default=256)
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--filter", default="")
ap.add_argument("--no-tma-store", action="store_true", help="A/B: LDS + STG copy-out instead of bulk tensor stores (b200f_debug_set(8, 1))")
ap.add_argument("--no-tma-aux", action="store_true", help="A/B: GEMMs with a residual / ReLU-mask block keep the LDS + STG copy-out (b200f_debug_set(11, 1))")
ap.add_argument("--six-stages", action="store_true", help="A/B: 6-stage pair GEMM for launches without an aux block (b200f_debug_set(7, 1))")
args = ap.parse_args()
if args.no_tma_store:
    pkg._lib.lib().b200f_debug_set(8, 1)
if args.no_tma_aux:
    pkg._lib.lib().b200f_debug_set(11, 1)
if args.six_stages:
    pkg._lib.lib().b200f_debug_set(7, 1)
dev = torch.device("cuda")
H = 512
bf = torch.bfloat16


def timeit(fn, iters=args.iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {}
if args.only in ("", "gemm"):
    M = args.chunk * 512
    x512 = torch.randn(M, 512, device=dev).to(bf)
    x2048 = torch.randn(M, 2048, device=dev).to(bf)
    x3072 = torch.randn(M, 3072, device=dev).to(bf)
    x1536 = x3072[:, :1536]
    res512 = torch.randn(M, 512, device=dev).to(bf)
    cases = []
    for name, n_out, k_in in [("proj6H", 3072, 512), ("selfqkv", 1536, 512), ("outproj", 512, 512), ("ffn1", 2048, 512), ("ffn2", 512, 2048)]:
        w = (torch.randn(n_out, k_in, device=dev) * 0.03).to(bf)
        bias = torch.randn(n_out, device=dev)
        xin = {512: x512, 2048: x2048}[k_in]
        dy = {512: res512, 2048: x2048, 3072: x3072, 1536: x1536}[n_out]
        yout = torch.empty(M, n_out, device=dev, dtype=bf)
        dxout = torch.empty(M, k_in, device=dev, dtype=bf)
        dw = torch.zeros(n_out, k_in, device=dev, dtype=torch.float32)
        resid = res512 if n_out == 512 else None
        cases.append((f"{name}_fwd M{M} N{n_out} K{k_in}", 2.0 * M * n_out * k_in,
                      lambda xin=xin, w=w, bias=bias, yout=yout, resid=resid, n_out=n_out: K.linear_fwd(xin, w, bias, relu=(n_out == 2048), residual=resid, out=yout),
                      lambda xin=xin, w=w: torch.matmul(xin, w.t())))
        cases.append((f"{name}_dgrad M{M} N{k_in} K{n_out}", 2.0 * M * n_out * k_in,
                      lambda dy=dy, w=w, dxout=dxout: K.linear_dgrad(dy, w, out=dxout),
                      lambda dy=dy, w=w: torch.matmul(dy, w)))
        if name == "ffn1":        # FFN1 input gradient + the residual path (K = 2048 -> N = 512, residual block in the epilogue)
            cases.append((f"{name}_dgrad_res M{M} N{k_in} K{n_out}", 2.0 * M * n_out * k_in,
                          lambda dy=dy, w=w, dxout=dxout: K.linear_dgrad(dy, w, out=dxout, residual=res512),
                          lambda dy=dy, w=w: torch.matmul(dy, w)))
        if name == "ffn2":        # FFN2 input gradient with the ReLU mask of the stored hidden layer in the epilogue (K = 512 -> N = 2048)
            cs = torch.zeros(k_in, device=dev)
            cases.append((f"{name}_dgrad_mask M{M} N{k_in} K{n_out}", 2.0 * M * n_out * k_in,
                          lambda dy=dy, w=w, dxout=dxout, cs=cs: K.linear_dgrad(dy, w, out=dxout, relu_mask=x2048, colsum=cs),
                          lambda dy=dy, w=w: torch.matmul(dy, w)))
            cases.append((f"{name}_dgrad_maskonly M{M} N{k_in} K{n_out}", 2.0 * M * n_out * k_in,
                          lambda dy=dy, w=w, dxout=dxout: K.linear_dgrad(dy, w, out=dxout, relu_mask=x2048),
                          lambda dy=dy, w=w: torch.matmul(dy, w)))
            cases.append((f"{name}_dgrad_colsumonly M{M} N{k_in} K{n_out}", 2.0 * M * n_out * k_in,
                          lambda dy=dy, w=w, dxout=dxout, cs=cs: K.linear_dgrad(dy, w, out=dxout, colsum=cs),
                          lambda dy=dy, w=w: torch.matmul(dy, w)))
        if name == "ffn2":        # ... and with the mask as one bit per element (written by the FFN1 forward epilogue)
            cs2 = torch.zeros(k_in, device=dev)
            hb = K.sign_bits_for(x512, k_in)
            hb.random_(-2**31, 2**31 - 1)
            cases.append((f"{name}_dgrad_bits M{M} N{k_in} K{n_out}", 2.0 * M * n_out * k_in,
                          lambda dy=dy, w=w, dxout=dxout, hb=hb: K.linear_dgrad(dy, w, out=dxout, sign_bits=hb),
                          lambda dy=dy, w=w: torch.matmul(dy, w)))
            cases.append((f"{name}_dgrad_bits_colsum M{M} N{k_in} K{n_out}", 2.0 * M * n_out * k_in,
                          lambda dy=dy, w=w, dxout=dxout, hb=hb, cs2=cs2: K.linear_dgrad(dy, w, out=dxout, sign_bits=hb, colsum=cs2),
                          lambda dy=dy, w=w: torch.matmul(dy, w)))
            for mode in (1, 2):
                def dbg(dy=dy, w=w, dxout=dxout, hb=hb, cs2=cs2, mode=mode):
                    pkg._lib.lib().b200f_debug_set(13, mode)
                    K.linear_dgrad(dy, w, out=dxout, sign_bits=hb, colsum=cs2)
                    pkg._lib.lib().b200f_debug_set(13, 0)
                cases.append((f"{name}_dgrad_bits_colsum_dbg{mode} M{M} N{k_in} K{n_out}", 2.0 * M * n_out * k_in, dbg, lambda dy=dy, w=w: torch.matmul(dy, w)))
        if name == "ffn1":
            hb1 = K.sign_bits_for(xin, n_out)
            cases.append((f"{name}_fwd_bits M{M} N{n_out} K{k_in}", 2.0 * M * n_out * k_in,
                          lambda xin=xin, w=w, bias=bias, yout=yout, hb1=hb1: K.linear_fwd(xin, w, bias, relu=True, out=yout, sign_bits_out=hb1),
                          lambda xin=xin, w=w: torch.matmul(xin, w.t())))
            cases.append((f"{name}_fwd_dropout_bits M{M} N{n_out} K{k_in}", 2.0 * M * n_out * k_in,
                          lambda xin=xin, w=w, bias=bias, yout=yout, hb1=hb1: K.linear_fwd(xin, w, bias, relu=True, out=yout, dropout=(0.1, 11, 22), sign_bits_out=hb1),
                          lambda xin=xin, w=w: torch.matmul(xin, w.t())))
        if name == "ffn1":        # FFN1 forward with the hidden-layer dropout generated in the epilogue (the training configuration)
            cases.append((f"{name}_fwd_dropout M{M} N{n_out} K{k_in}", 2.0 * M * n_out * k_in,
                          lambda xin=xin, w=w, bias=bias, yout=yout: K.linear_fwd(xin, w, bias, relu=True, out=yout, dropout=(0.1, 11, 22)),
                          lambda xin=xin, w=w: torch.matmul(xin, w.t())))
        cases.append((f"{name}_wgrad M{n_out} N{k_in} K{M}", 2.0 * M * n_out * k_in,
                      lambda dy=dy, xin=xin, dw=dw: K.linear_wgrad(dy, xin, dw),
                      lambda dy=dy, xin=xin: torch.matmul(dy.t(), xin)))
    for name, flops, ours, ref in cases:
        if args.filter and args.filter not in name:
            continue
        ms = timeit(ours)
        ms_ref = timeit(ref)
        out[name] = {"ms": ms, "tflops": flops / ms / 1e9, "cublas_ms": ms_ref, "cublas_tflops": flops / ms_ref / 1e9}
        print(name, {k: round(v, 3) for k, v in out[name].items()}, flush=True)

if args.only in ("", "attn"):
    B, heads = args.chunk, 8
    scale = 1.0 / math.sqrt(64)
    for Lq, Lk in [(512, 512), (512, 30), (30, 512)]:
        name = f"attn Lq{Lq} Lk{Lk} B{B}"
        if args.filter and args.filter not in name:
            continue
        pq = torch.randn(B, Lq, 6 * H, device=dev).to(bf)
        pkv = torch.randn(B, Lk, 6 * H, device=dev).to(bf)
        q, k, v = pq[:, :, :H], pkv[:, :, 2 * H:3 * H], pkv[:, :, 3 * H:4 * H]
        o = torch.empty(B, Lq, H, device=dev, dtype=bf)
        do = torch.randn(B, Lq, H, device=dev).to(bf)
        dpq, dpkv = torch.empty_like(pq), torch.empty_like(pkv)
        flops = 4.0 * B * heads * Lq * Lk * 64
        _, lse = K.attn_fwd(q, k, v, heads, scale, out=o)
        ms_f = timeit(lambda: K.attn_fwd(q, k, v, heads, scale, out=o))
        ms_b = timeit(lambda: K.attn_bwd(do, q, k, v, o, lse, heads, scale, dpq[:, :, :H], dpkv[:, :, 2 * H:3 * H], dpkv[:, :, 3 * H:4 * H]))
        qh = q.reshape(B, Lq, heads, 64).transpose(1, 2).contiguous().requires_grad_(True)
        kh = k.reshape(B, Lk, heads, 64).transpose(1, 2).contiguous().requires_grad_(True)
        vh = v.reshape(B, Lk, heads, 64).transpose(1, 2).contiguous().requires_grad_(True)
        ms_rf = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
        oo = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
        g = torch.randn_like(oo)
        ms_rfb = timeit(lambda: torch.autograd.grad(torch.nn.functional.scaled_dot_product_attention(qh, kh, vh), (qh, kh, vh), g))
        out[name] = {"fwd_ms": ms_f, "fwd_tflops": flops / ms_f / 1e9, "bwd_ms": ms_b, "bwd_tflops": 2.5 * flops / ms_b / 1e9,
                     "sdpa_fwd_ms": ms_rf, "sdpa_fwd_tflops": flops / ms_rf / 1e9, "sdpa_fwdbwd_ms": ms_rfb}
        print(name, {k_: round(v_, 3) for k_, v_ in out[name].items()}, flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = (args.only or "all") + ("_no_tma_aux" if args.no_tma_aux else "")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"microbench_{tag}.json"), "w"), indent=1)
