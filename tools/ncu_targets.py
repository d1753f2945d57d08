"""Launch the kernels that `ncu --set full` captures are wanted of, once each, at the benchmark shapes (not a bench, not a test):

    ncu --set full --clock-control none -k regex:'attn_|infonce_' -o gpurun_out/r02_attn_infonce python tools/ncu_targets.py

attention: one MulT chunk's block shapes (B=256, 8 heads, d=64; 512x512, 512x30, 30x512; dropout 0.1 as in the bench);
InfoNCE: the fused lse / grad kernels at B_g = 4096 (1 GPU) and B_l = 4096 against B_g = 32768 (the 8-GPU global batch)."""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simple-multimodal_b200")
K = pkg.kernels

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="", help="attn | infonce | gemm (gemm only on request)")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
heads, W = 8, 512

if args.only in ("", "attn", "attn512"):
    for Lq, Lk in ((512, 512), (512, 30), (30, 512))[:1 if args.only == "attn512" else 3]:
        q, k, v, do = (torch.randn(args.batch, L, W, device=dev, generator=g).to(torch.bfloat16) for L in (Lq, Lk, Lk, Lq))
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        db = torch.zeros(3, W, device=dev)
        drop = (0.1, 0x1234, 0x5678)
        for _ in range(args.reps):
            o, lse = K.attn_fwd(q, k, v, heads, 0.125, dropout=drop)
            K.attn_bwd(do, q, k, v, o, lse, heads, 0.125, dq, dk, dv, dbq=db[0], dbv=db[2], dropout=drop)
        torch.cuda.synchronize()

if args.only in ("", "infonce"):
    for Bl, Bg in ((4096, 4096), (4096, 32768)):
        x = torch.nn.functional.normalize(torch.randn(Bl, 256, device=dev, generator=g), dim=-1).to(torch.bfloat16)
        y = torch.nn.functional.normalize(torch.randn(Bg, 256, device=dev, generator=g), dim=-1).to(torch.bfloat16)
        y[:Bl] = x
        for _ in range(args.reps):
            lse_x, diag = K.infonce_lse(x, y, 0, 1.0 / 0.07)
            lse_y, _ = K.infonce_lse(y, x, 0, 1.0 / 0.07, want_diag=False) if Bg == Bl else (torch.zeros(Bg, device=dev), None)
            dx = torch.zeros(Bl, 256, device=dev)
            gs = torch.ones(1, device=dev)
            K.infonce_grad(x, y, lse_x, lse_y, 1.0 / (0.07 * 2 * Bg), gs, dx, True, 0, 1.0 / 0.07)
        torch.cuda.synchronize()
if args.only == "gemm":
    # the K = 512 -> N = 2048 GEMMs of one MulT block at chunk 256 (M = 131072), in launch order:
    #   0 FFN1 forward (ReLU)            1 + dropout + one-bit mask out      2 FFN2 input gradient, plain
    #   3 + column sums                  4 + stored-activation mask + sums   5 + one-bit mask + sums (the step's configuration)
    M = args.batch * 512
    x = torch.randn(M, 512, device=dev, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(2048, 512, device=dev, generator=g) * 0.04).to(torch.bfloat16)
    b1 = torch.randn(2048, device=dev, generator=g) * 0.1
    w2 = (torch.randn(512, 2048, device=dev, generator=g) * 0.02).to(torch.bfloat16)
    dy = torch.randn(M, 512, device=dev, generator=g).to(torch.bfloat16)
    hid = torch.empty(M, 2048, device=dev, dtype=torch.bfloat16)
    dh = torch.empty_like(hid)
    bits = K.sign_bits_for(x, 2048)
    cs = torch.zeros(2048, device=dev)
    for _ in range(args.reps):
        K.linear_fwd(x, w1, b1, relu=True, out=hid)
        K.linear_fwd(x, w1, b1, relu=True, out=hid, dropout=(0.1, 11, 22), sign_bits_out=bits)
        K.linear_dgrad(dy, w2, out=dh)
        K.linear_dgrad(dy, w2, out=dh, colsum=cs)
        K.linear_dgrad(dy, w2, out=dh, relu_mask=hid, colsum=cs)
        K.linear_dgrad(dy, w2, out=dh, sign_bits=bits, colsum=cs)
    torch.cuda.synchronize()
print("ncu_targets: done")
