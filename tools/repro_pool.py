"""Scratch: hunt an intermittent 1-ulp difference in the pooled self-attention path (see repro_chunk_invariance.py)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simple-multimodal_b200")
K = pkg.kernels
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
B, L, H, heads = 64, 512, 512, 8
qkv = torch.randn(B, L, 3 * H, device=dev, generator=g).to(torch.bfloat16)
q, k, v = qkv[:, :, :H], qkv[:, :, H:2 * H], qkv[:, :, 2 * H:]
w = (torch.randn(H, H, device=dev, generator=g) * 0.05).to(torch.bfloat16)
bias = torch.randn(H, device=dev, generator=g)


def poison(nbytes=1 << 30):
    t = torch.empty(nbytes // 4, device=dev)
    t.fill_(float("nan"))
    del t


ref = None
bad = {"o": 0, "pooled": 0, "nan": 0, "gemm": 0}
for it in range(200):
    poison()
    pooled = torch.empty(B, H, device=dev, dtype=torch.bfloat16)
    o, lse = K.attn_fwd(q, k, v, heads, 0.125, pooled=pooled)
    y = K.linear_fwd(pooled, w, bias)
    torch.cuda.synchronize()
    if ref is None:
        ref = (o.clone(), pooled.clone(), y.clone())
        continue
    if not torch.equal(o, ref[0]):
        bad["o"] += 1
    if not torch.equal(pooled, ref[1]):
        bad["pooled"] += 1
        d = (pooled.float() - ref[1].float()).abs()
        rows = (d.max(1).values > 0).nonzero().flatten().tolist()
        cols = (d.max(0).values > 0).nonzero().flatten().tolist()
        print(f"it {it}: pooled differs rows {rows} cols {cols[:10]}... n={int((d > 0).sum())} max {float(d.max()):.3e} nan={bool(torch.isnan(pooled.float()).any())}", flush=True)
    if torch.isnan(pooled.float()).any() or torch.isnan(o.float()).any():
        bad["nan"] += 1
    if not torch.equal(y, ref[2]) and torch.equal(pooled, ref[1]):
        bad["gemm"] += 1
        d = (y.float() - ref[2].float()).abs()
        print(f"it {it}: GEMM differs rows {(d.max(1).values > 0).nonzero().flatten().tolist()} n={int((d > 0).sum())}", flush=True)
print("summary", bad)
