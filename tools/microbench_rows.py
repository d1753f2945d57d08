"""GPU microbenchmark (not a pytest) of the HBM-bound row kernels of a MulT chunk through the C ABI: LayerNorm forward
(plain / with the residual post-adds) and backward (without / with the residual gradient), the TMA-staged kernels
(csrc/rownorm_tma.cu) next to the register-staged ones (b200f_debug_set(9, 1)), with the algorithmic bytes each moves and the
achieved GB/s.  Also checks that the two families agree.  Writes gpurun_out/microbench_rows.json.

    python tools/microbench_rows.py [--rows 131072] [--iters 20]
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("simple-multimodal_b200")
K = pkg.kernels

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=256 * 512)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda")
H, bf = 512, torch.bfloat16
rows = args.rows
NSET = 3                                                   # rotate operand sets: every launch reads cold (sets exceed the 126 MB L2)


def timeit(fn, iters=args.iters, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


xs = [torch.randn(rows, H, device=dev).to(bf) for _ in range(NSET)]
p1 = [torch.randn(rows, H, device=dev).to(bf) for _ in range(NSET)]
p2 = [torch.randn(rows, H, device=dev).to(bf) for _ in range(NSET)]
dys = [torch.randn(rows, H, device=dev).to(bf) for _ in range(NSET)]
gamma = torch.rand(H, device=dev) + 0.5
beta = torch.randn(H, device=dev)
tensor_bytes = rows * H * 2
out = {"rows": rows, "H": H, "dtype": "bf16"}
saved = {}
for label, key in (("tma", 0), ("regs", 1)):
    pkg._lib.lib().b200f_debug_set(9, key)
    y, mean, rstd = K.layernorm_fwd(xs[0], gamma, beta, 1e-5)
    y2, _, _ = K.layernorm_fwd(xs[0], gamma, beta, 1e-5, p1[0], p2[0])
    dg, db, dsum = (torch.zeros(H, device=dev) for _ in range(3))
    dx = K.layernorm_bwd(dys[0], xs[0], mean, rstd, gamma, dg, db)
    dg2, db2, dsum2 = (torch.zeros(H, device=dev) for _ in range(3))
    dx2 = K.layernorm_bwd(dys[0], xs[0], mean, rstd, gamma, dg2, db2, dres=p1[0], dxsum=dsum2)
    saved[label] = (y, mean, rstd, y2, dx, dg, db, dx2, dg2, db2, dsum2)
    sink = torch.zeros(H, device=dev)
    cases = {
        "ln_fwd_plain": (lambda i: K.layernorm_fwd(xs[i % NSET], gamma, beta, 1e-5), 2 * tensor_bytes),
        "ln_fwd_post2": (lambda i: K.layernorm_fwd(xs[i % NSET], gamma, beta, 1e-5, p1[i % NSET], p2[i % NSET]), 4 * tensor_bytes),
        "ln_bwd": (lambda i: K.layernorm_bwd(dys[i % NSET], xs[i % NSET], mean, rstd, gamma, sink, sink), 3 * tensor_bytes),
        "ln_bwd_dres_dxsum": (lambda i: K.layernorm_bwd(dys[i % NSET], xs[i % NSET], mean, rstd, gamma, sink, sink, dres=p1[i % NSET], dxsum=sink),
                              4 * tensor_bytes),
    }
    for name, (fn, nbytes) in cases.items():
        ms = timeit(fn)
        out[f"{name}.{label}"] = {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / ms / 1e6}
pkg._lib.lib().b200f_debug_set(9, 0)

a, b = saved["tma"], saved["regs"]
names = ["y", "mean", "rstd", "y_post2", "dx", "dgamma", "dbeta", "dx_dres", "dgamma2", "dbeta2", "dxsum"]
agree = {}
for n, u, v in zip(names, a, b):
    d = float((u.double() - v.double()).abs().max())
    agree[n] = d / max(float(v.double().abs().max()), 1e-30)
out["max_rel_diff_tma_vs_regs"] = agree

# mean-pool over the sequence: [B, L, H] -> [B, H]
x3 = [x.view(rows // 512, 512, H) for x in xs] if rows % 512 == 0 else None
if x3 is not None:
    ref = x3[0].float().mean(1)
    got = K.meanpool_fwd(x3[0]).float()
    out["meanpool_max_abs_err_vs_torch"] = float((got - ref).abs().max())
    ms = timeit(lambda i: K.meanpool_fwd(x3[i % NSET]))
    out["meanpool_fwd"] = {"ms": ms, "algorithmic_bytes": tensor_bytes, "GBps": tensor_bytes / ms / 1e6}

# torch's copy on the same footprint = the practical HBM ceiling for a 1-read-1-write kernel
dst = [torch.empty_like(x) for x in xs]
ms = timeit(lambda i: dst[i % NSET].copy_(xs[i % NSET]))
out["torch_copy"] = {"ms": ms, "algorithmic_bytes": 2 * tensor_bytes, "GBps": 2 * tensor_bytes / ms / 1e6}

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "microbench_rows.json"), "w") as f:
    json.dump(out, f, indent=1)
for k, v in out.items():
    print(k, v)
