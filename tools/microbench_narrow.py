"""GPU microbenchmark (not a pytest): the attention blocks of one MulT chunk that touch the 30-frame video stream -- forward and
backward, attn_narrow.cu (mma.sync, narrow side resident) against the 128-wide tcgen05 tiles (b200f_debug_set(10, 0)) on the
same operands, with the HBM time of the algorithmic bytes next to them.  Writes gpurun_out/microbench_narrow.json.

    python tools/microbench_narrow.py [--batch 256] [--iters 20] [--dropout 0.1]
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
lib = pkg._lib.lib()

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--dropout", type=float, default=0.1)
args = ap.parse_args()
dev, heads, W = torch.device("cuda"), 8, 512
peak = 6547.2
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        flush.zero_()                                   # 256 MB > the 126 MB L2: every timed launch starts cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


out = {"batch": args.batch, "dropout": args.dropout, "hbm_peak_gbs": peak, "cases": {}}
for Lq, Lk in ((512, 30), (30, 512), (30, 30)):
    g = torch.Generator(device=dev).manual_seed(0)
    q, k, v, do = (torch.randn(args.batch, L, W, device=dev, generator=g).to(torch.bfloat16) for L in (Lq, Lk, Lk, Lq))
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    db = torch.zeros(3, W, device=dev)
    drop = (args.dropout, 1, 2) if args.dropout > 0 else None
    o, lse = K.attn_fwd(q, k, v, heads, 0.125, dropout=drop)
    row = args.batch * W * 2
    fwd_bytes = row * (2 * Lq + 2 * Lk)                  # read Q, K, V; write O
    bwd_bytes = row * (4 * Lq + 4 * Lk)                  # read Q, dO, O, K, V; write dQ, dK, dV
    rec = {"fwd_bytes": fwd_bytes, "bwd_bytes": bwd_bytes, "fwd_hbm_us": fwd_bytes / peak / 1e3, "bwd_hbm_us": bwd_bytes / peak / 1e3}
    for name, flag in (("narrow", 1), ("tcgen05_128", 0)):
        lib.b200f_debug_set(10, flag)
        f = timeit(lambda: K.attn_fwd(q, k, v, heads, 0.125, out=o, dropout=drop))
        b = timeit(lambda: K.attn_bwd(do, q, k, v, o, lse, heads, 0.125, dq, dk, dv, dbq=db[0], dbv=db[2], dropout=drop))
        rec[name] = {"fwd_us": f * 1e3, "bwd_us": b * 1e3, "fwd_gbs": fwd_bytes / f / 1e6, "bwd_gbs": bwd_bytes / b / 1e6,
                     "fwd_frac_of_hbm_peak": fwd_bytes / f / 1e6 / peak, "bwd_frac_of_hbm_peak": bwd_bytes / b / 1e6 / peak}
    lib.b200f_debug_set(10, 1)
    out["cases"][f"{Lq}x{Lk}"] = rec
    print(f"{Lq}x{Lk}: " + json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench_narrow.json"), "w"), indent=1)
