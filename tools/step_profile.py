"""Live (hot-cache, un-serialised) per-entry-point timing of one fusion-head step: every C-ABI call is bracketed by CUDA
events on the launching stream.  Complements the ncu launch list (cold-cache, serialised).  Not a bench: the events add
launch gaps, so the SUM is an upper bound of the step's kernel time; use bench.py for the step time itself.

    python tools/step_profile.py [--workload mult_b256] [--steps 3] > gpurun_out/step_profile.md
"""
import argparse
import importlib
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402  (workload table, Cfg)

pkg = importlib.import_module("simple-multimodal_b200")
K, L = pkg.kernels, pkg._lib

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="mult_b256")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=0)
args = ap.parse_args()
kind, batch, lens, flag, _ = bench.WORKLOADS[args.workload]
batch = args.batch or batch
dev = torch.device("cuda")
torch.manual_seed(0)
head = getattr(pkg.fusion_layers, {"mult": "MultimodalTransformer", "hierarchical": "HierarchicalFusion",
                                   "contrastive": "ContrastiveFusion", "early": "EarlyFusion"}[kind])(bench.Cfg).to(dev)
head.train()
mt = head.mult_fusion if kind == "hierarchical" else (head if kind == "mult" else None)
if mt is not None:
    mt.graph_chunks = False          # events cannot be recorded inside replayed chunk graphs: this tool profiles the eagerly issued step
shapes = [(batch, bench.H)] * 3 if lens is None else [(batch, Ln, bench.H) for Ln in lens]
xs = [torch.randn(s, device=dev).to(torch.bfloat16).requires_grad_(True) for s in shapes]
kw = {"compute_contrastive_loss": flag} if kind in ("contrastive", "hierarchical") else {}


def step():
    for p in head.parameters():
        p.grad = None
    bench.objective(head(*xs, **kw), batch).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / args.steps

L.CALL_PROFILE = []
K.GEMM_PROFILE = []
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
prof_ms = e0.elapsed_time(e1) / args.steps
calls, L.CALL_PROFILE = L.CALL_PROFILE, None
gemms, K.GEMM_PROFILE = K.GEMM_PROFILE, None

agg = defaultdict(lambda: [0, 0.0])
for name, a, b in calls:
    agg[name][0] += 1
    agg[name][1] += a.elapsed_time(b)
total = sum(v[1] for v in agg.values()) / args.steps
print(f"# live step profile: {args.workload} (batch {batch}), {args.steps} steps\n")
print(f"step time without events {plain_ms:.3f} ms; with events {prof_ms:.3f} ms; sum of bracketed calls {total:.3f} ms/step\n")
print("| C-ABI entry | calls/step | ms/step | share of plain step | mean us |")
print("|---|---:|---:|---:|---:|")
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c / args.steps:.0f} | {ms / args.steps:.3f} | {100 * ms / args.steps / plain_ms:.1f}% | {1e3 * ms / c:.1f} |")
print("\n(b200f_gemm rows are bracketed twice -- by the entry proxy and by the GEMM profile below -- the inner one is listed.)\n")
print("| GEMM (flops-weighted) | launches/step | ms/step | TFLOP/s |")
print("|---|---:|---:|---:|")
g = defaultdict(lambda: [0, 0.0, 0.0])
for a, b, fl, tc in gemms:
    key = "tcgen05" if tc else "simt/ragged"
    g[key][0] += 1
    g[key][1] += a.elapsed_time(b)
    g[key][2] += fl
for k2, (c, ms, fl) in g.items():
    print(f"| {k2} | {c / args.steps:.0f} | {ms / args.steps:.3f} | {fl / (ms * 1e-3) / 1e12 if ms else 0:.1f} |")
