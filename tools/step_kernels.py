"""Per-kernel time of whole training steps as they really run (hot caches, captured chunk graphs replayed, small heads issued
eagerly), from CUPTI activity records (torch.profiler) -- the complement of the ncu launch list, which serialises every launch
and cannot afford the ~2,500 kernels x several steps of the hierarchical B = 4096 step.  Not a bench: tracing adds overhead.

    python tools/step_kernels.py [--workload hierarchical_b4096] [--steps 2] > gpurun_out/step_kernels.md
"""
import argparse
import importlib
import os
import re
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--no-graph", action="store_true")
a = ap.parse_args()
args = argparse.Namespace(batch=a.batch, chunk=a.chunk, no_graph=a.no_graph, dropout=a.dropout, warmup=3, fixed_warmup=True, workload=a.workload, steps=a.steps)
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
run = bench.Runner(args, a.workload, 1, 0, dev)
run.warmup()
ms, host_ms, launches = run.device_leg(a.steps)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        run.step(run.resident)
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
t_min, t_max = None, None
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    name = re.sub(r"\(.*", "", ev.name)
    name = re.sub(r"^void ", "", name).replace("(anonymous namespace)::", "")
    dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    agg[name][0] += 1
    agg[name][1] += dur
    s0 = ev.time_range.start
    t_min = s0 if t_min is None else min(t_min, s0)
    t_max = max(t_max or 0, ev.time_range.end)
total = sum(v[1] for v in agg.values())
print(f"# kernel time per step: {a.workload} (batch {run.batch}), {a.steps} traced steps, issue = {run.issue}\n")
print(f"untraced step {ms:.2f} ms; traced span {(t_max - t_min) / 1e3 / a.steps:.2f} ms/step; sum of kernel + memcpy/memset time {total / 1e3 / a.steps:.2f} ms/step "
      f"(GPU idle inside the traced span: {((t_max - t_min) - total) / 1e3 / a.steps:.2f} ms/step)\n")
print("| kernel | launches/step | ms/step | share | mean us |")
print("|---|---:|---:|---:|---:|")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"| `{n[:100]}` | {c / a.steps:.0f} | {us / 1e3 / a.steps:.3f} | {100 * us / total:.1f}% | {us / c:.1f} |")
