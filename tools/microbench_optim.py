"""GPU microbenchmark (not a pytest) of the fused optimizer step on the parameter set of the hierarchical fusion
head (graph 512/512/3: ~36 M fp32 parameters in ~140 tensors): FusedAdamW.clip_grad_norm_ + step next to
torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW (foreach and fused=True).  Reports ms per optimizer step and the achieved GB/s
against the algorithmic traffic (4 B/element for the norm + 28 B/element for the update).  Writes gpurun_out/microbench_optim.json.

    python tools/microbench_optim.py [--iters 20]
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402  (Cfg of the bench workloads)

pkg = importlib.import_module("simple-multimodal_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda")


def params():
    torch.manual_seed(0)
    head = pkg.HierarchicalFusion(bench.Cfg).to(dev)
    ps = [p for p in head.parameters()]
    g = torch.Generator(device="cuda").manual_seed(1)
    for p in ps:
        p.grad = torch.randn(p.shape, device=dev, generator=g) * 0.01
    return ps


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
ps = params()
n = sum(p.numel() for p in ps)
out["tensors"], out["elements"] = len(ps), n
out["algorithmic_bytes_per_step"] = 32 * n
ours = pkg.FusedAdamW(ps, lr=1e-4, weight_decay=1e-5)


def step_ours():
    ours.clip_grad_norm_(1.0)
    ours.step()


l0 = pkg._lib.launch_count()
step_ours()
out["fused_launches_per_step"] = pkg._lib.launch_count() - l0
ms = timeit(step_ours)
out["FusedAdamW"] = {"ms": ms, "GBps": 32 * n / ms / 1e6}
for label, kw in (("torch_foreach", {"foreach": True}), ("torch_fused", {"fused": True})):
    ps2 = params()
    ref = torch.optim.AdamW(ps2, lr=1e-4, weight_decay=1e-5, **kw)

    def step_ref():
        torch.nn.utils.clip_grad_norm_(ps2, 1.0)
        ref.step()

    ms = timeit(step_ref)
    out[label] = {"ms": ms, "GBps": 32 * n / ms / 1e6}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "microbench_optim.json"), "w") as f:
    json.dump(out, f, indent=1)
for k, v in out.items():
    print(k, v)
