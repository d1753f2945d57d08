"""Scratch: MulT B=256 at 512/512/30 -- input gradients of a 64-sample-chunk eager run vs the single-chunk run must be bit-equal.
Repeats with several library debug switches to localise a mismatch.  python tools/repro_chunk_invariance.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("simple-multimodal_b200")
from oracle import fusion_oracle as fo          # noqa: E402
from parity_util import Cfg                      # noqa: E402

FL = pkg.fusion_layers
lib = pkg._lib.lib()
H, B, LENS = 512, 256, (512, 512, 30)
P = {k: v.to(torch.bfloat16).float() for k, v in fo.init_params("mult", H=H, heads=8, seed=5).items()}
g = torch.Generator(device="cpu").manual_seed(77)
host = [torch.randn((B, L, H), generator=g).to(torch.bfloat16) for L in LENS]


def run(chunk, graphs, stash=None):
    head = FL.MultimodalTransformer(Cfg()).cuda()
    head.load_state_dict(P, strict=True)
    head.train()
    head.chunk_size, head.graph_chunks = chunk, graphs
    if stash is not None:
        head.stash_fraction = stash
    xs = [h.cuda().requires_grad_(True) for h in host]
    out = head(*xs)
    ((out["fused_features"].float() ** 2).sum() / (B * H)).backward()
    torch.cuda.synchronize()
    r = (out["fused_features"].detach().clone(), [x.grad.clone() for x in xs])
    head.release_graphs()
    return r


def cmp(a, b):
    return [bool(torch.equal(a[0], b[0]))] + [int((x != y).sum()) for x, y in zip(a[1], b[1])]


for name, knobs in (("default", {}), ("late_aux", {12: 1}), ("no_tma_aux", {11: 1}), ("narrow_v1", {10: 2}), ("no_narrow", {10: 0})):
    for k, v in knobs.items():
        lib.b200f_debug_set(k, v)
    ref = run(1024, True)
    print(name, "engine vs engine   ", cmp(ref, run(1024, True)), flush=True)
    print(name, "engine vs eager 256", cmp(ref, run(256, False)), flush=True)
    print(name, "engine vs eager 64 ", cmp(ref, run(64, False)), cmp(ref, run(64, False, 1e-9)), flush=True)
    for k in knobs:
        lib.b200f_debug_set(k, 1 if k == 10 else 0)
