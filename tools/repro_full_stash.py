"""Scratch: full eager fwd+bwd MulT runs (B=256, chunk 64, resident stash); compare every stashed forward tensor of every chunk with the
first run's and report the first one (in schedule order) that differs.  python tools/repro_full_stash.py"""
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

ME, FL = pkg.mult_engine, pkg.fusion_layers
lib = pkg._lib.lib()
for kv in os.environ.get("KNOBS", "").split(","):
    if kv:
        k_, v_ = kv.split("=")
        lib.b200f_debug_set(int(k_), int(v_))
H, B, LENS = 512, 256, (512, 512, 30)
P = {k: v.to(torch.bfloat16).float() for k, v in fo.init_params("mult", H=H, heads=8, seed=5).items()}
g = torch.Generator(device="cpu").manual_seed(77)
host = [torch.randn((B, L, H), generator=g).to(torch.bfloat16) for L in LENS]
BLK = [b for b, _, _ in ME.BLOCKS]
ORDER = BLK + list(ME.MODS)
FIELDS = {"blk": ["ctx", "lse", "s1", "x1", "mean1", "rstd1", "hid", "s2", "mean2", "rstd2"], "mod": ["enh", "qkv", "att", "lse", "pooled_ctx"]}
orig = ME._chunk_forward
rec = []


def patched(xs, W, H_, heads, pooled_out, keep, drop=None):
    st = orig(xs, W, H_, heads, pooled_out, keep, drop)
    if st is not None:
        out = [("proj%d" % m, t) for m, t in enumerate(st["proj"])]
        for n in ORDER:
            for f in FIELDS["blk" if n in BLK else "mod"]:
                out.append((f"{n}.{f}", st[n][f]))
        rec.append([(n, t.clone()) for n, t in out] + [("pooled_out", pooled_out.clone())])
    return st


ME._chunk_forward = patched


def run():
    rec.clear()
    head = FL.MultimodalTransformer(Cfg()).cuda()
    head.load_state_dict(P, strict=True)
    head.train()
    head.chunk_size, head.graph_chunks = 64, False
    xs = [h.cuda().requires_grad_(True) for h in host]
    out = head(*xs)
    ((out["fused_features"].float() ** 2).sum() / (B * H)).backward()
    torch.cuda.synchronize()
    return [list(c) for c in rec]


ref = run()
n_bad = 0
trials = int(os.environ.get("TRIALS", "40"))
for it in range(trials):
    cur = run()
    for ci, (rc, cc) in enumerate(zip(ref, cur)):
        bad = [(n, a, b) for (n, a), (_, b) in zip(rc, cc) if not torch.equal(a, b)]
        if bad:
            n_bad += 1
            n, a, b = bad[0]
            d = (a.float() - b.float()).abs().reshape(a.size(0), -1)
            rows = (d.max(1).values > 0).nonzero().flatten().tolist()
            inner = d[rows[0]].reshape(-1, a.size(-1)) if a.dim() >= 3 else d[rows[0]].reshape(1, -1)
            toks = (inner.max(1).values > 0).nonzero().flatten().tolist()
            cols = (inner.max(0).values > 0).nonzero().flatten().tolist()
            print(f"trial {it} chunk {ci}: FIRST differing tensor {n} {tuple(a.shape)}: rows(dim0) {rows[:6]} ({len(rows)}); inside row {rows[0]}: tokens {toks[:5]}..{toks[-2:]} ({len(toks)}) "
                  f"cols {cols[:3]}..{cols[-2:]} ({len(cols)}) max {float(d.max()):.3e}; next {[x[0] for x in bad[1:5]]}", flush=True)
print(f"{n_bad} chunk mismatches in {trials} trials; knobs {os.environ.get('KNOBS', '')}")
