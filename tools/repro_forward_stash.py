"""Scratch: run mult_engine._chunk_forward repeatedly on the same chunk and report the first stashed tensor (in schedule order) that is
not bit-identical to the first run's.  python tools/repro_forward_stash.py"""
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
H, B, LENS = 512, int(os.environ.get("BATCH", "64")), (512, 512, 30)
P = {k: v.to(torch.bfloat16).float() for k, v in fo.init_params("mult", H=H, heads=8, seed=5).items()}
g = torch.Generator(device="cpu").manual_seed(77)
xs = [torch.randn((B, L, H), generator=g).to(torch.bfloat16).cuda() for L in LENS]
head = FL.MultimodalTransformer(Cfg()).cuda()
head.load_state_dict(P, strict=True)
W, plist = head._operands(torch.bfloat16)
ORDER = [n for n, _, _ in ME.BLOCKS] + list(ME.MODS)
FIELDS = {"blk": ["ctx", "lse", "s1", "x1", "hid", "s2"], "mod": ["enh", "qkv", "att", "lse", "pooled_ctx"]}


def flat(st, pooled):
    out = [("proj%d" % m, t) for m, t in enumerate(st["proj"])]
    for n in ORDER:
        for f in FIELDS["blk" if n in [b for b, _, _ in ME.BLOCKS] else "mod"]:
            out.append((f"{n}.{f}", st[n][f]))
    out.append(("pooled_out", pooled))
    return out


def once():
    pooled = torch.empty((B, 3 * H), device="cuda", dtype=torch.bfloat16)
    st = ME._chunk_forward(xs, W, H, 8, pooled, keep=True)
    torch.cuda.synchronize()
    return flat(st, pooled)


ref = once()
keep_alive = []
n_bad = 0
for it in range(int(os.environ.get("TRIALS", "60"))):
    cur = once()
    if os.environ.get("HOLD", "1") == "1":
        keep_alive.append(cur)                     # like a resident stash: every run gets fresh memory
        if len(keep_alive) > 6:
            keep_alive.pop(0)
    bad = [(n, a, b) for (n, a), (_, b) in zip(ref, cur) if not torch.equal(a, b)]
    if bad:
        n_bad += 1
        n, a, b = bad[0]
        d = (a.float() - b.float()).abs().reshape(a.size(0), -1)
        rows = (d.max(1).values > 0).nonzero().flatten().tolist()
        sub = d[rows[0]].reshape(-1, a.size(-1)) if a.dim() == 3 else d[rows[0]].reshape(1, -1)
        toks = (sub.max(1).values > 0).nonzero().flatten().tolist()
        cols = (sub.max(0).values > 0).nonzero().flatten().tolist()
        print(f"it {it}: first differing tensor {n} {tuple(a.shape)}: samples {rows[:8]} tokens {toks[:6]}..{toks[-3:]} ({len(toks)}) cols {cols[:4]}..{cols[-3:]} ({len(cols)}) "
              f"max {float(d.max()):.3e}; then {[x[0] for x in bad[1:4]]}", flush=True)
print(f"{n_bad} of {it + 1} runs differ; knobs {os.environ.get('KNOBS', '')}")
