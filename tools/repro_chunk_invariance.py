"""Scratch: MulT B=256 at 512/512/30 -- outputs / input gradients of eagerly issued chunked runs vs the single-chunk engine run must be
bit-equal.  Repeats and prints WHERE they differ.  python tools/repro_chunk_invariance.py"""
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
    feats = torch.cat([out["text_features"], out["audio_features"], out["video_features"]], 1).detach().clone()
    ((out["fused_features"].float() ** 2).sum() / (B * H)).backward()
    torch.cuda.synchronize()
    r = (out["fused_features"].detach().clone(), [x.grad.clone() for x in xs], feats)
    head.release_graphs()
    return r


def where(a, b, name):
    d = (a.float() - b.float()).abs()
    if float(d.max()) == 0:
        return f"{name}: equal"
    rows = (d.reshape(d.size(0), -1).max(1).values > 0).nonzero().flatten().tolist()
    return f"{name}: {int((d > 0).sum())} elements differ, max abs {float(d.max()):.3e} (ref max {float(a.float().abs().max()):.3e}), samples {rows[:12]}{'...' if len(rows) > 12 else ''} ({len(rows)} samples)"


for kv in os.environ.get("KNOBS", "").split(","):
    if kv:
        k_, v_ = kv.split("=")
        lib.b200f_debug_set(int(k_), int(v_))
print("knobs:", os.environ.get("KNOBS", "(default)"), flush=True)
ref = run(1024, True)
for trial in range(int(os.environ.get("TRIALS", "8"))):
    r = run(64, False)
    msgs = [where(ref[0], r[0], "fused"), where(ref[2][:, :H], r[2][:, :H], "pooled_text"), where(ref[2][:, H:2 * H], r[2][:, H:2 * H], "pooled_audio"),
            where(ref[2][:, 2 * H:], r[2][:, 2 * H:], "pooled_video")] + [where(a, b, f"dx{i}") for i, (a, b) in enumerate(zip(ref[1], r[1]))]
    print(f"trial {trial}: " + ("ALL EQUAL" if all(m.endswith("equal") for m in msgs) else " | ".join(m for m in msgs if not m.endswith("equal"))), flush=True)
