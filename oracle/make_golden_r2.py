"""Second batch of golden vectors from the EXECUTED reference -- TEST INFRASTRUCTURE (build container only, needs /root/reference).

    python oracle/make_golden_r2.py            -> tests/golden/r2/*.pt

What round 1's fixtures (oracle/make_golden.py, left untouched so that they regenerate bit for bit) did not cover:
  * `CrossModalTransformer` on its own (reference models/fusion_layers.py:182-211) -- `cross_*`: forward(query, key_value) with
    different query / key lengths, outputs + input gradients + parameter gradients;
  * batch-size edge cases straight from the reference, B = 1 and B = 257 (ragged against every tile size);
  * MulT at the benchmark sequence lengths (512, 512, 30), H = 512, 8 heads, B = 1 -- gradients as digests plus sampled rows.
Parameters come from `fusion_oracle.init_params(seed)` (for `cross`: the `text_to_audio.` block of the MulT parameter set) and are
loaded with `load_state_dict(strict=True)` into the reference class; float64, dropout 0, seeded synthetic features.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import fusion_oracle as fo          # noqa: E402
from oracle import ref_shim                      # noqa: E402
from oracle.make_golden import digest, run_reference   # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden", "r2")
ROW_STRIDE = 61        # sampled rows of big input gradients: every 61st token

# name, kind, H, heads, graph_hidden, B, lens, contrastive flag, full parameter gradients
CASES = [
    ("early_h64_b1", "early", 64, 8, 64, 1, (None, None, None), False, True),
    ("early_h64_b257", "early", 64, 8, 64, 257, (None, None, None), False, True),
    ("late_h64_b1", "late", 64, 8, 64, 1, (None, None, None), False, True),
    ("adaptive_h64_b1", "adaptive", 64, 8, 64, 1, (None, None, None), False, True),
    ("adaptive_h64_b257", "adaptive", 64, 8, 64, 257, (None, None, None), False, True),
    ("contrastive_h64_b1", "contrastive", 64, 8, 64, 1, (None, None, None), True, True),
    ("contrastive_h64_b257", "contrastive", 64, 8, 64, 257, (None, None, None), True, True),
    ("graph_h64_b1", "graph", 64, 8, 64, 1, (None, None, None), False, True),
    ("mult2d_h32_b1", "mult", 32, 4, 32, 1, (None, None, None), False, True),
    ("mult3d_h32_b257", "mult", 32, 4, 32, 257, (3, 2, 4), False, True),
    ("hier2d_h32_b1", "hierarchical", 32, 4, 32, 1, (None, None, None), True, True),
    ("mult3d_h512_bench_b1", "mult", 512, 8, 512, 1, (512, 512, 30), False, False),       # the benchmark sequence lengths
]
# CrossModalTransformer alone: name, H, heads, B, Lq, Lk, full
CROSS = [
    ("cross_h64", 64, 8, 3, 7, 5, True),
    ("cross_h64_q1", 64, 8, 2, 1, 9, True),
    ("cross_h512", 512, 8, 2, 40, 30, False),
]


def cross_params(H, heads, seed=7):
    P = fo.init_params("mult", H=H, heads=heads, seed=seed)
    pre = "text_to_audio."
    return {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}


def cross_features(B, Lq, Lk, H, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, Lq, H, generator=g), torch.randn(B, Lk, H, generator=g)


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    for name, kind, H, heads, gh, B, lens, flag, full in CASES:
        cfg = ref_shim.RefConfig(H=H, heads=heads, graph_hidden=gh, graph_layers=3)
        P = fo.init_params(kind, H=H, heads=heads, graph_hidden=gh, graph_layers=3, seed=7)
        feats = fo.synthetic_features(B, lens, H=H, seed=1234)
        out, loss, xg, pg = run_reference(kind, cfg, P, feats, flag)
        big = not full
        rec = {"meta": {"kind": kind, "H": H, "heads": heads, "graph_hidden": gh, "graph_layers": 3, "B": B, "lens": lens, "flag": flag,
                        "param_seed": 7, "feat_seed": 1234, "full": full, "input_digest": big, "row_stride": ROW_STRIDE},
               "loss": loss.detach(), "outputs": {}, "losses": {}}
        if big:
            rec["input_grad_digest"] = [digest(g) for g in xg]
            rec["input_grad_rows"] = [g.detach()[:, ::ROW_STRIDE].clone() for g in xg]
        else:
            rec["input_grads"] = [g.detach() for g in xg]
        if isinstance(out, torch.Tensor):
            rec["outputs"]["__tensor__"] = out.detach()
        else:
            for k, v in out.items():
                if k == "contrastive_losses":
                    rec["losses"] = {n: t.detach() for n, t in v.items()}
                elif isinstance(v, torch.Tensor):
                    rec["outputs"][k] = v.detach()
        rec["param_grads"] = {k: (g.detach() if full else digest(g)) for k, g in pg.items() if g is not None}
        path = os.path.join(OUT_DIR, name + ".pt")
        torch.save(rec, path)
        print(f"{name}: loss={float(loss.detach()):.12f} -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")

    ref = ref_shim.load_reference_fusion()
    for name, H, heads, B, Lq, Lk, full in CROSS:
        cfg = ref_shim.RefConfig(H=H, heads=heads)
        P = cross_params(H, heads)
        block = ref.CrossModalTransformer(cfg).double()
        res = block.load_state_dict({k: v.double() for k, v in P.items()}, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        block.train()
        q, kv = (t.double().requires_grad_(True) for t in cross_features(B, Lq, Lk, H))
        out = block(q, kv)
        loss = fo.objective(out)
        loss.backward()
        rec = {"meta": {"kind": "cross", "H": H, "heads": heads, "B": B, "Lq": Lq, "Lk": Lk, "param_seed": 7, "feat_seed": 1234, "full": full},
               "loss": loss.detach(), "output": out.detach(), "input_grads": [q.grad.detach(), kv.grad.detach()],
               "param_grads": {k: (p.grad.detach() if full else digest(p.grad)) for k, p in block.named_parameters()}}
        path = os.path.join(OUT_DIR, name + ".pt")
        torch.save(rec, path)
        print(f"{name}: loss={float(loss.detach()):.12f} -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()
