"""Generate tests/golden/*.pt from the EXECUTED reference -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
Each case: parameters from `fusion_oracle.init_params(seed)` are loaded with
`load_state_dict(strict=True)` into the reference class (which also pins names and shapes),
the reference runs forward + backward in float64 with dropout 0 on seeded synthetic
features, and inputs' seeds, outputs, input-grads and parameter-grads (full at H=64,
digests at H=512 to keep the fixtures small) are stored.  The 3-D hierarchical case is
the reference's own sub-modules composed with one explicit mean-pool.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import fusion_oracle as fo          # noqa: E402
from oracle import ref_shim                      # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name, kind, H, heads, graph_hidden, B, lens, contrastive flag, store full param grads
CASES = [
    ("early_h64", "early", 64, 8, 64, 5, (None, None, None), False, True),
    ("late_h64", "late", 64, 8, 64, 5, (None, None, None), False, True),
    ("mult2d_h32", "mult", 32, 4, 32, 5, (None, None, None), False, True),
    ("mult3d_h32", "mult", 32, 4, 32, 3, (7, 9, 4), False, True),
    ("graph_h64", "graph", 64, 8, 64, 5, (None, None, None), False, True),
    ("graph_1layer_h64", "graph1", 64, 8, 32, 5, (None, None, None), False, True),
    ("contrastive_h64", "contrastive", 64, 8, 64, 6, (None, None, None), True, True),
    ("adaptive_h64", "adaptive", 64, 8, 64, 5, (None, None, None), False, True),
    ("hier2d_h32", "hierarchical", 32, 4, 32, 4, (None, None, None), True, True),
    ("hier3d_h32", "hierarchical", 32, 4, 32, 3, (6, 5, 3), True, True),
    ("early_h512_b16", "early", 512, 8, 512, 16, (None, None, None), False, False),   # BASELINE config 1
    ("mult3d_h512", "mult", 512, 8, 512, 2, (33, 65, 7), False, False),
    ("contrastive_h512", "contrastive", 512, 8, 512, 64, (None, None, None), True, False),
]


def digest(t: torch.Tensor):
    f = t.detach().double().flatten()
    return torch.cat([f.sum()[None], f.norm()[None], f[:4] if f.numel() >= 4 else torch.nn.functional.pad(f, (0, 4 - f.numel()))])


def run_reference(kind, cfg, P, feats, flag):
    real_kind = "graph" if kind == "graph1" else kind
    head = ref_shim.build_reference_head(real_kind, cfg).double()
    missing = head.load_state_dict({k: v.double() for k, v in P.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    head.train()                                    # dropout p=0: train == eval numerically
    xs = [f.double().clone().requires_grad_(True) for f in feats]
    if real_kind == "hierarchical" and xs[0].dim() == 3:
        # reference sub-modules + one explicit pooling step (fusion_layers.py:486-508)
        p2 = [x.mean(dim=1) for x in xs]
        early = head.early_fusion(*p2)
        mult = head.mult_fusion(*xs)["fused_features"]
        graph = head.graph_fusion(*p2)
        con = head.contrastive_fusion(*p2, flag)
        ada = head.adaptive_fusion(*p2)
        fused = head.meta_fusion(torch.cat([early, mult, graph, con["fused_features"], ada["fused_features"]], dim=-1))
        out = {"fused_features": fused, "early_features": early, "mult_features": mult, "graph_features": graph,
               "contrastive_features": con["fused_features"], "adaptive_features": ada["fused_features"],
               "contrastive_losses": con["contrastive_losses"], "attention_weights": ada["attention_weights"],
               "adaptive_weights": ada["adaptive_weights"]}
    elif real_kind in ("contrastive", "hierarchical"):
        out = head(*xs, compute_contrastive_loss=flag)
    else:
        out = head(*xs)
    loss = fo.objective(out)
    loss.backward()
    pgrads = {k: p.grad for k, p in head.named_parameters()}
    return out, loss, [x.grad for x in xs], pgrads


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    for name, kind, H, heads, gh, B, lens, flag, full in CASES:
        glayers = 1 if kind == "graph1" else 3
        real_kind = "graph" if kind == "graph1" else kind
        cfg = ref_shim.RefConfig(H=H, heads=heads, graph_hidden=gh, graph_layers=glayers)
        P = fo.init_params(real_kind, H=H, heads=heads, graph_hidden=gh, graph_layers=glayers, seed=7)
        feats = fo.synthetic_features(B, lens, H=H, seed=1234)
        out, loss, xg, pg = run_reference(kind, cfg, P, feats, flag)
        rec = {"meta": {"kind": real_kind, "H": H, "heads": heads, "graph_hidden": gh, "graph_layers": glayers,
                        "B": B, "lens": lens, "flag": flag, "param_seed": 7, "feat_seed": 1234, "full": full},
               "loss": loss.detach(), "input_grads": [g.detach() for g in xg], "outputs": {}, "losses": {}}
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


if __name__ == "__main__":
    main()
