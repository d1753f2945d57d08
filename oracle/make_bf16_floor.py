"""Measure the bf16 noise floor of the REFERENCE ITSELF -- TEST INFRASTRUCTURE (build container only).

For every case of tests/parity_cases.py the executed reference (via ref_shim) runs once in float64 and once
under `torch.autocast("cpu", dtype=torch.bfloat16)` on the same bf16-representable weights and features; the
norm-wise deviation of every output / input-gradient / parameter-gradient tensor is written to
tests/golden/bf16_floor.json.  tests/test_parity_gpu.py uses it to bound the CUDA bf16 path: a gradient tensor
passes if it is within north_star's 2e-2 of the fp64 oracle OR within 2x of what the reference's own bf16
execution deviates (sign flips of near-zero ReLU pre-activations put a ~sqrt(flip fraction) floor of a few per
cent on any bf16 implementation's gradients behind a hidden ReLU -- SURVEY F7).
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import fusion_oracle as fo          # noqa: E402
from oracle import ref_shim                      # noqa: E402
from parity_cases import CASES                   # noqa: E402


def run(kind, head, xs, flag, autocast):
    import contextlib
    ctx = torch.autocast("cpu", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with ctx:
        if kind == "hierarchical" and xs[0].dim() == 3:
            p2 = [x.mean(dim=1) for x in xs]
            early, mult = head.early_fusion(*p2), head.mult_fusion(*xs)["fused_features"]
            graph, con, ada = head.graph_fusion(*p2), head.contrastive_fusion(*p2, flag), head.adaptive_fusion(*p2)
            fused = head.meta_fusion(torch.cat([early, mult, graph, con["fused_features"], ada["fused_features"]], dim=-1))
            out = {"fused_features": fused, "early_features": early, "mult_features": mult, "graph_features": graph,
                   "contrastive_features": con["fused_features"], "adaptive_features": ada["fused_features"],
                   "contrastive_losses": con["contrastive_losses"], "attention_weights": ada["attention_weights"],
                   "adaptive_weights": ada["adaptive_weights"]}
        elif kind in ("contrastive", "hierarchical"):
            out = head(*xs, compute_contrastive_loss=flag)
        else:
            out = head(*xs)
    loss = fo.objective(out)
    loss.backward()
    return out, loss


def rel(x, r):
    return float((x.double() - r.double()).norm() / r.double().norm().clamp_min(1e-30))


def main():
    floors = {}
    for cid, c in CASES.items():
        if c.get("fp32_only"):
            continue
        kind = c["kind"]
        cfgkw = dict(H=512, heads=8, graph_hidden=512, graph_layers=3)
        cfgkw.update(c["cfg"])
        cfg = ref_shim.RefConfig(**cfgkw)
        P = fo.init_params(kind, H=512, heads=8, graph_hidden=cfgkw["graph_hidden"], graph_layers=cfgkw["graph_layers"], seed=7)
        P = {k: v.to(torch.bfloat16).float() for k, v in P.items()}
        feats = [f.to(torch.bfloat16).float() for f in fo.synthetic_features(c["B"], c["lens"], H=512, seed=1234)]
        mask = fo.modality_keep_mask(c["B"], 0.4, torch.Generator().manual_seed(c["mask_seed"])) if c["mask_seed"] else None
        res = []
        for autocast in (False, True):
            head = ref_shim.build_reference_head(kind, cfg)
            head = head.float() if autocast else head.double()
            head.load_state_dict({k: (v.float() if autocast else v.double()) for k, v in P.items()}, strict=True)
            head.train()
            xs = [(f.float() if autocast else f.double()).clone().requires_grad_(True) for f in feats]
            xin = fo.apply_modality_mask(*xs, mask.to(xs[0].dtype) if mask is not None else None)
            out, loss = run(kind, head, list(xin), c["flag"], autocast)
            res.append((out, loss, [x.grad for x in xs], {k: p.grad for k, p in head.named_parameters()}))
        (o0, l0, xg0, pg0), (o1, l1, xg1, pg1) = res
        f = {"objective": abs(float(l1.detach()) - float(l0.detach()))}
        if isinstance(o0, torch.Tensor):
            f["out"] = rel(o1.detach(), o0.detach())
        else:
            for k, v in o0.items():
                if k == "contrastive_losses":
                    for n in v:
                        f[f"loss.{n}"] = abs(float(o1[k][n].detach()) - float(v[n].detach()))
                elif isinstance(v, torch.Tensor):
                    f[f"out.{k}"] = rel(o1[k].detach(), v.detach())
        for i, (a, b) in enumerate(zip(xg1, xg0)):
            f[f"dx{i}"] = rel(a, b)
        for k in pg0:
            if pg0[k] is not None and pg1[k] is not None:
                f[f"dP.{k}"] = rel(pg1[k], pg0[k])
        floors[cid] = f
        worst = sorted(f.items(), key=lambda kv: -kv[1])[:3]
        print(cid, [(k, f"{v:.3g}") for k, v in worst], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "bf16_floor.json"), "w") as fh:
        json.dump(floors, fh, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
