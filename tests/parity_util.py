"""Shared helpers for the GPU parity tests: run a b200 head and the CPU oracle on identical seeded
features + weights and compare outputs, loss, input grads and parameter grads norm-wise."""
import importlib

import torch

from oracle import fusion_oracle as fo

pkg = importlib.import_module("simple-multimodal_b200")
FL = pkg.fusion_layers

CLS = {"early": "EarlyFusion", "late": "LateFusion", "mult": "MultimodalTransformer", "graph": "GraphFusion",
       "contrastive": "ContrastiveFusion", "adaptive": "AdaptiveFusion", "hierarchical": "HierarchicalFusion"}
# tolerances from BASELINE.json north_star, applied norm-wise per tensor against the fp64 oracle
TOL = {torch.float32: dict(rel=1e-5, loss=1e-5), torch.bfloat16: dict(rel=2e-2, loss=1e-3)}


class Cfg:
    def __init__(self, H=512, heads=8, graph_hidden=512, graph_layers=3, dropout=0.0, temperature=0.07, num_emotions=7):
        self.fusion_hidden_size, self.fusion_num_heads, self.fusion_dropout = H, heads, dropout
        self.num_emotions, self.graph_hidden_size, self.graph_num_layers = num_emotions, graph_hidden, graph_layers
        self.graph_dropout, self.contrastive_temperature = dropout, temperature


def rel(x, ref):
    x, ref = x.detach().double().cpu(), ref.detach().double().cpu()
    return float((x - ref).norm() / ref.norm().clamp_min(1e-30))


def oracle_run(kind, cfg, P, feats, flag, mask=None):
    P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
    xs = [f.double().requires_grad_(True) for f in feats]
    kw = {}
    if kind in ("mult", "adaptive", "hierarchical"):
        kw["heads"] = cfg.fusion_num_heads
    if kind in ("contrastive", "hierarchical"):
        kw["compute_contrastive_loss"] = flag
        kw["temperature"] = cfg.contrastive_temperature
    if kind == "graph":
        kw["num_layers"] = cfg.graph_num_layers
    if kind == "hierarchical":
        kw["graph_layers"] = cfg.graph_num_layers
    xin = fo.apply_modality_mask(*xs, mask.double() if mask is not None else None)
    out = fo.HEADS[kind](*xin, P64, **kw)
    loss = fo.objective(out)
    loss.backward()
    return out, loss, [x.grad for x in xs], {k: v.grad for k, v in P64.items()}


def device_run(kind, cfg, P, feats, flag, dtype, mask=None, chunk=None):
    head = getattr(FL, CLS[kind])(cfg).cuda()
    head.load_state_dict(P, strict=True)
    head.train()
    if chunk is not None:                      # small chunks + a one-chunk stash budget: exercises resident AND recomputed chunks
        mt = head.mult_fusion if kind == "hierarchical" else head
        mt.chunk_size = chunk
        mt.stash_fraction = 1e-9 if chunk == 2 else mt.stash_fraction
    xs = [f.to("cuda", dtype).requires_grad_(True) for f in feats]
    kw = {}
    if kind in ("contrastive", "hierarchical"):
        kw["compute_contrastive_loss"] = flag
    if mask is not None:
        kw["mask"] = mask.cuda()
    out = head(*xs, **kw)
    loss = fo.objective(out)
    loss.backward()
    torch.cuda.synchronize()
    return out, loss, [x.grad for x in xs], {k: p.grad for k, p in head.named_parameters()}


import json
import os

_FLOOR_PATH = os.path.join(os.path.dirname(__file__), "golden", "bf16_floor.json")
BF16_FLOOR = json.load(open(_FLOOR_PATH)) if os.path.exists(_FLOOR_PATH) else {}
GRAD_CAP = 0.10       # no bf16 gradient tensor may be further than 10 % (norm-wise) from the fp64 oracle, whatever the reference's own
                      # bf16 deviation is; the measured errors of every tensor are in profiles/r02_bf16_parity_errors.json
ZERO_FLOOR = 1e-4     # a gradient tensor whose oracle norm is < 1e-4 of the largest one is "numerically zero":
                      # its error is measured against that floor instead of its own (cancellation-level) norm


def _group_errors(dev, ref, prefix):
    """norm-wise error per tensor, denominators floored at ZERO_FLOOR * (largest oracle norm of the group)."""
    norms = {k: float(r.detach().double().norm()) for k, r in ref.items()}
    top = max(norms.values()) if norms else 0.0
    out = {}
    for k, r in ref.items():
        d = float((dev[k].detach().double().cpu() - r.detach().double()).norm())
        out[f"{prefix}{k}"] = d / max(norms[k], ZERO_FLOOR * top, 1e-30)
    return out


def compare(kind, cfg, B, lens, dtype, flag=False, mask=None, chunk=None, seed=7, case_id=None):
    """Run the CUDA head and the fp64 oracle on identical data and check every tensor.
    fp32: 1e-5 norm-wise everywhere.  bf16: outputs 2e-2, losses 1e-3 (north_star); gradients 2e-2 or, where the
    reference's OWN bf16 autocast execution deviates more than 1e-2 from fp64 (tests/golden/bf16_floor.json, made by
    oracle/make_bf16_floor.py), within 2x of that measured floor and never more than GRAD_CAP = 10 %.  Tensors whose oracle
    norm is cancellation-level (< ZERO_FLOOR of the group's largest) are measured against that floor, not their own norm."""
    P = fo.init_params(kind, H=cfg.fusion_hidden_size, heads=cfg.fusion_num_heads, graph_hidden=cfg.graph_hidden_size,
                       graph_layers=cfg.graph_num_layers, seed=seed)
    feats = fo.synthetic_features(B, lens, H=cfg.fusion_hidden_size, seed=1234)
    if dtype == torch.bfloat16:
        # identical inputs AND weights on both sides: make them bf16-representable once, up front, so the only
        # difference left is the arithmetic (bf16 storage of intermediates, fp32 accumulation) -- not the data
        feats = tuple(f.to(torch.bfloat16).float() for f in feats)
        P = {k: v.to(torch.bfloat16).float() for k, v in P.items()}
    o_out, o_loss, o_xg, o_pg = oracle_run(kind, cfg, P, feats, flag, mask)
    d_out, d_loss, d_xg, d_pg = device_run(kind, cfg, P, feats, flag, dtype, mask, chunk)
    tol = TOL[dtype]
    floor = BF16_FLOOR.get(case_id, {}) if dtype == torch.bfloat16 else {}
    errs, limits = {}, {}
    if isinstance(o_out, torch.Tensor):
        errs["out"] = rel(d_out, o_out)
    else:
        assert set(d_out) == set(o_out), (set(d_out), set(o_out))
        for k, v in o_out.items():
            if k == "contrastive_losses":
                assert set(d_out[k]) == set(v)
                for n in v:
                    errs[f"loss.{n}"] = abs(float(d_out[k][n].detach()) - float(v[n].detach()))
            else:
                errs[f"out.{k}"] = rel(d_out[k], v)
    errs["objective"] = abs(float(d_loss.detach()) - float(o_loss.detach()))
    errs.update(_group_errors({f"{i}": g for i, g in enumerate(d_xg)}, {f"{i}": g for i, g in enumerate(o_xg)}, "dx"))
    assert set(d_pg) == set(o_pg), set(d_pg) ^ set(o_pg)
    for k in o_pg:
        assert d_pg[k] is not None, f"no gradient for {k}"
    errs.update(_group_errors(d_pg, o_pg, "dP."))
    for k in errs:
        if k.startswith("loss.") or k == "objective":
            limits[k] = tol["loss"]
        elif k.startswith("out"):
            limits[k] = tol["rel"]
        else:
            # the six MulT blocks (and the three projectors ...) are statistically identical: use the largest measured
            # reference-autocast deviation among tensors with the same role (same trailing name) as this tensor's floor
            role = ".".join(k.split(".")[-2:])
            fl = max([v for n, v in floor.items() if n.endswith(role) and n.split(".")[0] == k.split(".")[0]] + [0.0])
            limits[k] = max(tol["rel"], min(2.0 * fl, GRAD_CAP))
    table = os.environ.get("B200F_PARITY_TABLE")              # optional dump of every measured error next to its limit (JSON lines)
    if table:
        with open(table, "a") as f:
            f.write(json.dumps({"case": case_id, "kind": kind, "dtype": str(dtype).replace("torch.", ""),
                                "errors": {k: [float(v), float(limits[k])] for k, v in errs.items()}}) + "\n")
    bad = {k: (v, limits[k]) for k, v in errs.items() if not v <= limits[k]}
    worst = dict(sorted(((k, round(v / limits[k], 3)) for k, v in errs.items()), key=lambda kv: -kv[1])[:8])
    assert not bad, f"{kind} {dtype}: {len(bad)} tensors out of tolerance (err, limit): {dict(list(bad.items())[:10])}; worst err/limit: {worst}"
    return errs
