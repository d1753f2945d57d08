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
    if chunk is not None:
        (head.mult_fusion if kind == "hierarchical" else head).chunk_size = chunk
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


def compare(kind, cfg, B, lens, dtype, flag=False, mask=None, chunk=None, seed=7, report=None):
    P = fo.init_params(kind, H=cfg.fusion_hidden_size, heads=cfg.fusion_num_heads, graph_hidden=cfg.graph_hidden_size,
                       graph_layers=cfg.graph_num_layers, seed=seed)
    feats = fo.synthetic_features(B, lens, H=cfg.fusion_hidden_size, seed=1234)
    if dtype == torch.bfloat16:                       # identical inputs for both sides: round once, up front
        feats = tuple(f.to(torch.bfloat16).float() for f in feats)
    o_out, o_loss, o_xg, o_pg = oracle_run(kind, cfg, P, feats, flag, mask)
    d_out, d_loss, d_xg, d_pg = device_run(kind, cfg, P, feats, flag, dtype, mask, chunk)
    tol = TOL[dtype]
    errs = {}
    if isinstance(o_out, torch.Tensor):
        errs["out"] = rel(d_out, o_out)
    else:
        assert set(d_out) == set(o_out), (set(d_out), set(o_out))
        for k, v in o_out.items():
            if k == "contrastive_losses":
                assert set(d_out[k]) == set(v)
                for n in v:
                    errs[f"loss.{n}"] = abs(float(d_out[k][n]) - float(v[n]))
            else:
                errs[f"out.{k}"] = rel(d_out[k], v)
    errs["objective"] = abs(float(d_loss) - float(o_loss))
    for i, (g, r) in enumerate(zip(d_xg, o_xg)):
        errs[f"dx{i}"] = rel(g, r)
    assert set(d_pg) == set(o_pg), set(d_pg) ^ set(o_pg)
    for k, r in o_pg.items():
        assert d_pg[k] is not None, f"no gradient for {k}"
        errs[f"dP.{k}"] = rel(d_pg[k], r)
    if report is not None:
        report.update(errs)
    bad = {k: v for k, v in errs.items()
           if v > (tol["loss"] if k.startswith("loss.") or k == "objective" else tol["rel"])}
    assert not bad, f"{kind} {dtype}: out of tolerance: {bad}"
    return errs
