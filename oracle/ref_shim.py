"""Load the *real* reference fusion module for pinning the oracle -- TEST INFRASTRUCTURE.

Only usable where /root/reference exists (the build container); never on the GPU box
and never from the product package.  `models/fusion_layers.py` imports
`torch_geometric` at module top (fusion_layers.py:4-5), which is not installed and cannot
be (no network), so a stand-in package is placed in `sys.modules` first.  Every class
except `GraphFusion` then runs as the reference's own unmodified code.

For `GraphFusion` the stand-ins are an *edge-list* restatement of the published
GATConv algorithm (PyG 2.3-2.6 semantics) plus minimal `Data`, `Batch`
and `global_mean_pool`; the reference's own wrapper code (type embeddings, per-sample
loop, ReLU, pooling, projection -- fusion_layers.py:240-291) executes around them.
It is deliberately written differently from the dense `gat_layer` in
`fusion_oracle.py` so the two restatements cross-check each other, but neither is the
real PyG: GATConv parity stays UNPINNED.
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("B200F_REFERENCE_ROOT", "/root/reference")


class _EdgeListGATConv(nn.Module):
    """GATConv(in, out, heads, concat=False, dropout) restated over an edge list."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, bias=True):
        super().__init__()
        assert not concat and add_self_loops and bias
        self.heads, self.out_channels, self.slope, self.p = heads, out_channels, negative_slope, dropout
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        bw = math.sqrt(6.0 / (in_channels + heads * out_channels))
        nn.init.uniform_(self.lin.weight, -bw, bw)
        ba = math.sqrt(6.0 / (heads + out_channels))
        nn.init.uniform_(self.att_src, -ba, ba)
        nn.init.uniform_(self.att_dst, -ba, ba)

    def forward(self, x, edge_index):
        n = x.size(0)
        src, dst = edge_index[0], edge_index[1]
        keep = src != dst                                    # remove_self_loops
        loops = torch.arange(n, device=x.device)
        src = torch.cat([src[keep], loops])                  # add_self_loops
        dst = torch.cat([dst[keep], loops])
        xp = self.lin(x).view(n, self.heads, self.out_channels)
        a_s = (xp * self.att_src).sum(-1)
        a_d = (xp * self.att_dst).sum(-1)
        e = torch.nn.functional.leaky_relu(a_s[src] + a_d[dst], self.slope)       # [E,h]
        emax = torch.full((n, self.heads), -float("inf"), dtype=e.dtype).scatter_reduce(
            0, dst[:, None].expand_as(e), e, reduce="amax")
        w = torch.exp(e - emax[dst])
        denom = torch.zeros(n, self.heads, dtype=e.dtype).index_add(0, dst, w)
        alpha = w / denom[dst]
        alpha = torch.nn.functional.dropout(alpha, self.p, self.training)
        out = torch.zeros(n, self.heads, self.out_channels, dtype=x.dtype).index_add(
            0, dst, alpha.unsqueeze(-1) * xp[src])
        return out.mean(dim=1) + self.bias


class _Data:
    def __init__(self, x=None, edge_index=None):
        self.x, self.edge_index = x, edge_index


class _Batch:
    @staticmethod
    def from_data_list(items):
        out = _Batch()
        xs, eis, bs, off = [], [], [], 0
        for i, d in enumerate(items):
            xs.append(d.x)
            eis.append(d.edge_index + off)
            bs.append(torch.full((d.x.size(0),), i, dtype=torch.long))
            off += d.x.size(0)
        out.x, out.edge_index, out.batch = torch.cat(xs, 0), torch.cat(eis, 1), torch.cat(bs)
        return out


def _global_mean_pool(x, batch):
    n = int(batch.max()) + 1
    s = torch.zeros(n, x.size(1), dtype=x.dtype).index_add(0, batch, x)
    c = torch.zeros(n, dtype=x.dtype).index_add(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return s / c[:, None]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "fusion_layers.py"))


_cached = None


def load_reference_fusion():
    """Return the reference `models/fusion_layers.py` as a module object (executed from where
    it lies under /root/reference; nothing is copied)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(f"reference not present under {REFERENCE_ROOT}")
    saved = {k: sys.modules.get(k) for k in ("torch_geometric", "torch_geometric.nn", "torch_geometric.data")}
    tg, tgn, tgd = (types.ModuleType(n) for n in ("torch_geometric", "torch_geometric.nn", "torch_geometric.data"))
    tgn.GATConv, tgn.global_mean_pool = _EdgeListGATConv, _global_mean_pool
    tgd.Data, tgd.Batch = _Data, _Batch
    tg.nn, tg.data = tgn, tgd
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tgn, "torch_geometric.data": tgd})
    try:
        spec = importlib.util.spec_from_file_location(
            "_reference_fusion_layers", os.path.join(REFERENCE_ROOT, "models", "fusion_layers.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached = mod
    return mod


class RefConfig:
    """Duck-typed stand-in for config.ModelConfig (config.py:6-79) -- only the attributes the
    fusion path reads.  Importing the real config.py would mkdir in CWD."""

    def __init__(self, H=512, heads=8, dropout=0.0, num_emotions=7, graph_hidden=512, graph_layers=3,
                 temperature=0.07):
        self.fusion_hidden_size = H
        self.fusion_num_heads = heads
        self.fusion_dropout = dropout
        self.num_emotions = num_emotions
        self.graph_hidden_size = graph_hidden
        self.graph_num_layers = graph_layers
        self.graph_dropout = dropout
        self.contrastive_temperature = temperature


REF_CLASS = {"early": "EarlyFusion", "late": "LateFusion", "mult": "MultimodalTransformer",
             "graph": "GraphFusion", "contrastive": "ContrastiveFusion", "adaptive": "AdaptiveFusion",
             "hierarchical": "HierarchicalFusion"}


def build_reference_head(kind: str, cfg: RefConfig):
    return getattr(load_reference_fusion(), REF_CLASS[kind])(cfg)
