"""CPU oracle for the fusion hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, explicit-math restatement (plain torch tensors on the CPU, fp32 or
fp64) of the reference's fusion heads.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this
file; the product package (`simple-multimodal_b200/`) never does, and fails
loudly when its CUDA library is missing instead of falling back to this.

Every function takes the activations plus a flat parameter dict `P` keyed by
the reference's own `state_dict()` names (with an optional dotted `prefix`), so
a reference checkpoint can be fed to it unchanged.  Nothing here calls
`nn.MultiheadAttention`, `nn.LayerNorm` or `F.cross_entropy`: the arithmetic is
written out so that the CUDA kernels can be checked against each intermediate.

Pinning status (see oracle/README.md and DESIGN.md):
  * every head except GraphFusion is pinned against the *executed* reference
    (`/root/reference/models/fusion_layers.py` imported through
    `oracle/ref_shim.py`) by `oracle/make_golden.py`; the resulting vectors are
    committed under `tests/golden/` and re-checked by `tests/test_oracle.py`.
  * GraphFusion depends on `torch_geometric.nn.GATConv`, a third-party
    dependency (`torch-geometric>=2.3.0`, un-pinned, `requirements.txt:34`) that
    is neither vendored under /root/reference nor installed: its published
    algorithm is restated in `gat_layer` below -- PARITY UNPINNED for that head.

Reference citations are `file:line` relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

LN_EPS = 1e-5          # nn.LayerNorm default, models/fusion_layers.py:192-193
L2_EPS = 1e-12         # F.normalize default, models/fusion_layers.py:338-340
GAT_HEADS = 4          # models/fusion_layers.py:227
GAT_SLOPE = 0.2        # GATConv default negative_slope
# Timing knob for bench.py's CPU arm ONLY: > 0 adds the reference's training-mode dropouts of the MulT blocks (attention
# probabilities, torch nn/functional.py:6647-6650; FFN hidden, models/fusion_layers.py:197-199) with torch's own RNG, so that
# the CPU baseline does the same work as the CUDA arm at fusion_dropout = 0.1.  Parity (tests/) always runs with 0.
TRAIN_DROPOUT = 0.0


def _p(P: Params, prefix: str, name: str) -> Tensor:
    return P[f"{prefix}{name}"]


# ----------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------
def affine(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """y = x W^T + b  (nn.Linear)."""
    y = torch.matmul(x, w.transpose(0, 1))
    return y if b is None else y + b


def layer_norm(x: Tensor, g: Tensor, b: Tensor) -> Tensor:
    """Biased-variance LayerNorm over the last dim, eps inside the sqrt
    (models/fusion_layers.py:192-193, 205, 209)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + LN_EPS) * g + b


def softmax_lastdim(s: Tensor) -> Tensor:
    m = s.max(dim=-1, keepdim=True).values
    e = torch.exp(s - m)
    return e / e.sum(dim=-1, keepdim=True)


def multi_head_attention(q_in: Tensor, kv_in: Tensor, P: Params, prefix: str,
                         heads: int) -> Tuple[Tensor, Tensor]:
    """nn.MultiheadAttention(batch_first=True), need_weights=True path, dropout off.

    Follows torch/nn/functional.py multi_head_attention_forward:
    packed in-projection, q scaled by 1/sqrt(head_dim) after the bias, softmax
    over keys, out-projection; returns (output [B,Lq,H], head-averaged weights
    [B,Lq,Lk]).  Call sites: models/fusion_layers.py:161-163, 204, 432-434.
    """
    w_in = _p(P, prefix, "in_proj_weight")
    b_in = _p(P, prefix, "in_proj_bias")
    w_out = _p(P, prefix, "out_proj.weight")
    b_out = _p(P, prefix, "out_proj.bias")
    B, Lq, H = q_in.shape
    Lk = kv_in.shape[1]
    d = H // heads
    q = affine(q_in, w_in[0:H], b_in[0:H]) * (1.0 / math.sqrt(d))
    k = affine(kv_in, w_in[H:2 * H], b_in[H:2 * H])
    v = affine(kv_in, w_in[2 * H:3 * H], b_in[2 * H:3 * H])
    q = q.reshape(B, Lq, heads, d).permute(0, 2, 1, 3)
    k = k.reshape(B, Lk, heads, d).permute(0, 2, 1, 3)
    v = v.reshape(B, Lk, heads, d).permute(0, 2, 1, 3)
    scores = torch.matmul(q, k.transpose(-1, -2))          # [B,h,Lq,Lk]
    prob = softmax_lastdim(scores)
    if TRAIN_DROPOUT > 0.0:
        prob = torch.nn.functional.dropout(prob, TRAIN_DROPOUT, True)
    ctx = torch.matmul(prob, v)                            # [B,h,Lq,d]
    ctx = ctx.permute(0, 2, 1, 3).reshape(B, Lq, H)
    return affine(ctx, w_out, b_out), prob.mean(dim=1)


def cross_block(query: Tensor, key_value: Tensor, P: Params, prefix: str, heads: int) -> Tensor:
    """CrossModalTransformer.forward, models/fusion_layers.py:202-211 (post-LN)."""
    attn, _ = multi_head_attention(query, key_value, P, prefix + "attention.", heads)
    x = layer_norm(query + attn, _p(P, prefix, "norm1.weight"), _p(P, prefix, "norm1.bias"))
    hid = torch.relu(affine(x, _p(P, prefix, "ffn.0.weight"), _p(P, prefix, "ffn.0.bias")))
    if TRAIN_DROPOUT > 0.0:
        hid = torch.nn.functional.dropout(hid, TRAIN_DROPOUT, True)
    y = affine(hid, _p(P, prefix, "ffn.3.weight"), _p(P, prefix, "ffn.3.bias"))
    return layer_norm(x + y, _p(P, prefix, "norm2.weight"), _p(P, prefix, "norm2.bias"))


# ----------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------
def early_fusion(t: Tensor, a: Tensor, v: Tensor, P: Params, prefix: str = "") -> Tensor:
    """EarlyFusion.forward, models/fusion_layers.py:30-43."""
    x = torch.cat([t, a, v], dim=-1)
    x = torch.relu(affine(x, _p(P, prefix, "fusion_layers.0.weight"), _p(P, prefix, "fusion_layers.0.bias")))
    return torch.relu(affine(x, _p(P, prefix, "fusion_layers.3.weight"), _p(P, prefix, "fusion_layers.3.bias")))


def late_fusion(t: Tensor, a: Tensor, v: Tensor, P: Params, prefix: str = "") -> Dict[str, Tensor]:
    """LateFusion.forward, models/fusion_layers.py:62-90."""
    lt = affine(t, _p(P, prefix, "text_classifier.weight"), _p(P, prefix, "text_classifier.bias"))
    la = affine(a, _p(P, prefix, "audio_classifier.weight"), _p(P, prefix, "audio_classifier.bias"))
    lv = affine(v, _p(P, prefix, "video_classifier.weight"), _p(P, prefix, "video_classifier.bias"))
    w = softmax_lastdim(_p(P, prefix, "fusion_weights"))
    return {"fused_logits": w[0] * lt + w[1] * la + w[2] * lv,
            "text_logits": lt, "audio_logits": la, "video_logits": lv, "fusion_weights": w}


def mult_fusion(t: Tensor, a: Tensor, v: Tensor, P: Params, prefix: str = "", heads: int = 8) -> Dict[str, Tensor]:
    """MultimodalTransformer.forward, models/fusion_layers.py:130-179."""
    if t.dim() == 2:                                        # :140-143
        t, a, v = t.unsqueeze(1), a.unsqueeze(1), v.unsqueeze(1)
    blk = lambda name, q, kv: cross_block(q, kv, P, f"{prefix}{name}.", heads)
    et = t + blk("text_to_audio", t, a) + blk("text_to_video", t, v)      # :146-147,156
    ea = a + blk("audio_to_text", a, t) + blk("audio_to_video", a, v)     # :149-150,157
    ev = v + blk("video_to_text", v, t) + blk("video_to_audio", v, a)     # :152-153,158
    pooled = []
    for name, x in (("text", et), ("audio", ea), ("video", ev)):          # :161-168
        y, _ = multi_head_attention(x, x, P, f"{prefix}{name}_self_attn.", heads)
        pooled.append(y.mean(dim=1))
    fused = torch.relu(affine(torch.cat(pooled, dim=-1),                   # :171-172
                              _p(P, prefix, "final_fusion.0.weight"), _p(P, prefix, "final_fusion.0.bias")))
    return {"fused_features": fused, "text_features": pooled[0],
            "audio_features": pooled[1], "video_features": pooled[2]}


def gat_layer(x: Tensor, P: Params, prefix: str, lin_key: str = "lin.weight") -> Tensor:
    """Dense restatement of torch_geometric GATConv(heads=4, concat=False,
    add_self_loops=True, negative_slope=0.2, bias=True) on B independent,
    fully connected 3-node graphs (PARITY UNPINNED -- PyG is not available).

    x: [B,3,C_in].  With the edge list of models/fusion_layers.py:267-270 plus
    the self-loops GATConv adds, every node attends over all 3 nodes of its own
    sample.  e[i<-j,h] = leaky_relu(a_src[j,h] + a_dst[i,h]); softmax over j;
    out[i,h] = sum_j alpha[i<-j,h] * x'[j,h]; mean over heads; + bias.
    """
    w = _p(P, prefix, lin_key)                              # [4*C, C_in]
    att_src = _p(P, prefix, "att_src").reshape(GAT_HEADS, -1)
    att_dst = _p(P, prefix, "att_dst").reshape(GAT_HEADS, -1)
    bias = _p(P, prefix, "bias")
    B, N, _ = x.shape
    C = att_src.shape[1]
    xp = affine(x, w, None).reshape(B, N, GAT_HEADS, C)
    a_src = (xp * att_src).sum(dim=-1)                      # [B,N,h]
    a_dst = (xp * att_dst).sum(dim=-1)
    e = a_dst.unsqueeze(2) + a_src.unsqueeze(1)             # [B, i, j, h]
    e = torch.where(e > 0, e, GAT_SLOPE * e)
    alpha = softmax_lastdim(e.permute(0, 1, 3, 2))          # [B,i,h,j]
    out = torch.einsum("bihj,bjhc->bihc", alpha, xp)
    return out.mean(dim=2) + bias


def graph_fusion(t: Tensor, a: Tensor, v: Tensor, P: Params, prefix: str = "", num_layers: int = 3) -> Tensor:
    """GraphFusion.forward, models/fusion_layers.py:240-291, with the per-sample
    Data/Batch construction replaced by a dense [B,3,C] tensor."""
    x = torch.stack([t, a, v], dim=1) + _p(P, prefix, "node_type_embedding.weight")   # :255-264
    for i in range(num_layers):                                                        # :281-282
        x = torch.relu(gat_layer(x, P, f"{prefix}gcn_layers.{i}."))
    pooled = x.mean(dim=1)                                                             # :286
    return affine(pooled, _p(P, prefix, "output_projection.weight"), _p(P, prefix, "output_projection.bias"))


def l2_normalize(x: Tensor) -> Tensor:
    n = torch.sqrt((x * x).sum(dim=-1, keepdim=True))
    return x / torch.clamp(n, min=L2_EPS)


def info_nce(z1: Tensor, z2: Tensor, temperature: float) -> Tensor:
    """ContrastiveFusion.contrastive_loss, models/fusion_layers.py:361-375:
    0.5 * [CE(S, arange) + CE(S^T, arange)], S = z1 z2^T / tau, mean over rows."""
    s = torch.matmul(z1, z2.transpose(0, 1)) / temperature
    diag = torch.diagonal(s)
    lse_r = torch.logsumexp(s, dim=1)
    lse_c = torch.logsumexp(s, dim=0)
    return 0.5 * ((lse_r - diag).mean() + (lse_c - diag).mean())


def contrastive_fusion(t: Tensor, a: Tensor, v: Tensor, P: Params, prefix: str = "",
                       temperature: float = 0.07, compute_contrastive_loss: bool = False) -> Dict[str, object]:
    """ContrastiveFusion.forward, models/fusion_layers.py:329-359."""
    def project(x, name):
        h = torch.relu(affine(x, _p(P, prefix, f"{name}.0.weight"), _p(P, prefix, f"{name}.0.bias")))
        return l2_normalize(affine(h, _p(P, prefix, f"{name}.2.weight"), _p(P, prefix, f"{name}.2.bias")))
    zt, za, zv = project(t, "text_projector"), project(a, "audio_projector"), project(v, "video_projector")
    losses = {}
    if compute_contrastive_loss:                            # :344-347
        losses = {"text_audio": info_nce(zt, za, temperature),
                  "text_video": info_nce(zt, zv, temperature),
                  "audio_video": info_nce(za, zv, temperature)}
    fused = torch.relu(affine(torch.cat([t, a, v], dim=-1),
                              _p(P, prefix, "fusion_layer.0.weight"), _p(P, prefix, "fusion_layer.0.bias")))
    return {"fused_features": fused, "text_proj": zt, "audio_proj": za, "video_proj": zv,
            "contrastive_losses": losses}


def adaptive_fusion(t: Tensor, a: Tensor, v: Tensor, P: Params, prefix: str = "", heads: int = 8) -> Dict[str, Tensor]:
    """AdaptiveFusion.forward, models/fusion_layers.py:414-452."""
    tokens = torch.stack([affine(t, _p(P, prefix, "text_transform.weight"), _p(P, prefix, "text_transform.bias")),
                          affine(a, _p(P, prefix, "audio_transform.weight"), _p(P, prefix, "audio_transform.bias")),
                          affine(v, _p(P, prefix, "video_transform.weight"), _p(P, prefix, "video_transform.bias"))], dim=1)
    attended, weights = multi_head_attention(tokens, tokens, P, prefix + "attention.", heads)   # :432-434
    gate_h = torch.relu(affine(torch.cat([t, a, v], dim=-1),                                   # :437-438 (raw inputs)
                               _p(P, prefix, "weight_predictor.0.weight"), _p(P, prefix, "weight_predictor.0.bias")))
    gate = softmax_lastdim(affine(gate_h, _p(P, prefix, "weight_predictor.2.weight"), _p(P, prefix, "weight_predictor.2.bias")))
    mixed = (attended * gate.unsqueeze(-1)).sum(dim=1)                                          # :441-443
    fused = torch.relu(affine(mixed, _p(P, prefix, "fusion_layer.0.weight"), _p(P, prefix, "fusion_layer.0.bias")))
    return {"fused_features": fused, "attention_weights": weights, "adaptive_weights": gate}


def encoder_pool(seq: Tensor, pooling: str, attention_mask: Optional[Tensor] = None) -> Tensor:
    """How the reference encoders reduce `sequence_output` [B,L,D] to one vector per sample before `projection`:
    'cls' = token 0 (text encoder when 'bert' is in the backbone's model_type -- true for deberta-v2, models/encoders.py:86-87);
    'masked_mean' = attention-mask mean with the 1e-9 clamp (text encoder otherwise, :88-93);
    'mean' = plain mean over the sequence (audio :157, video :241)."""
    if pooling == "cls":
        return seq[:, 0]
    if pooling == "mean":
        return seq.mean(dim=1)
    if pooling == "masked_mean":
        m = attention_mask.to(seq.dtype).unsqueeze(-1).expand(seq.size())
        return (seq * m).sum(1) / m.sum(1).clamp(min=1e-9)
    raise ValueError(pooling)


def sequence_projection(seq: Tensor, w: Tensor, b: Tensor, pooling: str, attention_mask: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """SURVEY 8f rank 2 (extension feeding MulT real sequences): the encoder's own `projection` Linear (models/encoders.py:35,134,207)
    applied per token, and the encoder's pooled `features` = projection(pool(sequence_output)) (:96, :160, :244; dropout p = 0).
    `features` is computed the reference's way (pool first, then project); pooling the projected tokens gives the same vector
    because the pooling weights sum to one (checked in tests/test_oracle.py)."""
    return {"sequence_features": affine(seq, w, b), "features": affine(encoder_pool(seq, pooling, attention_mask), w, b)}


def pool_sequence(x: Tensor) -> Tensor:
    """The one explicit step the hierarchical head needs for [B,L,H] inputs
    (SURVEY F3): mean over L, the rule of models/fusion_layers.py:166-168."""
    return x if x.dim() == 2 else x.mean(dim=1)


def hierarchical_fusion(t: Tensor, a: Tensor, v: Tensor, P: Params, prefix: str = "", heads: int = 8,
                        graph_layers: int = 3, temperature: float = 0.07,
                        compute_contrastive_loss: bool = False, pooled=None) -> Dict[str, object]:
    """HierarchicalFusion.forward, models/fusion_layers.py:478-520.  For 2-D
    inputs this is the literal reference; for 3-D inputs MulT sees the sequences
    and the four 2-D-only heads see `pool_sequence` of them, or the
    encoders' own pooled features when `pooled=(t2, a2, v2)` is given."""
    t2, a2, v2 = pooled if pooled is not None else (pool_sequence(t), pool_sequence(a), pool_sequence(v))
    early = early_fusion(t2, a2, v2, P, prefix + "early_fusion.")
    mult = mult_fusion(t, a, v, P, prefix + "mult_fusion.", heads)["fused_features"]
    graph = graph_fusion(t2, a2, v2, P, prefix + "graph_fusion.", graph_layers)
    con = contrastive_fusion(t2, a2, v2, P, prefix + "contrastive_fusion.", temperature, compute_contrastive_loss)
    ada = adaptive_fusion(t2, a2, v2, P, prefix + "adaptive_fusion.", heads)
    cat = torch.cat([early, mult, graph, con["fused_features"], ada["fused_features"]], dim=-1)   # :503-506
    hid = torch.relu(affine(cat, _p(P, prefix, "meta_fusion.0.weight"), _p(P, prefix, "meta_fusion.0.bias")))
    fused = affine(hid, _p(P, prefix, "meta_fusion.3.weight"), _p(P, prefix, "meta_fusion.3.bias"))
    return {"fused_features": fused, "early_features": early, "mult_features": mult,
            "graph_features": graph, "contrastive_features": con["fused_features"],
            "adaptive_features": ada["fused_features"], "contrastive_losses": con["contrastive_losses"],
            "attention_weights": ada["attention_weights"], "adaptive_weights": ada["adaptive_weights"]}


# ----------------------------------------------------------------------------
# modality dropout (models/encoders.py:280-321)
# ----------------------------------------------------------------------------
def modality_keep_mask(batch: int, rate: float, generator: torch.Generator) -> Tensor:
    """[B,3] 0/1 keep-mask: keep iff U(0,1) > rate per (sample, modality); rows with
    all three dropped get exactly one uniformly chosen modality re-enabled
    (models/encoders.py:303-314)."""
    keep = torch.rand(batch, 3, generator=generator) > rate
    dead = ~keep.any(dim=1)
    n_dead = int(dead.sum())
    if n_dead:
        pick = torch.randint(0, 3, (n_dead,), generator=generator)
        keep[dead] = torch.nn.functional.one_hot(pick, 3).bool()
    return keep.to(torch.float32)


def apply_modality_mask(t: Tensor, a: Tensor, v: Tensor, mask: Optional[Tensor]):
    """feat * mask, no 1/(1-p) rescale (models/encoders.py:317-319); broadcasts
    over L for [B,L,H] inputs (extension)."""
    if mask is None:
        return t, a, v
    def mul(x, col):
        m = mask[:, col].to(x.dtype)
        return x * (m[:, None] if x.dim() == 2 else m[:, None, None])
    return mul(t, 0), mul(a, 1), mul(v, 2)


# ----------------------------------------------------------------------------
# parameter construction (reference default init, SURVEY 8b "Default init")
# ----------------------------------------------------------------------------
def _linear_init(g, out_f, in_f):
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=g) * 2 - 1) * bound
    return w, b


def _mha_init(g, P, prefix, H):
    bound = math.sqrt(6.0 / (H + 3 * H))                    # xavier_uniform on [3H,H]
    P[prefix + "in_proj_weight"] = (torch.rand(3 * H, H, generator=g) * 2 - 1) * bound
    P[prefix + "in_proj_bias"] = torch.zeros(3 * H)
    w, _ = _linear_init(g, H, H)
    P[prefix + "out_proj.weight"] = w
    P[prefix + "out_proj.bias"] = torch.zeros(H)


def _put_linear(g, P, name, out_f, in_f):
    P[name + ".weight"], P[name + ".bias"] = _linear_init(g, out_f, in_f)


def init_params(kind: str, H: int = 512, heads: int = 8, num_emotions: int = 7, graph_hidden: int = 512,
                graph_layers: int = 3, seed: int = 0, randomize_affine: bool = True) -> Params:
    """Random parameters with the reference's names, shapes and init
    distributions for head `kind` in {early, late, mult, graph, contrastive,
    adaptive, hierarchical}.  With `randomize_affine` the LayerNorm gains/biases
    and MHA biases are perturbed away from their 1/0 defaults so that parity
    tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    P: Params = {}

    def early(pre):
        _put_linear(g, P, pre + "fusion_layers.0", 2 * H, 3 * H)
        _put_linear(g, P, pre + "fusion_layers.3", H, 2 * H)

    def late(pre):
        for m in ("text", "audio", "video"):
            _put_linear(g, P, pre + f"{m}_classifier", num_emotions, H)
        P[pre + "fusion_weights"] = torch.ones(3) / 3

    def mult(pre):
        for blk in ("text_to_audio", "text_to_video", "audio_to_text", "audio_to_video", "video_to_text", "video_to_audio"):
            b = f"{pre}{blk}."
            _mha_init(g, P, b + "attention.", H)
            for n in ("norm1", "norm2"):
                P[b + n + ".weight"] = torch.ones(H)
                P[b + n + ".bias"] = torch.zeros(H)
            _put_linear(g, P, b + "ffn.0", 4 * H, H)
            _put_linear(g, P, b + "ffn.3", H, 4 * H)
        for m in ("text", "audio", "video"):
            _mha_init(g, P, f"{pre}{m}_self_attn.", H)
        _put_linear(g, P, pre + "final_fusion.0", H, 3 * H)

    def graph(pre):
        for i in range(graph_layers):
            b = f"{pre}gcn_layers.{i}."
            C = graph_hidden
            bw = math.sqrt(6.0 / (H + GAT_HEADS * C))        # glorot
            P[b + "lin.weight"] = (torch.rand(GAT_HEADS * C, H, generator=g) * 2 - 1) * bw
            ba = math.sqrt(6.0 / (GAT_HEADS + C))
            P[b + "att_src"] = (torch.rand(1, GAT_HEADS, C, generator=g) * 2 - 1) * ba
            P[b + "att_dst"] = (torch.rand(1, GAT_HEADS, C, generator=g) * 2 - 1) * ba
            P[b + "bias"] = torch.zeros(C)
        P[pre + "node_type_embedding.weight"] = torch.randn(3, H, generator=g)
        _put_linear(g, P, pre + "output_projection", H, graph_hidden)

    def contrastive(pre):
        for m in ("text", "audio", "video"):
            _put_linear(g, P, pre + f"{m}_projector.0", H, H)
            _put_linear(g, P, pre + f"{m}_projector.2", H // 2, H)
        _put_linear(g, P, pre + "fusion_layer.0", H, 3 * H)

    def adaptive(pre):
        _mha_init(g, P, pre + "attention.", H)
        for m in ("text", "audio", "video"):
            _put_linear(g, P, pre + f"{m}_transform", H, H)
        _put_linear(g, P, pre + "weight_predictor.0", H, 3 * H)
        _put_linear(g, P, pre + "weight_predictor.2", 3, H)
        _put_linear(g, P, pre + "fusion_layer.0", H, H)

    builders = {"early": early, "late": late, "mult": mult, "graph": graph,
                "contrastive": contrastive, "adaptive": adaptive}
    if kind == "hierarchical":
        early("early_fusion."); mult("mult_fusion."); graph("graph_fusion.")
        contrastive("contrastive_fusion."); adaptive("adaptive_fusion.")
        _put_linear(g, P, "meta_fusion.0", 2 * H, 5 * H)
        _put_linear(g, P, "meta_fusion.3", H, 2 * H)
    else:
        builders[kind]("")
    if randomize_affine:
        for k in list(P):
            if k.endswith(("norm1.weight", "norm2.weight")):
                P[k] = P[k] + 0.1 * torch.randn(P[k].shape, generator=g)
            elif k.endswith(("norm1.bias", "norm2.bias", "in_proj_bias", "out_proj.bias")) or \
                    (k.endswith(".bias") and "gcn_layers" in k):
                P[k] = 0.05 * torch.randn(P[k].shape, generator=g)
    return P


def synthetic_features(batch: int, lens=(None, None, None), H: int = 512, seed: int = 1234,
                       dtype=torch.float32):
    """Seeded N(0,1) stand-ins for the encoder features: `lens[i]` None -> [B,H],
    int L -> [B,L,H]."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for L in lens:
        shape = (batch, H) if L is None else (batch, L, H)
        out.append(torch.randn(shape, generator=g).to(dtype))
    return tuple(out)


# ----------------------------------------------------------------------------
# heads directly downstream of the fusion output
# ----------------------------------------------------------------------------
def emotion_classifier(f: Tensor, P: Params, prefix: str = "") -> Tensor:
    """EmotionClassifier.forward, models/multimodal_model.py:210-219: only the main logits are returned."""
    h = torch.relu(affine(f, _p(P, prefix, "classifier.0.weight"), _p(P, prefix, "classifier.0.bias")))
    return affine(h, _p(P, prefix, "classifier.3.weight"), _p(P, prefix, "classifier.3.bias"))


def auxiliary_heads(f: Tensor, P: Params, prefix: str = "") -> Dict[str, Tensor]:
    """valence / arousal regressors and the uncertainty head of MultimodalEmotionModel (multimodal_model.py:55-60,146-164)."""
    lin = lambda n: affine(f, _p(P, prefix, n + ".weight"), _p(P, prefix, n + ".bias"))
    return {"valence": lin("valence_regressor"), "arousal": lin("arousal_regressor"),
            "uncertainty": softmax_lastdim(lin("uncertainty_head"))}


def cross_entropy_label_smoothing(logits: Tensor, target: Tensor, eps: float = 0.1) -> Tensor:
    """nn.CrossEntropyLoss(label_smoothing=eps), mean reduction (training/advanced_trainer.py:53,139):
    (1-eps) * NLL(target) + eps * mean_c(-log p_c), averaged over the batch."""
    logp = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    nll = -logp.gather(1, target[:, None]).squeeze(1)
    return ((1.0 - eps) * nll + eps * (-logp.mean(dim=-1))).mean()


def init_head_params(H: int = 512, num_emotions: int = 7, seed: int = 0) -> Params:
    """nn.Linear default init for EmotionClassifier + the three auxiliary heads (reference state_dict names)."""
    g = torch.Generator().manual_seed(seed)
    P: Params = {}
    for name, o, i in (("classifier.0", H // 2, H), ("classifier.3", num_emotions, H // 2), ("sentiment_classifier", 3, H),
                       ("positive_classifier", 2, H), ("negative_classifier", 4, H), ("valence_regressor", 1, H),
                       ("arousal_regressor", 1, H), ("uncertainty_head", num_emotions, H)):
        _put_linear(g, P, name, o, i)
    return P


HEADS = {
    "early": early_fusion, "late": late_fusion, "mult": mult_fusion, "graph": graph_fusion,
    "contrastive": contrastive_fusion, "adaptive": adaptive_fusion, "hierarchical": hierarchical_fusion,
}


# ----------------------------------------------------------------------------
# scalar objective used by parity tests, golden vectors and the bench:
# mean(fused^2) + 0.1 * sum(contrastive losses), plus a small weight on every other
# tensor output so that each returned tensor carries gradient in the tests.
# ----------------------------------------------------------------------------
MAIN_KEYS = ("fused_features", "fused_logits")
AUX_WEIGHT = 0.05
CONTRASTIVE_WEIGHT = 0.1      # training/advanced_trainer.py:163


def objective(out) -> Tensor:
    if isinstance(out, torch.Tensor):
        return (out.double() ** 2).mean() if out.dtype != torch.float64 else (out ** 2).mean()
    total = None
    for k in sorted(out):
        v = out[k]
        if k == "contrastive_losses":
            for name in sorted(v):
                term = CONTRASTIVE_WEIGHT * v[name].double()
                total = term if total is None else total + term
        elif isinstance(v, torch.Tensor) and v.is_floating_point():
            w = 1.0 if k in MAIN_KEYS else AUX_WEIGHT
            term = w * (v.double() ** 2).mean()
            total = term if total is None else total + term
    return total
