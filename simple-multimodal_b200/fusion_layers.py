"""Drop-in replacements for the reference's fusion heads (models/fusion_layers.py) on B200.

Same class names, `Cls(config)` constructors, `forward(text, audio, video[, compute_contrastive_loss])`
signatures, return types / dict keys and `state_dict()` layout as the reference, so
`models/multimodal_model.py:29-46,110-144` and reference checkpoints work unchanged
(see INTEGRATION.md).  The sub-modules registered here are *parameter containers only*: no
`nn.Linear.forward`, `nn.MultiheadAttention.forward` or `nn.LayerNorm.forward` is ever called -- all
arithmetic runs in libb200fusion.so through `ops.py` / `mult_engine.py`, and there is no fallback.

Extensions over the reference: an optional trailing `mask=None` keyword ([B,3] keep-mask,
the ModalityDropout multiply of models/encoders.py:317-319 fused into the first load), and
HierarchicalFusion accepts [B,L,H] sequences (MulT sees them, the other heads see their mean over L, or the
`pooled_features=(t, a, v)` given by the caller -- see sequence_features.SequenceProjector).

Precision: bfloat16 inputs (or `compute_dtype=torch.bfloat16`, or CUDA autocast) run the tcgen05
kernels; float32 inputs run the fp32 parity kernels.  Parameters are always fp32 masters.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import kernels as K
from . import mult_engine
from . import ops
from ._lib import B200FusionError

Tensor = torch.Tensor
GAT_HEADS = 4           # fusion_layers.py:227
GAT_SLOPE = 0.2         # GATConv default negative_slope


def _mlp_container(sizes, dropout, final_relu):
    """nn.Sequential with the reference's child indices (Linear, ReLU, Dropout, Linear, ...) so that
    parameter names line up; used only to hold parameters."""
    mods = []
    for i in range(len(sizes) - 1):
        mods.append(nn.Linear(sizes[i], sizes[i + 1]))
        last = i == len(sizes) - 2
        if not last or final_relu:
            mods.append(nn.ReLU())
            if dropout is not None:
                mods.append(nn.Dropout(dropout))
    return nn.Sequential(*mods)


class _FusionBase(nn.Module):
    compute_dtype: Optional[torch.dtype] = None

    def _prepare(self, feats, mask):
        """Pick the compute dtype, cast inputs with the library's cast kernels, validate the mask."""
        t = feats[0]
        if not t.is_cuda:
            raise B200FusionError("b200 fusion heads run on CUDA tensors only (there is no CPU fallback)")
        dt = self.compute_dtype
        if dt is None:
            if torch.is_autocast_enabled("cuda"):
                dt = torch.bfloat16
            else:
                dt = t.dtype if t.dtype in (torch.float32, torch.bfloat16) else torch.bfloat16
        out = []
        for x in feats:
            if x.dtype != dt:
                x = _CastFn.apply(x, dt)
            out.append(x)
        if mask is not None:
            if mask.shape != (t.size(0), 3):
                raise B200FusionError(f"mask must be [B,3], got {tuple(mask.shape)}")
            mask = mask.to(device=t.device, dtype=torch.float32).contiguous()
        return out, mask, dt

    @property
    def _p(self) -> float:
        return float(getattr(self.config, "fusion_dropout", 0.0))


class _CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dt):
        ctx.src = x.dtype
        if x.dtype == torch.float32 and dt == torch.bfloat16:
            return K.cast_to_bf16(x)
        if x.dtype == torch.bfloat16 and dt == torch.float32:
            return K.cast_to_f32(x)
        if x.dtype == torch.float16:
            return K.cast_to_bf16(x.float()) if dt == torch.bfloat16 else x.float()
        raise B200FusionError(f"cannot cast {x.dtype} -> {dt}")

    @staticmethod
    def backward(ctx, g):
        if ctx.src == torch.float32:
            return (K.cast_to_f32(g) if g.dtype == torch.bfloat16 else g), None
        if ctx.src == torch.bfloat16:
            return (K.cast_to_bf16(g) if g.dtype == torch.float32 else g), None
        return g.to(ctx.src), None


def _masked_cat(t, a, v, mask):
    return ops.Concat3Fn.apply(t, a, v, mask)


def _masked_split(t, a, v, mask):
    """Per-modality features with the keep-mask applied (one fused concat+mask pass, then views)."""
    if mask is None:
        return t, a, v
    return ops.Split3Fn.apply(_masked_cat(t, a, v, mask))


class EarlyFusion(_FusionBase):
    """reference models/fusion_layers.py:9-43."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H = config.fusion_hidden_size
        self.fusion_layers = _mlp_container([3 * H, 2 * H, H], config.fusion_dropout, final_relu=True)

    def forward(self, text_features, audio_features, video_features, mask=None) -> Tensor:
        (t, a, v), mask, _ = self._prepare((text_features, audio_features, video_features), mask)
        return self._run(_masked_cat(t, a, v, mask))

    def _run(self, cat):
        l0, l3 = self.fusion_layers[0], self.fusion_layers[3]
        h = ops.dropout(ops.linear(cat, l0.weight, l0.bias, relu=True), self._p, self.training)
        return ops.dropout(ops.linear(h, l3.weight, l3.bias, relu=True), self._p, self.training)


class LateFusion(_FusionBase):
    """reference models/fusion_layers.py:46-90."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H, E = config.fusion_hidden_size, config.num_emotions
        self.text_classifier = nn.Linear(H, E)
        self.audio_classifier = nn.Linear(H, E)
        self.video_classifier = nn.Linear(H, E)
        self.fusion_weights = nn.Parameter(torch.ones(3) / 3)

    def forward(self, text_features, audio_features, video_features, mask=None) -> Dict[str, Tensor]:
        (t, a, v), mask, _ = self._prepare((text_features, audio_features, video_features), mask)
        t, a, v = _masked_split(t, a, v, mask)
        lt = ops.linear(t, self.text_classifier.weight, self.text_classifier.bias)
        la = ops.linear(a, self.audio_classifier.weight, self.audio_classifier.bias)
        lv = ops.linear(v, self.video_classifier.weight, self.video_classifier.bias)
        fused, w = ops.LateCombineFn.apply(lt, la, lv, self.fusion_weights)
        return {"fused_logits": fused, "text_logits": lt, "audio_logits": la, "video_logits": lv, "fusion_weights": w}


class CrossModalTransformer(_FusionBase):
    """reference models/fusion_layers.py:182-211.  Parameter container for MultimodalTransformer; its own
    forward (one block) is also available and runs the same kernels op by op."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H = config.fusion_hidden_size
        self.attention = nn.MultiheadAttention(H, config.fusion_num_heads, dropout=config.fusion_dropout, batch_first=True)
        self.norm1 = nn.LayerNorm(H)
        self.norm2 = nn.LayerNorm(H)
        self.ffn = _mlp_container([H, 4 * H, H], config.fusion_dropout, final_relu=False)

    def forward(self, query: Tensor, key_value: Tensor) -> Tensor:
        (q, kv), _, _ = self._prepare((query, key_value), None)
        H, heads = self.config.fusion_hidden_size, self.config.fusion_num_heads
        att = self.attention
        pq = ops.linear(q, att.in_proj_weight[:H], att.in_proj_bias[:H])
        pkv = ops.linear(kv, att.in_proj_weight[H:], att.in_proj_bias[H:])
        drop = (self._p, *ops.next_drop_seed()) if (self.training and self._p > 0.0) else None
        ctx = ops.AttentionFn.apply(pq, pkv, 0, 0, H, H, heads, ops.mha_scale(H, heads), drop)
        x = ops.LayerNormFn.apply(ops.AddFn.apply(q, ops.linear(ctx, att.out_proj.weight, att.out_proj.bias), None),
                                  self.norm1.weight, self.norm1.bias)
        h = ops.dropout(ops.linear(x, self.ffn[0].weight, self.ffn[0].bias, relu=True), self._p, self.training)
        y = ops.linear(h, self.ffn[3].weight, self.ffn[3].bias)
        return ops.LayerNormFn.apply(ops.AddFn.apply(x, y, None), self.norm2.weight, self.norm2.bias)


class MultimodalTransformer(_FusionBase):
    """reference models/fusion_layers.py:93-179, executed by mult_engine.MulTFn (chunked, fused schedule)."""

    chunk_size = 1024           # samples per MulT chunk (~30 GB of bf16 activations at L=512/512/30, H=512; measured on the B=4096 step:
                                # 256 -> 324.3 ms, 512 -> 318.1 ms, 1024 -> 314.7 ms: fewer, larger launches -- shorter GEMM tails, more items
                                # per persistent CTA; the graph engine halves it by itself when memory is short)
    stash_fraction = 0.72       # share of the currently free device memory that forward may keep resident for backward
    graph_chunks = True         # bf16 training steps replay captured per-chunk CUDA graphs (mult_engine.ChunkGraphEngine)
    graph_min_tokens = 16384    # ... when a chunk is big enough for launch overhead to matter (tokens per chunk, all modalities)
    _engine = None              # the ChunkGraphEngine of the last shape seen (one is kept: its pool holds the activation stash)
    _engine_key = None
    _engine_builds = 0
    _weights = None             # dtype -> mult_engine._Weights (operand copies at static addresses)

    def __init__(self, config):
        super().__init__()
        self.config = config
        H = config.fusion_hidden_size
        for name, _, _ in mult_engine.BLOCKS:
            setattr(self, name, CrossModalTransformer(config))
        for m in mult_engine.MODS:
            setattr(self, f"{m}_self_attn", nn.MultiheadAttention(H, config.fusion_num_heads, dropout=config.fusion_dropout,
                                                                  batch_first=True))
        self.final_fusion = _mlp_container([3 * H, H], config.fusion_dropout, final_relu=True)
        self._names = mult_engine.param_names()

    def __getstate__(self):                                  # copies / pickles of the module never carry graphs or operand caches
        state = dict(self.__dict__)
        state["_engine"], state["_engine_key"], state["_weights"] = None, None, {}
        return state

    def _operands(self, dt):
        """the persistent operand copies of the parameters for compute dtype `dt` (rebuilt if the parameters were re-created,
        e.g. by .to(device))"""
        params = dict(self.named_parameters())
        P = {n: params[n] for n in self._names}
        cache = self.__dict__.setdefault("_weights", {})
        W = cache.get(dt)
        if W is None or any(W.P[n] is not P[n] for n in self._names) or W.w_stack[0].device != P[self._names[0]].device:
            self.release_graphs()
            W = cache[dt] = mult_engine._Weights(P, self.config.fusion_hidden_size, dt)
        return W, [P[n] for n in self._names]

    def release_graphs(self) -> None:
        """free the captured chunk graphs and the activation stash their pool holds"""
        if self.__dict__.get("_engine") is not None:
            self._engine.release()
        self._engine, self._engine_key = None, None

    def _graph_engine(self, W, xs, need_dx, p_drop, want_mean=False):
        """ChunkGraphEngine for this step, or None when the step is issued eagerly: fp32 parity mode, inference, inside an outer
        CUDA-graph capture, tiny chunks, shapes that keep changing, or too little free memory to keep every chunk's stash."""
        t = xs[0]
        if not (self.graph_chunks and t.dtype == torch.bfloat16 and torch.is_grad_enabled() and not torch.cuda.is_current_stream_capturing()):
            return None
        B, Ls, chunk = t.size(0), [x.size(1) for x in xs], int(self.chunk_size)
        if min(B, chunk) * sum(Ls) < self.graph_min_tokens:
            return None
        key = (B, tuple(Ls), chunk, t.device, float(p_drop), bool(need_dx), bool(want_mean), id(W))
        if self._engine is not None and self._engine_key == key:
            return self._engine
        if self._engine_builds >= 4:                     # shapes keep changing: capturing each of them costs more than it saves
            return None
        self.release_graphs()
        H = self.config.fusion_hidden_size
        torch.cuda.empty_cache()
        free_bytes, _ = torch.cuda.mem_get_info(t.device)
        stash = mult_engine.stash_bytes_per_sample(Ls, H, t.element_size())
        static = (2 if need_dx else 1) * B * sum(Ls) * H * t.element_size()        # masked inputs (+ input gradients)
        if B * stash > self.stash_fraction * free_bytes:
            return None                                      # not every chunk can stay resident: eager issue with recomputed chunks
        # the backward temporaries of the chunk in flight (~0.55 x its stash, measured) must fit next to the resident stash: halve the
        # engine's chunk until they do (larger chunks are faster -- B = 4096: 256 -> 324 ms, 512 -> 318 ms, 1024 -> 315 ms per step)
        eng_chunk = min(B, chunk)
        while B * stash + static + 4 * eng_chunk * stash // 5 > 0.92 * free_bytes and eng_chunk > 128:
            eng_chunk //= 2
        if B * stash + static + 4 * eng_chunk * stash // 5 > 0.92 * free_bytes:
            return None
        self._engine = mult_engine.ChunkGraphEngine(W, self._names, H, self.config.fusion_num_heads, eng_chunk, B, Ls, t.dtype, t.device,
                                                    p_drop, need_dx, want_mean)
        self._engine_key = key
        self._engine_builds += 1
        return self._engine

    def _pooled(self, t, a, v, mask, want_mean=False):
        """pooled attended features [B,3H]; with `want_mean` also the masked mean over L of the inputs ([B,3H], see MulTFn)"""
        if t.dim() == 2:                                     # fusion_layers.py:140-143
            t, a, v = t.unsqueeze(1), a.unsqueeze(1), v.unsqueeze(1)
        training_drop = self.training and self._p > 0.0
        H, heads = self.config.fusion_hidden_size, self.config.fusion_num_heads
        W, plist = self._operands(t.dtype)
        need_dx = torch.is_grad_enabled() and any(x.requires_grad for x in (t, a, v))
        engine = self._graph_engine(W, (t, a, v), need_dx, self._p if training_drop else 0.0, want_mean)
        drop = None
        if engine is None and training_drop:
            drop = (self._p, *ops.next_drop_seed())          # one seed pair per call
        # memory the stash may use: what the driver reports free plus what torch's caching allocator holds but has not handed out
        if engine is not None:
            budget = 0
        elif torch.cuda.is_current_stream_capturing():
            budget = 1 << 62                                 # CUDA-graph capture: the graph's private pool keeps every chunk anyway
        else:
            free_bytes, _ = torch.cuda.mem_get_info(t.device)
            free_bytes += torch.cuda.memory_reserved(t.device) - torch.cuda.memory_allocated(t.device)
            budget = int(self.stash_fraction * free_bytes) if torch.is_grad_enabled() else 0
        pooled, xmean = mult_engine.MulTFn.apply(t, a, v, mask, bool(want_mean), H, heads, int(self.chunk_size), budget, drop, self._names, W,
                                                 engine, *plist)
        return (pooled, xmean) if want_mean else pooled

    def forward(self, text_features, audio_features, video_features, mask=None) -> Dict[str, Tensor]:
        (t, a, v), mask, _ = self._prepare((text_features, audio_features, video_features), mask)
        pooled = self._pooled(t, a, v, mask)
        H = self.config.fusion_hidden_size
        ff = self.final_fusion[0]
        fused = ops.dropout(ops.linear(pooled, ff.weight, ff.bias, relu=True), self._p, self.training)
        tp, ap, vp = ops.Split3Fn.apply(pooled)
        return {"fused_features": fused, "text_features": tp, "audio_features": ap, "video_features": vp}


class _GATParams(nn.Module):
    """Parameter container with torch_geometric.nn.GATConv's names and shapes (PyG >= 2.5 naming, `lin.weight`;
    the 2.3/2.4 names `lin_src.weight`/`lin_dst.weight` are accepted on load).  Initialised like PyG (glorot)."""

    def __init__(self, in_channels, out_channels, heads):
        super().__init__()
        self.heads, self.out_channels = heads, out_channels
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        nn.init.xavier_uniform_(self.lin.weight)
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for old in ("lin_src.weight", "lin_dst.weight"):
            if prefix + old in state_dict:
                state_dict.setdefault(prefix + "lin.weight", state_dict[prefix + old])
                del state_dict[prefix + old]
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class GraphFusion(_FusionBase):
    """reference models/fusion_layers.py:214-291 on dense [B,3,C] node tensors: the per-sample Python loop,
    Data/Batch construction and scatter-softmax disappear (every sample is the same complete 3-graph)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H, G = config.fusion_hidden_size, config.graph_hidden_size
        self.gcn_layers = nn.ModuleList([_GATParams(H, G, GAT_HEADS) for _ in range(config.graph_num_layers)])
        self.node_type_embedding = nn.Embedding(3, H)
        self.output_projection = nn.Linear(G, H)
        if config.graph_num_layers > 1 and G != H:
            # the reference builds every layer with in=H (fusion_layers.py:223-232): layer 2 cannot consume G != H channels
            raise B200FusionError(f"GraphFusion: graph_hidden_size ({G}) must equal fusion_hidden_size ({H}) when "
                                  f"graph_num_layers > 1 -- the reference constructor has the same constraint (SURVEY F4)")

    def forward(self, text_features, audio_features, video_features, mask=None) -> Tensor:
        (t, a, v), mask, dt = self._prepare((text_features, audio_features, video_features), mask)
        return self._run(_masked_cat(t, a, v, mask), dt)

    def _run(self, cat, dt):
        B, H = cat.size(0), self.config.fusion_hidden_size
        emb = _EmbedAddFn.apply(cat, self.node_type_embedding.weight)            # nodes + type embedding, [B,3H]
        x = emb.view(B, 3, H)
        for layer in self.gcn_layers:
            xp = ops.linear(x, layer.lin.weight, None)                            # [B,3,heads*C]
            gp = float(getattr(self.config, "graph_dropout", 0.0))            # GATConv(dropout=config.graph_dropout), fusion_layers.py:228
            drop = (gp, *ops.next_drop_seed()) if (self.training and gp > 0.0) else None
            x = ops.GatFn.apply(xp, layer.att_src, layer.att_dst, layer.bias, layer.heads, GAT_SLOPE, drop)
        pooled = ops.MeanPoolFn.apply(x)                                          # global_mean_pool over the 3 nodes
        return ops.linear(pooled, self.output_projection.weight, self.output_projection.bias)


class _EmbedAddFn(torch.autograd.Function):
    """cat[B,3H] + flatten(embedding[3,H]) broadcast over the batch, via the GEMM-free add kernel."""

    @staticmethod
    def forward(ctx, cat, emb):
        B = cat.size(0)
        e = emb.detach().reshape(1, -1)
        e = (K.cast_to_bf16(e) if cat.dtype == torch.bfloat16 else e.contiguous()).expand(B, -1).contiguous()
        return K.add(cat.contiguous(), e)

    @staticmethod
    def backward(ctx, g):
        demb = torch.zeros(g.size(1), device=g.device, dtype=torch.float32)
        K.colsum_accum(g.contiguous(), demb)
        return g, demb.view(3, -1)


class ContrastiveFusion(_FusionBase):
    """reference models/fusion_layers.py:294-375.  The InfoNCE negatives span the GLOBAL batch when a
    torch.distributed process group is initialised (ops.InfoNCE3Fn)."""

    process_group = None

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.temperature = config.contrastive_temperature
        H = config.fusion_hidden_size
        self.text_projector = _mlp_container([H, H, H // 2], None, final_relu=False)
        self.audio_projector = _mlp_container([H, H, H // 2], None, final_relu=False)
        self.video_projector = _mlp_container([H, H, H // 2], None, final_relu=False)
        self.fusion_layer = _mlp_container([3 * H, H], config.fusion_dropout, final_relu=True)

    def contrastive_loss(self, z1: Tensor, z2: Tensor) -> Tensor:
        """Single-pair loss with the reference's signature (fusion_layers.py:361-375)."""
        lt, _, _ = ops.InfoNCE3Fn.apply(z1, z2, z2, self.temperature, self.process_group)
        return lt

    def forward(self, text_features, audio_features, video_features, compute_contrastive_loss: bool = False, mask=None):
        (t, a, v), mask, _ = self._prepare((text_features, audio_features, video_features), mask)
        cat = _masked_cat(t, a, v, mask)
        feats = ops.Split3Fn.apply(cat) if mask is not None else (t, a, v)
        return self._run(cat, feats, compute_contrastive_loss)

    def _run(self, cat, feats, compute_contrastive_loss):
        z = []
        for x, proj in zip(feats, (self.text_projector, self.audio_projector, self.video_projector)):
            h = ops.linear(x, proj[0].weight, proj[0].bias, relu=True)
            z.append(ops.L2NormFn.apply(ops.linear(h, proj[2].weight, proj[2].bias)))
        losses = {}
        if compute_contrastive_loss:
            l = ops.InfoNCE3Fn.apply(z[0], z[1], z[2], self.temperature, self.process_group)
            losses = {name: l[i] for i, (_, _, name) in enumerate(ops.PAIRS)}
        fl = self.fusion_layer[0]
        fused = ops.dropout(ops.linear(cat, fl.weight, fl.bias, relu=True), self._p, self.training)
        return {"fused_features": fused, "text_proj": z[0], "audio_proj": z[1], "video_proj": z[2], "contrastive_losses": losses}


class AdaptiveFusion(_FusionBase):
    """reference models/fusion_layers.py:378-452."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H = config.fusion_hidden_size
        self.attention = nn.MultiheadAttention(H, config.fusion_num_heads, dropout=config.fusion_dropout, batch_first=True)
        self.text_transform = nn.Linear(H, H)
        self.audio_transform = nn.Linear(H, H)
        self.video_transform = nn.Linear(H, H)
        self.weight_predictor = nn.Sequential(nn.Linear(3 * H, H), nn.ReLU(), nn.Linear(H, 3), nn.Softmax(dim=-1))
        self.fusion_layer = _mlp_container([H, H], config.fusion_dropout, final_relu=True)

    def forward(self, text_features, audio_features, video_features, mask=None) -> Dict[str, Tensor]:
        (t, a, v), mask, _ = self._prepare((text_features, audio_features, video_features), mask)
        cat = _masked_cat(t, a, v, mask)
        feats = ops.Split3Fn.apply(cat) if mask is not None else (t, a, v)
        return self._run(cat, feats)

    def _run(self, cat, feats):
        H, heads = self.config.fusion_hidden_size, self.config.fusion_num_heads
        drop = (self._p, *ops.next_drop_seed()) if (self.training and self._p > 0.0) else None
        tr = [ops.linear(x, m.weight, m.bias) for x, m in zip(feats, (self.text_transform, self.audio_transform, self.video_transform))]
        tokens = ops.Concat3Fn.apply(tr[0], tr[1], tr[2], None).view(-1, 3, H)           # stack(dim=1)
        qkv = ops.linear(tokens, self.attention.in_proj_weight, self.attention.in_proj_bias)  # [B,3,3H]
        ctx, avgw = ops.Tok3AttnFn.apply(qkv, heads, ops.mha_scale(H, heads), drop)
        attended = ops.linear(ctx, self.attention.out_proj.weight, self.attention.out_proj.bias)
        wp0, wp2 = self.weight_predictor[0], self.weight_predictor[2]
        logits = ops.linear(ops.linear(cat, wp0.weight, wp0.bias, relu=True), wp2.weight, wp2.bias)
        mixed, gate = ops.GateMixFn.apply(attended, logits)
        fl = self.fusion_layer[0]
        fused = ops.dropout(ops.linear(mixed, fl.weight, fl.bias, relu=True), self._p, self.training)
        return {"fused_features": fused, "attention_weights": avgw, "adaptive_weights": gate}


class HierarchicalFusion(_FusionBase):
    """reference models/fusion_layers.py:455-520.  [B,L,H] inputs: MulT on the sequences, the four 2-D heads
    on their mean over L."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        H = config.fusion_hidden_size
        self.early_fusion = EarlyFusion(config)
        self.mult_fusion = MultimodalTransformer(config)
        self.graph_fusion = GraphFusion(config)
        self.contrastive_fusion = ContrastiveFusion(config)
        self.adaptive_fusion = AdaptiveFusion(config)
        self.meta_fusion = _mlp_container([5 * H, 2 * H, H], config.fusion_dropout, final_relu=False)
        # the reference's Sequential is Linear, ReLU, Dropout, Linear: drop the trailing activation slots
        self.meta_fusion = nn.Sequential(*list(self.meta_fusion)[:4])

    def forward(self, text_features, audio_features, video_features, compute_contrastive_loss: bool = False, mask=None, pooled_features=None):
        (t, a, v), mask, dt = self._prepare((text_features, audio_features, video_features), mask)
        if pooled_features is not None:        # extension: the encoders' own pooled features for the 2-D heads
            if len(pooled_features) != 3 or any(x.dim() != 2 or x.size(0) != t.size(0) for x in pooled_features):
                raise B200FusionError("pooled_features must be three [B,H] tensors")
            pooled2d, _, _ = self._prepare(tuple(pooled_features), None)
        else:
            pooled2d = None if t.dim() == 3 else [t, a, v]
        mt = self.mult_fusion
        if pooled2d is None:
            # [B,L,H] inputs and no pooled features given: the 2-D heads see the (masked) mean over L, which MulT's staging pass
            # produces from the same read of the sequences; its gradient joins MulT's input gradient inside the chunk backward
            mult_pooled, cat = mt._pooled(t, a, v, mask, want_mean=True)
            feats = ops.Split3Fn.apply(cat)
        else:
            mult_pooled = mt._pooled(t, a, v, mask)
            cat = _masked_cat(*pooled2d, mask)                               # shared by early / graph / contrastive / adaptive
            feats = ops.Split3Fn.apply(cat) if mask is not None else pooled2d
        early = self.early_fusion._run(cat)
        ff = mt.final_fusion[0]
        mult = ops.dropout(ops.linear(mult_pooled, ff.weight, ff.bias, relu=True), self._p, self.training)
        graph = self.graph_fusion._run(cat, dt)
        con = self.contrastive_fusion._run(cat, feats, compute_contrastive_loss)
        ada = self.adaptive_fusion._run(cat, feats)
        H = self.config.fusion_hidden_size
        allf = _Concat5Fn.apply(early, mult, graph, con["fused_features"], ada["fused_features"])
        m0, m3 = self.meta_fusion[0], self.meta_fusion[3]
        hid = ops.dropout(ops.linear(allf, m0.weight, m0.bias, relu=True), self._p, self.training)
        fused = ops.linear(hid, m3.weight, m3.bias)
        return {"fused_features": fused, "early_features": early, "mult_features": mult, "graph_features": graph,
                "contrastive_features": con["fused_features"], "adaptive_features": ada["fused_features"],
                "contrastive_losses": con["contrastive_losses"], "attention_weights": ada["attention_weights"],
                "adaptive_weights": ada["adaptive_weights"]}


class _Concat5Fn(torch.autograd.Function):
    """cat of the five head outputs [B,H] -> [B,5H]: device-to-device row copies into one buffer (layout plumbing)."""

    @staticmethod
    def forward(ctx, *xs):
        B, H = xs[0].shape
        out = torch.empty((B, len(xs) * H), device=xs[0].device, dtype=xs[0].dtype)
        for i, x in enumerate(xs):
            out[:, i * H:(i + 1) * H].copy_(x)
        ctx.H = H
        return out

    @staticmethod
    def backward(ctx, g):
        H = ctx.H
        return tuple(g[:, i * H:(i + 1) * H].contiguous() for i in range(g.size(1) // H))


class ModalityDropout(nn.Module):
    """reference models/encoders.py:280-321, sync-free: the keep-mask (with the keep-at-least-one repair) is
    generated on the device by a counter-based RNG; pass `mask=` to a fusion head to fuse the multiply, or
    call this module for the reference's (text, audio, video) -> masked features behaviour."""

    def __init__(self, dropout_rate: float = 0.1, seed: int = 4321):
        super().__init__()
        self.dropout_rate, self.seed, self._offset = dropout_rate, seed, 0

    def sample_mask(self, batch: int, device) -> Tensor:
        m = K.modality_mask(batch, self.dropout_rate, self.seed, self._offset, device)
        self._offset += 4 * batch
        return m

    def forward(self, text_features, audio_features, video_features, training: bool = True):
        if not training:
            return text_features, audio_features, video_features
        mask = self.sample_mask(text_features.size(0), text_features.device)
        if text_features.dim() == 2:
            return ops.Split3Fn.apply(ops.Concat3Fn.apply(text_features, audio_features, video_features, mask))
        return tuple(ops.RowMaskFn.apply(x, mask, i) for i, x in enumerate((text_features, audio_features, video_features)))
