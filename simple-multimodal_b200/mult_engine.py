"""MulT (MultimodalTransformer, reference models/fusion_layers.py:93-211) as one chunked
autograd.Function with a hand-scheduled forward and backward.

Why one Function: MulT is >99 % of the hierarchical head's FLOPs and its activations at the
benchmark size (B=4096, L=512/512/30, H=512) are ~29 MB per sample.  MulT is per-sample
independent, so the batch is processed in chunks of `chunk` samples: forward keeps only the pooled
[B,3H] output; backward recomputes a chunk's forward into a bounded stash and immediately runs its
backward, accumulating parameter gradients in fp32.  Kernel-level scheduling per chunk:

  * each modality's four projections that share an input (Q of its two query blocks, K/V of the two
    blocks it serves as key/value) run as ONE GEMM with N = 6H over a stacked weight;
  * attention reads Q/K/V as column slices of those packed projections and writes dQ/dK/dV straight
    into the packed gradient, which then needs one dgrad + one wgrad GEMM per modality;
  * out-proj / FFN2 GEMMs add the residual in their epilogue; the FFN2 input-gradient GEMM applies the
    ReLU mask in its epilogue; LayerNorm of the second block of each modality adds the 3-way residual;
  * mean-pool commutes with the self-attention out-projection (linear), so the pooled [B,H] rows are
    projected instead of all B*L tokens.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

from . import kernels as K
from .ops import LN_EPS, mha_scale, operand

Tensor = torch.Tensor

MODS = ("text", "audio", "video")
# (block name, query modality, key/value modality) in the reference's order (fusion_layers.py:146-153)
BLOCKS = (("text_to_audio", 0, 1), ("text_to_video", 0, 2), ("audio_to_text", 1, 0),
          ("audio_to_video", 1, 2), ("video_to_text", 2, 0), ("video_to_audio", 2, 1))
BLOCK_PARAMS = ("attention.in_proj_weight", "attention.in_proj_bias", "attention.out_proj.weight", "attention.out_proj.bias",
                "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias", "ffn.0.weight", "ffn.0.bias", "ffn.3.weight", "ffn.3.bias")
SELF_PARAMS = ("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias")


def param_names() -> List[str]:
    names = [f"{b}.{p}" for b, _, _ in BLOCKS for p in BLOCK_PARAMS]
    names += [f"{m}_self_attn.{p}" for m in MODS for p in SELF_PARAMS]
    return names


def _layout(H: int):
    """Column layout of the packed per-modality projection [*, 6H]:
    [ Q(block A) | Q(block B) | K,V(block C) | K,V(block D) ], A,B = blocks querying from this modality,
    C,D = blocks that use it as key/value."""
    q_slot: Dict[str, int] = {}
    kv_slot: Dict[str, int] = {}
    order = {m: [] for m in range(3)}
    for name, qm, _ in BLOCKS:
        q_slot[name] = len(order[qm]) * H
        order[qm].append(("q", name))
    for name, _, km in BLOCKS:
        n_q = sum(1 for kind, _ in order[km] if kind == "q")
        n_kv = sum(1 for kind, _ in order[km] if kind == "kv")
        kv_slot[name] = n_q * H + n_kv * 2 * H
        order[km].append(("kv", name))
    return q_slot, kv_slot, order


class _Weights:
    """Per-step operand copies (stacked projections, bf16 casts) built once and shared by all chunks."""

    def __init__(self, P: Dict[str, Tensor], H: int, dtype: torch.dtype):
        self.P = P
        self.q_slot, self.kv_slot, self.order = _layout(H)
        self.w_stack, self.b_stack = [], []
        for m in range(3):
            ws, bs = [], []
            for kind, name in self.order[m]:
                w, b = P[f"{name}.attention.in_proj_weight"].detach(), P[f"{name}.attention.in_proj_bias"].detach()
                ws.append(w[:H] if kind == "q" else w[H:])
                bs.append(b[:H] if kind == "q" else b[H:])
            self.w_stack.append(operand(torch.cat(ws, 0), dtype))
            self.b_stack.append(torch.cat(bs, 0).contiguous())
        self.op = {k: operand(v, dtype) for k, v in P.items() if v.dim() == 2}

    def w(self, name: str) -> Tensor:
        return self.op[name]

    def f32(self, name: str) -> Tensor:
        return self.P[name].detach()


def _site(drop, site: int):
    """Dropout arguments of one site of a chunk: drop = None or (p, seed_lo, seed_hi of this chunk).  Sites: 2*block = attention
    probabilities of cross block `block`, 2*block + 1 = its FFN hidden layer, 12 + m = self-attention of modality m."""
    if drop is None:
        return None
    p, lo, hi = drop
    return p, (lo + 0x632BE5AB * (site + 1)) & 0xFFFFFFFF, hi


def _chunk_forward(xs, W: _Weights, H: int, heads: int, pooled_out: Tensor, keep: bool, drop=None):
    """Forward of one chunk.  xs: 3 contiguous [Bc,L,H] tensors.  Writes the pooled attended features into
    pooled_out [Bc,3H]; returns the stash needed by `_chunk_backward` when `keep`.  `drop` (training with fusion_dropout > 0):
    (p, seed_lo, seed_hi) of this chunk -- attention-probability dropout inside the attention kernels and FFN hidden dropout in
    the FFN1 GEMM epilogue, both regenerated (not stored) by backward and by the recomputed forward."""
    scale = mha_scale(H, heads)
    Bc = xs[0].size(0)
    Ls = [x.size(1) for x in xs]
    x2 = [x.reshape(-1, H) for x in xs]
    st = {"x2": x2, "Ls": Ls} if keep else None
    proj = [K.linear_fwd(x2[m], W.w_stack[m], W.b_stack[m]).view(Bc, Ls[m], 6 * H) for m in range(3)]
    blk_out = {}
    first_of = {}
    enhanced = [None, None, None]
    for bi, (name, qm, km) in enumerate(BLOCKS):
        qo, ko = W.q_slot[name], W.kv_slot[name]
        ctx_, lse = K.attn_fwd(proj[qm][:, :, qo:qo + H], proj[km][:, :, ko:ko + H], proj[km][:, :, ko + H:ko + 2 * H], heads, scale,
                               dropout=_site(drop, 2 * bi))
        ctx2 = ctx_.view(-1, H)
        s1 = K.linear_fwd(ctx2, W.w(f"{name}.attention.out_proj.weight"), W.f32(f"{name}.attention.out_proj.bias"), residual=x2[qm])
        x1, mean1, rstd1 = K.layernorm_fwd(s1, W.f32(f"{name}.norm1.weight"), W.f32(f"{name}.norm1.bias"), LN_EPS)
        hid = K.linear_fwd(x1, W.w(f"{name}.ffn.0.weight"), W.f32(f"{name}.ffn.0.bias"), relu=True, dropout=_site(drop, 2 * bi + 1))
        s2 = K.linear_fwd(hid, W.w(f"{name}.ffn.3.weight"), W.f32(f"{name}.ffn.3.bias"), residual=x1)
        if qm not in first_of:      # first block of this query modality: plain LN2
            first_of[qm] = name
            y, mean2, rstd2 = K.layernorm_fwd(s2, W.f32(f"{name}.norm2.weight"), W.f32(f"{name}.norm2.bias"), LN_EPS)
            blk_out[name] = y
        else:                       # second block: LN2 + first block's output + the input  (3-way residual, :156-158)
            y, mean2, rstd2 = K.layernorm_fwd(s2, W.f32(f"{name}.norm2.weight"), W.f32(f"{name}.norm2.bias"), LN_EPS,
                                              post1=blk_out[first_of[qm]], post2=x2[qm])
            enhanced[qm] = y
        if keep:
            st[name] = dict(ctx=ctx_, lse=lse, s1=s1, x1=x1, mean1=mean1, rstd1=rstd1, hid=hid, s2=s2, mean2=mean2, rstd2=rstd2)
    if keep:
        st["proj"] = proj
    for m, mod in enumerate(MODS):
        pre = f"{mod}_self_attn."
        qkv = K.linear_fwd(enhanced[m], W.w(pre + "in_proj_weight"), W.f32(pre + "in_proj_bias")).view(Bc, Ls[m], 3 * H)
        att, lse = K.attn_fwd(qkv[:, :, :H], qkv[:, :, H:2 * H], qkv[:, :, 2 * H:], heads, scale, dropout=_site(drop, 12 + m))
        pooled_ctx = K.meanpool_fwd(att)
        K.linear_fwd(pooled_ctx, W.w(pre + "out_proj.weight"), W.f32(pre + "out_proj.bias"), out=pooled_out[:, m * H:(m + 1) * H])
        if keep:
            st[mod] = dict(enh=enhanced[m], qkv=qkv, att=att, lse=lse, pooled_ctx=pooled_ctx)
    return st


def _chunk_backward(st, W: _Weights, H: int, heads: int, dpooled: Tensor, G: Dict[str, Tensor], dstack_w, dstack_b, need_dx: bool, drop=None):
    """Backward of one chunk.  dpooled [Bc,3H]; G: fp32 gradient accumulators keyed like the parameters;
    dstack_w/dstack_b: accumulators of the stacked projections.  Returns [dx_text, dx_audio, dx_video] or None."""
    scale = mha_scale(H, heads)
    x2, Ls, proj = st["x2"], st["Ls"], st["proj"]
    Bc = dpooled.size(0)
    dt = x2[0].dtype
    d_enh = [None, None, None]
    for m, mod in enumerate(MODS):
        pre, s = f"{mod}_self_attn.", st[mod]
        g = dpooled[:, m * H:(m + 1) * H]
        K.linear_wgrad(g, s["pooled_ctx"], G[pre + "out_proj.weight"])
        K.colsum_accum(g, G[pre + "out_proj.bias"])
        d_pc = K.linear_dgrad(g, W.w(pre + "out_proj.weight"))
        d_att = K.meanpool_bwd(d_pc, Ls[m])
        dqkv = torch.empty_like(s["qkv"])
        qkv = s["qkv"]
        db = G[pre + "in_proj_bias"]        # Q / V bias gradients are summed in the attention-backward epilogue; the K-bias
        K.attn_bwd(d_att, qkv[:, :, :H], qkv[:, :, H:2 * H], qkv[:, :, 2 * H:], s["att"], s["lse"], heads, scale,   # gradient is
                   dqkv[:, :, :H], dqkv[:, :, H:2 * H], dqkv[:, :, 2 * H:], dbq=db[:H], dbv=db[2 * H:],           # identically 0 (softmax shift invariance)
                   dropout=_site(drop, 12 + m))
        dqkv2 = dqkv.view(-1, 3 * H)
        K.linear_wgrad(dqkv2, s["enh"], G[pre + "in_proj_weight"])
        d_enh[m] = K.linear_dgrad(dqkv2, W.w(pre + "in_proj_weight"))
    dproj = [torch.empty_like(p) for p in proj]
    d_s1 = {}
    inv_keep = 1.0 if drop is None else 1.0 / (1.0 - drop[0])
    for bi, (name, qm, km) in enumerate(BLOCKS):
        s = st[name]
        # d(enhanced) reaches both blocks' LN2 outputs and the input unchanged
        ds2 = K.layernorm_bwd(d_enh[qm], s["s2"], s["mean2"], s["rstd2"], W.f32(f"{name}.norm2.weight"),
                              G[f"{name}.norm2.weight"], G[f"{name}.norm2.bias"], dxsum=G[f"{name}.ffn.3.bias"])   # + bias grad of FFN2
        K.linear_wgrad(ds2, s["hid"], G[f"{name}.ffn.3.weight"])
        dhid = K.linear_dgrad(ds2, W.w(f"{name}.ffn.3.weight"), relu_mask=s["hid"],       # ReLU' and the FFN1 bias gradient
                              colsum=G[f"{name}.ffn.0.bias"], alpha=inv_keep)              # (column sums) fused in the epilogue;
        # with dropout the stored hidden is zero where dropped, so the same mask covers it and alpha carries 1/(1-p)
        K.linear_wgrad(dhid, s["x1"], G[f"{name}.ffn.0.weight"])
        dx1 = K.linear_dgrad(dhid, W.w(f"{name}.ffn.0.weight"), residual=ds2)             # + residual path x1 -> s2
        ds1 = K.layernorm_bwd(dx1, s["s1"], s["mean1"], s["rstd1"], W.f32(f"{name}.norm1.weight"),
                              G[f"{name}.norm1.weight"], G[f"{name}.norm1.bias"], dxsum=G[f"{name}.attention.out_proj.bias"])
        d_s1[name] = ds1
        ctx2 = s["ctx"].view(-1, H)
        K.linear_wgrad(ds1, ctx2, G[f"{name}.attention.out_proj.weight"])
        dctx = K.linear_dgrad(ds1, W.w(f"{name}.attention.out_proj.weight")).view(Bc, Ls[qm], H)
        qo, ko = W.q_slot[name], W.kv_slot[name]
        K.attn_bwd(dctx, proj[qm][:, :, qo:qo + H], proj[km][:, :, ko:ko + H], proj[km][:, :, ko + H:ko + 2 * H], s["ctx"], s["lse"],
                   heads, scale, dproj[qm][:, :, qo:qo + H], dproj[km][:, :, ko:ko + H], dproj[km][:, :, ko + H:ko + 2 * H],
                   dbq=dstack_b[qm][qo:qo + H], dbv=dstack_b[km][ko + H:ko + 2 * H], dropout=_site(drop, 2 * bi))
    dxs = []
    for m in range(3):
        dp2 = dproj[m].view(-1, 6 * H)
        K.linear_wgrad(dp2, x2[m], dstack_w[m])
        if need_dx:
            a_name, b_name = [n for n, qm, _ in BLOCKS if qm == m]
            direct = K.add(d_enh[m], d_s1[a_name], d_s1[b_name])      # residual paths into the input
            dxs.append(K.linear_dgrad(dp2, W.w_stack[m], residual=direct).view(Bc, Ls[m], H))
    return dxs if need_dx else None


def stash_bytes_per_sample(Ls, H: int, elem: int) -> int:
    """Activation bytes `_chunk_forward(keep=True)` holds per sample (packed projections, per-block ctx/s1/x1/hid/s2,
    enhanced, self-attention qkv/ctx) -- used to decide how many chunks can stay resident between forward and backward."""
    tok = sum(Ls)
    per_tok = 6 * H + H + 3 * H + H                 # packed projection, enhanced, self qkv, self ctx
    per_qtok = 2 * (H + H + H + 4 * H + H)          # two query blocks per token: ctx, s1, x1, hid, s2
    return (per_tok + per_qtok) * tok * elem


def _chunk_drop(drop, ci: int):
    """(p, seed_lo, seed_hi) of chunk `ci`: every chunk (rows restart at 0 inside a chunk) gets its own high seed word."""
    if drop is None:
        return None
    p, lo, hi = drop
    return p, lo, (hi + 0x9E3779B1 * (ci + 1)) & 0xFFFFFFFF


class MulTFn(torch.autograd.Function):
    """(text, audio, video [B,L,H]) + MulT parameters -> pooled attended features [B,3H].

    The batch runs in chunks of `chunk` samples.  Chunks whose activations fit in `stash_budget` bytes stay resident
    for backward; the remaining chunks are recomputed chunk by chunk in backward (bounded memory at any batch)."""

    @staticmethod
    def forward(ctx, t, a, v, H, heads, chunk, stash_budget, drop, names, *params):
        xs = [x.contiguous() for x in (t, a, v)]
        B = xs[0].size(0)
        P = dict(zip(names, params))
        W = _Weights(P, H, xs[0].dtype)
        pooled = torch.empty((B, 3 * H), device=t.device, dtype=t.dtype)
        need_grad = any(ctx.needs_input_grad)
        per_chunk = stash_bytes_per_sample([x.size(1) for x in xs], H, xs[0].element_size()) * min(chunk, B)
        n_keep = int(stash_budget // max(per_chunk, 1)) if need_grad else 0
        stash = {}
        for ci, b0 in enumerate(range(0, B, chunk)):
            b1 = min(B, b0 + chunk)
            keep = ci < n_keep
            st = _chunk_forward([x[b0:b1] for x in xs], W, H, heads, pooled[b0:b1], keep=keep, drop=_chunk_drop(drop, ci))
            if keep:
                stash[b0] = st
        ctx.cfg = (H, heads, chunk, names, drop)
        ctx.W, ctx.stash, ctx.xs = (W if need_grad else None), stash, (xs if need_grad else None)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        H, heads, chunk, names, drop = ctx.cfg
        W, xs = ctx.W, ctx.xs
        B = xs[0].size(0)
        dev = xs[0].device
        dpooled = dpooled.contiguous()
        # fp32 gradient accumulators of every parameter (+ the stacked projections) as views of ONE zeroed buffer: one fill kernel
        # instead of ~100; every view starts on a 256-byte boundary (the GEMM / attention epilogues use 16-byte vector REDs)
        shapes = [tuple(W.P[n].shape) for n in names] + [tuple(w.shape) for w in W.w_stack] + [tuple(b.shape) for b in W.b_stack]
        offs, total = [], 0
        for shp in shapes:
            offs.append(total)
            total += (math.prod(shp) + 63) // 64 * 64
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        views = [flat[o:o + math.prod(shp)].view(shp) for o, shp in zip(offs, shapes)]
        G = dict(zip(names, views[:len(names)]))
        dstack_w = views[len(names):len(names) + len(W.w_stack)]
        dstack_b = views[len(names) + len(W.w_stack):]
        need_dx = any(ctx.needs_input_grad[:3])
        dxs = [torch.empty_like(x) for x in xs] if need_dx else None
        scratch = None
        for b0 in reversed(range(0, B, chunk)):        # recomputed (late) chunks first, then the resident ones are released in turn
            b1 = min(B, b0 + chunk)
            st = ctx.stash.pop(b0, None)
            if st is None:
                if scratch is None:
                    scratch = torch.empty((min(chunk, B), 3 * H), device=dev, dtype=xs[0].dtype)
                st = _chunk_forward([x[b0:b1] for x in xs], W, H, heads, scratch[:b1 - b0], keep=True, drop=_chunk_drop(drop, b0 // chunk))
            out = _chunk_backward(st, W, H, heads, dpooled[b0:b1], G, dstack_w, dstack_b, need_dx, drop=_chunk_drop(drop, b0 // chunk))
            if need_dx:
                for m in range(3):
                    dxs[m][b0:b1].copy_(out[m])
            del st, out
        ctx.stash = None
        # scatter the stacked-projection gradients back onto the blocks' in_proj parameters
        for m in range(3):
            row = 0
            for kind, name in W.order[m]:
                n = H if kind == "q" else 2 * H
                dst_w, dst_b = G[f"{name}.attention.in_proj_weight"], G[f"{name}.attention.in_proj_bias"]
                lo = 0 if kind == "q" else H
                dst_w[lo:lo + n].copy_(dstack_w[m][row:row + n])
                dst_b[lo:lo + n].copy_(dstack_b[m][row:row + n])
                row += n
        grads = [G[n] if ctx.needs_input_grad[9 + i] else None for i, n in enumerate(names)]
        return (dxs[0] if need_dx else None, dxs[1] if need_dx else None, dxs[2] if need_dx else None,
                None, None, None, None, None, None, *grads)
