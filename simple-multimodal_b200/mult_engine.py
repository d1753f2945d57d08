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
from ._lib import B200FusionError
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
    """Contraction operands of the MulT parameters (stacked projections, bf16 casts) at STATIC addresses, shared by all chunks
    and all steps.  `refresh()` re-derives them from the fp32 masters only when a parameter changed (tensor version counters,
    bumped by every in-place optimizer update), so a step whose weights did not move issues no cast kernel and the captured
    chunk graphs (`ChunkGraphEngine`) keep reading the same buffers."""

    def __init__(self, P: Dict[str, Tensor], H: int, dtype: torch.dtype):
        self.P, self.H, self.dtype = P, H, dtype
        self.q_slot, self.kv_slot, self.order = _layout(H)
        dev = next(iter(P.values())).device
        self.w_stack = [torch.empty((6 * H, H), device=dev, dtype=dtype) for _ in range(3)]
        self.b_stack = [torch.empty(6 * H, device=dev, dtype=torch.float32) for _ in range(3)]
        self.op = {k: torch.empty(tuple(v.shape), device=dev, dtype=dtype) for k, v in P.items() if v.dim() == 2}
        self._key = None
        self.refresh()

    def _fingerprint(self):
        return tuple((v.data_ptr(), v._version) for v in self.P.values())

    def refresh(self, force: bool = False) -> None:
        key = self._fingerprint()
        if key == self._key and not force:
            return
        H, P = self.H, self.P

        def put(dst: Tensor, src: Tensor):
            if dst.dtype == torch.bfloat16:
                K.cast_to_bf16(src.detach(), out=dst)
            else:
                dst.copy_(src.detach())

        for m in range(3):
            row = 0
            for kind, name in self.order[m]:
                w, b = P[f"{name}.attention.in_proj_weight"], P[f"{name}.attention.in_proj_bias"]
                lo, n = (0, H) if kind == "q" else (H, 2 * H)
                put(self.w_stack[m][row:row + n], w[lo:lo + n])
                self.b_stack[m][row:row + n].copy_(b.detach()[lo:lo + n])
                row += n
        for k, dst in self.op.items():
            put(dst, P[k])
        self._key = key

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
        # the ReLU' (and dropout) mask of the hidden layer leaves the FFN1 epilogue as one bit per element: the FFN2 input-gradient
        # GEMM reads 1/16 of the bytes it needed when it looked at the stored hidden layer itself
        hbits = K.sign_bits_for(x1, 4 * H) if keep else None
        hid = K.linear_fwd(x1, W.w(f"{name}.ffn.0.weight"), W.f32(f"{name}.ffn.0.bias"), relu=True, dropout=_site(drop, 2 * bi + 1),
                           sign_bits_out=hbits)
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
            st[name] = dict(ctx=ctx_, lse=lse, s1=s1, x1=x1, mean1=mean1, rstd1=rstd1, hid=hid, hbits=hbits, s2=s2, mean2=mean2, rstd2=rstd2)
    if keep:
        st["proj"] = proj
    for m, mod in enumerate(MODS):
        pre = f"{mod}_self_attn."
        qkv = K.linear_fwd(enhanced[m], W.w(pre + "in_proj_weight"), W.f32(pre + "in_proj_bias")).view(Bc, Ls[m], 3 * H)
        q_ = qkv[:, :, :H]
        # the mean over L of the attended features comes out of the attention epilogue (column sums of the stored O) where the
        # kernel family supports it, instead of a pass that re-reads [Bc, L, H]
        pooled_ctx = torch.empty((Bc, H), device=q_.device, dtype=q_.dtype) if K.attn_pool_supported(q_, heads) else None
        att, lse = K.attn_fwd(q_, qkv[:, :, H:2 * H], qkv[:, :, 2 * H:], heads, scale, dropout=_site(drop, 12 + m), pooled=pooled_ctx)
        if pooled_ctx is None:
            pooled_ctx = K.meanpool_fwd(att)
        K.linear_fwd(pooled_ctx, W.w(pre + "out_proj.weight"), W.f32(pre + "out_proj.bias"), out=pooled_out[:, m * H:(m + 1) * H])
        if keep:
            st[mod] = dict(enh=enhanced[m], qkv=qkv, att=att, lse=lse, pooled_ctx=pooled_ctx)
    return st


def _chunk_backward(st, W: _Weights, H: int, heads: int, dpooled: Tensor, G: Dict[str, Tensor], dstack_w, dstack_b, dx_out, drop=None, dmean=None):
    """Backward of one chunk.  dpooled [Bc,3H]; G: fp32 gradient accumulators keyed like the parameters;
    dstack_w/dstack_b: accumulators of the stacked projections; dx_out: None, or the three contiguous [Bc,L,H] slices of the
    input-gradient buffers this chunk's input gradients are written into (the last dgrad GEMM stores there directly);
    dmean: None or [Bc,3H], the gradient of the mean-over-L of the inputs (MulTFn's second output) -- added to the residual paths."""
    need_dx = dx_out is not None
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
        dhid = K.linear_dgrad(ds2, W.w(f"{name}.ffn.3.weight"),                           # ReLU' and the FFN1 bias gradient
                              relu_mask=s["hid"] if s["hbits"] is None else None, sign_bits=s["hbits"],   # (column sums) fused in the
                              colsum=G[f"{name}.ffn.0.bias"], alpha=inv_keep)                             # epilogue;
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
    for m in range(3):
        dp2 = dproj[m].view(-1, 6 * H)
        K.linear_wgrad(dp2, x2[m], dstack_w[m])
        if need_dx:
            a_name, b_name = [n for n, qm, _ in BLOCKS if qm == m]
            if dmean is None:
                direct = K.add(d_enh[m], d_s1[a_name], d_s1[b_name])      # residual paths into the input
            else:                                                         # ... plus dmean / L broadcast over the sample's tokens
                direct = K.add_rowbcast(d_enh[m], d_s1[a_name], d_s1[b_name], dmean[:, m * H:(m + 1) * H], 1.0 / Ls[m], Ls[m])
            K.linear_dgrad(dp2, W.w_stack[m], residual=direct, out=dx_out[m].view(-1, H))


def stash_bytes_per_sample(Ls, H: int, elem: int) -> int:
    """Activation bytes `_chunk_forward(keep=True)` holds per sample (packed projections, per-block ctx/s1/x1/hid/s2,
    enhanced, self-attention qkv/ctx) -- used to decide how many chunks can stay resident between forward and backward."""
    tok = sum(Ls)
    per_tok = 6 * H + H + 3 * H + H                 # packed projection, enhanced, self qkv, self ctx
    per_qtok = 2 * (H + H + H + 4 * H + H)          # two query blocks per token: ctx, s1, x1, hid, s2
    bits = 2 * (4 * H // 8) if elem == 2 else 0     # one-bit ReLU' masks of the two hidden layers (bf16 path)
    return (per_tok + per_qtok) * tok * elem + bits * tok


def _chunk_drop(drop, ci: int):
    """(p, seed_lo, seed_hi) of chunk `ci`: every chunk (rows restart at 0 inside a chunk) gets its own high seed word."""
    if drop is None:
        return None
    p, lo, hi = drop
    return p, lo, (hi + 0x9E3779B1 * (ci + 1)) & 0xFFFFFFFF


def _grad_buffers(W: _Weights, names, dev):
    """fp32 gradient accumulators of every parameter (+ the stacked projections) as views of ONE buffer: one fill kernel
    instead of ~100; every view starts on a 256-byte boundary (the GEMM / attention epilogues use 16-byte vector REDs)."""
    shapes = [tuple(W.P[n].shape) for n in names] + [tuple(w.shape) for w in W.w_stack] + [tuple(b.shape) for b in W.b_stack]
    offs, total = [], 0
    for shp in shapes:
        offs.append(total)
        total += (math.prod(shp) + 63) // 64 * 64
    flat = torch.empty(total, device=dev, dtype=torch.float32)
    views = [flat[o:o + math.prod(shp)].view(shp) for o, shp in zip(offs, shapes)]
    G = dict(zip(names, views[:len(names)]))
    dstack_w = views[len(names):len(names) + len(W.w_stack)]
    dstack_b = views[len(names) + len(W.w_stack):]
    return flat, G, dstack_w, dstack_b


def _scatter_stacked(W: _Weights, H: int, G, dstack_w, dstack_b) -> None:
    """the stacked-projection gradients go back onto the blocks' in_proj parameters"""
    for m in range(3):
        row = 0
        for kind, name in W.order[m]:
            n = H if kind == "q" else 2 * H
            lo = 0 if kind == "q" else H
            G[f"{name}.attention.in_proj_weight"][lo:lo + n].copy_(dstack_w[m][row:row + n])
            G[f"{name}.attention.in_proj_bias"][lo:lo + n].copy_(dstack_b[m][row:row + n])
            row += n


class ChunkGraphEngine:
    """The MulT step of one (batch, sequence lengths) shape as CUDA graphs: one captured forward and one captured backward PER
    CHUNK, replayed back to back (reference models/fusion_layers.py:93-179 is per-sample independent, so chunks never interact).

    Why: a chunk's forward + backward is ~150 kernel launches issued from Python (~18 ms of host time), about what the B200
    needs to execute them; at B = 4096 (16 chunks) the step was issued in ~300 ms for ~330 ms of GPU work, and several ranks on a
    CPU-limited host made it host-bound.  Replaying 32 graphs costs well under a millisecond of host time per step.

    Everything a graph touches lives at a static address: the (masked) inputs `xs`, the pooled output, `dpooled`, the input
    gradients `dx`, the fp32 parameter-gradient accumulators, the operand copies of the weights (`_Weights`) and, per chunk, the
    activation stash its backward graph reads -- the same ~29 MB per sample the eager path keeps, held in the graphs' shared
    private pool.  Dropout: the per-chunk seeds are frozen in the graphs; the device-side epoch (csrc/common.cuh) is raised by
    this step's number around the forward replays and around the backward replays (and lowered again afterwards, so eager kernels
    issued between them see the epoch they started with), giving every step fresh masks that forward and backward agree on.

    One forward may be in flight: a second forward overwrites the static buffers, and the first one's backward then raises."""

    def __init__(self, W: _Weights, names, H: int, heads: int, chunk: int, B: int, Ls, dtype, dev, drop_p: float, need_dx: bool,
                 want_mean: bool = False):
        from .ops import next_drop_seed
        self.W, self.names, self.H, self.heads, self.chunk, self.B, self.Ls = W, list(names), H, heads, chunk, B, list(Ls)
        self.need_dx, self.drop_p, self.want_mean = need_dx, drop_p, want_mean
        self.dmean = torch.zeros((B, 3 * H), device=dev, dtype=dtype) if (want_mean and need_dx) else None
        self.drop = (drop_p, *next_drop_seed()) if drop_p > 0.0 else None
        self.bounds = [(b0, min(B, b0 + chunk)) for b0 in range(0, B, chunk)]
        self.xs = [torch.empty((B, L, H), device=dev, dtype=dtype) for L in Ls]
        self.dx = [torch.empty_like(x) for x in self.xs] if need_dx else None
        self.pooled = torch.empty((B, 3 * H), device=dev, dtype=dtype)
        self.dpooled = torch.zeros((B, 3 * H), device=dev, dtype=dtype)
        self.flat, self.G, self.dstack_w, self.dstack_b = _grad_buffers(W, names, dev)
        self.step_no, self.live = 0, None
        self.kernel_launches = 0
        self._capture()

    def _fwd(self, ci, keep=True):
        b0, b1 = self.bounds[ci]
        return _chunk_forward([x[b0:b1] for x in self.xs], self.W, self.H, self.heads, self.pooled[b0:b1], keep=keep,
                              drop=_chunk_drop(self.drop, ci))

    def _bwd(self, ci, st):
        b0, b1 = self.bounds[ci]
        _chunk_backward(st, self.W, self.H, self.heads, self.dpooled[b0:b1], self.G, self.dstack_w, self.dstack_b,
                        [d[b0:b1] for d in self.dx] if self.need_dx else None, drop=_chunk_drop(self.drop, ci),
                        dmean=None if self.dmean is None else self.dmean[b0:b1])

    def _capture(self):
        from . import _lib
        for x in self.xs:
            x.zero_()
        self.flat.zero_()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                       # every kernel configures itself (opt-in shared memory) on first use:
            last = len(self.bounds) - 1                     # run the two distinct chunk shapes once outside capture
            for ci in sorted({0, last}):
                st = self._fwd(ci)
                self._bwd(ci, st)
                del st
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.pool = torch.cuda.graph_pool_handle()
        self.fwd_graphs, self.bwd_graphs, self.stash = [], [None] * len(self.bounds), []
        n0 = _lib.launch_count()
        for ci in range(len(self.bounds)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
                st = self._fwd(ci)
            self.fwd_graphs.append(g)
            self.stash.append(st)
        for ci in reversed(range(len(self.bounds))):        # captured in replay order
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.pool, capture_error_mode="thread_local"):
                self._bwd(ci, self.stash[ci])
            self.bwd_graphs[ci] = g
        self.final_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.final_graph, pool=self.pool, capture_error_mode="thread_local"):
            _scatter_stacked(self.W, self.H, self.G, self.dstack_w, self.dstack_b)
        self.kernel_launches = _lib.launch_count() - n0     # library kernels replayed by one forward + backward
        self._views = [self.G[n] for n in self.names]       # held here so that autograd copies (never adopts) them into .grad

    def _epoch(self, delta: int) -> None:
        if self.drop is not None:
            K.dropout_epoch(delta & 0xFFFFFFFF, add=True)

    def forward(self, xs, mask):
        self.W.refresh()
        xmean = torch.empty((self.B, 3 * self.H), device=self.pooled.device, dtype=self.pooled.dtype) if self.want_mean else None
        for m in range(3):                                  # one read of the inputs: masked copy into the static buffers (+ their mean over L)
            if self.want_mean:
                K.stage_pool(xs[m], self.xs[m], xmean[:, m * self.H:(m + 1) * self.H], mask, m)
            else:
                K.rowmask_copy(xs[m], self.xs[m], mask, m)
        self.step_no += 1
        self.live = self.step_no
        self._epoch(self.step_no)
        for g in self.fwd_graphs:
            g.replay()
        self._epoch(-self.step_no)
        return self.pooled.clone(), xmean, self.step_no

    def backward(self, token: int, dpooled: Tensor, mask, dxmean=None):
        if token != self.live:
            raise B200FusionError("MulT chunk graphs: the static activation buffers of this forward were overwritten by a later forward "
                                    "(one forward may be in flight; set MultimodalTransformer.graph_chunks = False for several)")
        self.dpooled.copy_(dpooled)
        if self.dmean is not None:
            if dxmean is None:
                self.dmean.zero_()
            else:
                self.dmean.copy_(dxmean)
        self.flat.zero_()
        self._epoch(token)
        for ci in reversed(range(len(self.bounds))):
            self.bwd_graphs[ci].replay()
        self._epoch(-token)
        self.final_graph.replay()
        if self.need_dx and mask is not None:
            for m in range(3):
                K.rowmask_apply_(self.dx[m], mask, m)
        return self.dx, self._views

    def release(self) -> None:
        """drop the graphs and their pool (tens of GB): the engine is unusable afterwards"""
        self.fwd_graphs = self.bwd_graphs = self.final_graph = self.stash = None
        self.live = None


class MulTFn(torch.autograd.Function):
    """(text, audio, video [B,L,H]) [+ modality keep-mask [B,3]] + MulT parameters -> (pooled attended features [B,3H], xmean).

    The modality-dropout multiply (models/encoders.py:317-319) is applied here: on the copy into the engine's static input
    buffers, or (eager path) by one masked copy; backward masks the input gradients in place.  With `want_mean` the second output
    is [mask_t * mean_L(t) | mask_a * mean_L(a) | mask_v * mean_L(v)] ([B,3H]), the input of HierarchicalFusion's 2-D heads
    (SURVEY F3), computed by the same single read of the inputs; its gradient joins the residual paths inside the last add of the
    chunk backward instead of materialising a second [B,L,H] gradient.  Without `want_mean` it is an empty placeholder.
    `engine` (a ChunkGraphEngine for exactly this shape) replays captured chunk graphs; without it the batch runs in eagerly
    issued chunks of `chunk` samples: chunks whose activations fit in `stash_budget` bytes stay resident for backward, the
    remaining chunks are recomputed chunk by chunk in backward (bounded memory at any batch)."""

    @staticmethod
    def forward(ctx, t, a, v, mask, want_mean, H, heads, chunk, stash_budget, drop, names, W, engine, *params):
        ctx.mask, ctx.engine, ctx.want_mean = mask, engine, want_mean
        need_grad = any(ctx.needs_input_grad)
        ctx.cfg = (H, heads, chunk, names, drop)
        if engine is not None:
            pooled, xmean, ctx.token = engine.forward([x.contiguous() for x in (t, a, v)], mask)
        else:
            B = t.size(0)
            xin = [x.contiguous() for x in (t, a, v)]
            xmean = torch.empty((B, 3 * H), device=t.device, dtype=t.dtype) if want_mean else None
            if mask is None:
                xs = xin
                if want_mean:
                    for m in range(3):
                        K.meanpool_fwd(xs[m], out=xmean[:, m * H:(m + 1) * H])
            else:
                xs = [torch.empty(x.shape, device=x.device, dtype=x.dtype) for x in xin]
                for m in range(3):
                    if want_mean:
                        K.stage_pool(xin[m], xs[m], xmean[:, m * H:(m + 1) * H], mask, m)
                    else:
                        K.rowmask_copy(xin[m], xs[m], mask, m)
            W.refresh(force=torch.cuda.is_current_stream_capturing())     # inside a captured step the casts must be part of the graph
            pooled = torch.empty((B, 3 * H), device=t.device, dtype=t.dtype)
            per_chunk = stash_bytes_per_sample([x.size(1) for x in xs], H, xs[0].element_size()) * min(chunk, B)
            n_keep = int(stash_budget // max(per_chunk, 1)) if need_grad else 0
            stash = {}
            for ci, b0 in enumerate(range(0, B, chunk)):
                b1 = min(B, b0 + chunk)
                keep = ci < n_keep
                st = _chunk_forward([x[b0:b1] for x in xs], W, H, heads, pooled[b0:b1], keep=keep, drop=_chunk_drop(drop, ci))
                if keep:
                    stash[b0] = st
            ctx.W, ctx.stash, ctx.xs = (W if need_grad else None), stash, (xs if need_grad else None)
        if xmean is None:
            xmean = pooled.new_empty(0)
            ctx.mark_non_differentiable(xmean)
        return pooled, xmean

    @staticmethod
    def backward(ctx, dpooled, dxmean):
        H, heads, chunk, names, drop = ctx.cfg
        need_dx = any(ctx.needs_input_grad[:3])
        n_fixed = 13                                   # positional arguments before *params
        if not ctx.want_mean:
            dxmean = None
        if ctx.engine is not None:
            dxs, views = ctx.engine.backward(ctx.token, dpooled, ctx.mask, dxmean)
            grads = [g if ctx.needs_input_grad[n_fixed + i] else None for i, g in enumerate(views)]
            dxs = dxs if need_dx else (None, None, None)
            return (dxs[0], dxs[1], dxs[2], *([None] * (n_fixed - 3)), *grads)
        W, xs = ctx.W, ctx.xs
        B = xs[0].size(0)
        dev = xs[0].device
        dpooled = dpooled.contiguous()
        if dxmean is not None:
            dxmean = dxmean.contiguous()
        flat, G, dstack_w, dstack_b = _grad_buffers(W, names, dev)
        flat.zero_()
        dxs = [torch.empty_like(x) for x in xs] if need_dx else None
        scratch = None
        for b0 in reversed(range(0, B, chunk)):        # recomputed (late) chunks first, then the resident ones are released in turn
            b1 = min(B, b0 + chunk)
            st = ctx.stash.pop(b0, None)
            if st is None:
                if scratch is None:
                    scratch = torch.empty((min(chunk, B), 3 * H), device=dev, dtype=xs[0].dtype)
                st = _chunk_forward([x[b0:b1] for x in xs], W, H, heads, scratch[:b1 - b0], keep=True, drop=_chunk_drop(drop, b0 // chunk))
            _chunk_backward(st, W, H, heads, dpooled[b0:b1], G, dstack_w, dstack_b, [d[b0:b1] for d in dxs] if need_dx else None,
                            drop=_chunk_drop(drop, b0 // chunk), dmean=None if dxmean is None else dxmean[b0:b1])
            del st
        ctx.stash = {}
        _scatter_stacked(W, H, G, dstack_w, dstack_b)
        if need_dx and ctx.mask is not None:
            for m in range(3):
                K.rowmask_apply_(dxs[m], ctx.mask, m)
        grads = [G[n] if ctx.needs_input_grad[n_fixed + i] else None for i, n in enumerate(names)]
        dxs = dxs if need_dx else (None, None, None)
        return (dxs[0], dxs[1], dxs[2], *([None] * (n_fixed - 3)), *grads)
