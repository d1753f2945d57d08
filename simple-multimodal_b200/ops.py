"""Differentiable ops of the fusion path: `torch.autograd.Function`s whose forward and backward
both run in libb200fusion.so (via `kernels.py`).  Activations are float32 (CUDA-core parity mode)
or bfloat16 (tcgen05 path); parameters stay float32 masters and receive float32 gradients."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist

from . import kernels as K
from ._lib import B200FusionError

Tensor = torch.Tensor
LN_EPS = 1e-5
L2_EPS = 1e-12


def operand(w: Tensor, dtype: torch.dtype) -> Tensor:
    """fp32 master parameter -> contraction operand in the activation dtype."""
    w = w.detach()
    if dtype == torch.float32:
        return w.contiguous()
    if dtype == torch.bfloat16:
        return K.cast_to_bf16(w)
    raise B200FusionError(f"unsupported compute dtype {dtype}")


def _c(x: Tensor) -> Tensor:
    return x if x.is_contiguous() else x.contiguous()


class LinearFn(torch.autograd.Function):
    """y = [relu](x W^T + b); nn.Linear (+ReLU) on the path."""

    @staticmethod
    def forward(ctx, x, w, b, relu: bool):
        x = _c(x)
        x2 = x.reshape(-1, x.size(-1))
        wc = operand(w, x.dtype)
        y2 = K.linear_fwd(x2, wc, None if b is None else b.detach(), relu=relu)
        ctx.relu, ctx.has_bias = relu, b is not None
        ctx.save_for_backward(x2, wc, y2 if relu else None)
        return y2.reshape(*x.shape[:-1], w.size(0))

    @staticmethod
    def backward(ctx, dy):
        x2, wc, y2 = ctx.saved_tensors
        g = _c(dy).reshape(-1, dy.size(-1))
        if ctx.relu:
            g = K.relu_bwd(g, y2)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = K.linear_dgrad(g, wc).reshape(*dy.shape[:-1], x2.size(1))
        if ctx.needs_input_grad[1]:
            dw = torch.zeros((wc.size(0), wc.size(1)), device=g.device, dtype=torch.float32)
            K.linear_wgrad(g, x2, dw)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(wc.size(0), device=g.device, dtype=torch.float32)
            K.colsum_accum(g, db)
        return dx, dw, db, None


def linear(x, w, b, relu=False):
    return LinearFn.apply(x, w, b, relu)


class DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed, offset):
        ctx.cfg = (p, seed, offset)
        return K.dropout(_c(x), p, seed, offset)

    @staticmethod
    def backward(ctx, dy):
        p, seed, offset = ctx.cfg
        return K.dropout(_c(dy), p, seed, offset), None, None, None


class _DropoutState:
    """Counter-based dropout stream: (seed, running offset) so forward/backward regenerate the same mask."""
    seed = 0x5EED
    offset = 0
    per_rank = True       # fold the data-parallel rank into the seed (every rank draws its own masks, like per-rank torch RNG)


def _rank_seed() -> int:
    """The package seed of THIS rank: with a default process group initialised, rank r uses seed ^ splitmix(r), so
    data-parallel replicas do not share dropout masks (`manual_seed(s)` on every rank still gives distinct streams)."""
    s = _DropoutState.seed
    if _DropoutState.per_rank and dist.is_available() and dist.is_initialized():
        r = dist.get_rank()
        if r:
            z = (r * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
            s ^= (z ^ (z >> 27)) & 0x7FFFFFFFFFFFFFFF
    return s


def dropout(x: Tensor, p: float, training: bool) -> Tensor:
    if not training or p <= 0.0:
        return x
    off = _DropoutState.offset
    _DropoutState.offset += x.numel()
    return DropoutFn.apply(x, p, _rank_seed(), off)


def manual_seed(seed: int, per_rank: bool = True) -> None:
    _DropoutState.seed, _DropoutState.offset, _DropoutState.per_rank = int(seed), 0, bool(per_rank)


def next_drop_seed():
    """A fresh (seed_lo, seed_hi) pair for one in-kernel dropout site (attention probabilities, FFN hidden, GAT / 3-token
    attention weights): splitmix64 of the running offset under the package seed.  The pair fully determines the mask, so
    backward (and the recomputed forward of MulT chunks) regenerate it instead of storing it."""
    _DropoutState.offset += 1
    z = (_rank_seed() * 0x9E3779B97F4A7C15 + _DropoutState.offset * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    z ^= z >> 31
    return z & 0xFFFFFFFF, z >> 32


class Concat3Fn(torch.autograd.Function):
    """[t*m0 | a*m1 | v*m2] (fusion_layers.py:38,350,437 with the modality-dropout multiply folded in)."""

    @staticmethod
    def forward(ctx, t, a, v, mask):
        ctx.save_for_backward(mask)
        ctx.H = t.size(1)
        return K.concat3_fwd(_c(t), _c(a), _c(v), mask)

    @staticmethod
    def backward(ctx, dcat):
        (mask,) = ctx.saved_tensors
        dt, da, dv = K.concat3_bwd(_c(dcat), mask, ctx.H)
        return dt, da, dv, None


class Split3Fn(torch.autograd.Function):
    """Inverse view: cat[B,3H] -> three [B,H] tensors (used to hand masked features to per-modality ops)."""

    @staticmethod
    def forward(ctx, cat):
        H = cat.size(1) // 3
        outs = K.concat3_bwd(_c(cat), None, H)
        return tuple(outs)

    @staticmethod
    def backward(ctx, dt, da, dv):
        return K.concat3_fwd(_c(dt), _c(da), _c(dv), None)


class MeanPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.L = x.size(1)
        return K.meanpool_fwd(_c(x))

    @staticmethod
    def backward(ctx, dy):
        return K.meanpool_bwd(_c(dy), ctx.L)


class WeightedPoolFn(torch.autograd.Function):
    """y[b,:] = sum_l w[b,l] x[b,l,:]; w is a constant fp32 [B,L] (normalised attention mask)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(w)
        return K.weighted_pool_fwd(_c(x), w)

    @staticmethod
    def backward(ctx, dy):
        (w,) = ctx.saved_tensors
        return K.weighted_pool_bwd(_c(dy), w), None


class L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y):
        z, norm = K.l2norm_fwd(_c(y), L2_EPS)
        ctx.save_for_backward(z, norm)
        return z

    @staticmethod
    def backward(ctx, dz):
        z, norm = ctx.saved_tensors
        return K.l2norm_bwd(_c(dz), z, norm, L2_EPS)


class AttentionFn(torch.autograd.Function):
    """softmax(scale Q K^T) V per head on a packed projection: qkv_q [B,Lq,*], qkv_kv [B,Lk,*] with the
    Q / K / V column offsets given (so no slicing copies)."""

    @staticmethod
    def forward(ctx, pq, pkv, q_off, k_off, v_off, H, heads, scale, dropout=None):
        q, k, v = pq[:, :, q_off:q_off + H], pkv[:, :, k_off:k_off + H], pkv[:, :, v_off:v_off + H]
        o, lse = K.attn_fwd(q, k, v, heads, scale, dropout=dropout)
        ctx.cfg = (q_off, k_off, v_off, H, heads, scale, pq is pkv, dropout)
        ctx.save_for_backward(pq, pkv, o, lse)
        return o

    @staticmethod
    def backward(ctx, do):
        pq, pkv, o, lse = ctx.saved_tensors
        q_off, k_off, v_off, H, heads, scale, same, dropout = ctx.cfg
        dpq = torch.zeros_like(pq)
        dpkv = dpq if same else torch.zeros_like(pkv)
        K.attn_bwd(_c(do), pq[:, :, q_off:q_off + H], pkv[:, :, k_off:k_off + H], pkv[:, :, v_off:v_off + H], o, lse, heads, scale,
                   dpq[:, :, q_off:q_off + H], dpkv[:, :, k_off:k_off + H], dpkv[:, :, v_off:v_off + H], dropout=dropout)
        return dpq, (None if same else dpkv), None, None, None, None, None, None, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta):
        x = _c(x)
        y, mean, rstd = K.layernorm_fwd(x, gamma.detach(), beta.detach(), LN_EPS)
        ctx.save_for_backward(x, mean, rstd, gamma.detach())
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma = ctx.saved_tensors
        dg = torch.zeros_like(gamma)
        db = torch.zeros_like(gamma)
        dx = K.layernorm_bwd(_c(dy), x, mean, rstd, gamma, dg, db)
        return dx, dg, db


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        return K.add(_c(a), _c(b), None if c is None else _c(c))

    @staticmethod
    def backward(ctx, dy):
        return dy, dy, (dy if ctx.needs_input_grad[2] else None)


class GatFn(torch.autograd.Function):
    """relu(GATConv core) on dense 3-node graphs given the `lin` projection xp [B,3,heads*C]."""

    @staticmethod
    def forward(ctx, xp, att_src, att_dst, bias, heads, slope, dropout=None):
        xp = _c(xp)
        a_s, a_d = att_src.detach().reshape(-1).contiguous(), att_dst.detach().reshape(-1).contiguous()
        out, alpha = K.gat_fwd(xp, a_s, a_d, bias.detach(), heads, slope, dropout=dropout)
        ctx.cfg = (heads, slope, att_src.shape, dropout)
        ctx.save_for_backward(xp, out, alpha, a_s, a_d)
        return out

    @staticmethod
    def backward(ctx, dout):
        xp, out, alpha, a_s, a_d = ctx.saved_tensors
        heads, slope, ashape, dropout = ctx.cfg
        dxp, ds, dd, dbias = K.gat_bwd(_c(dout), out, xp, alpha, a_s, a_d, heads, slope, dropout=dropout)
        return dxp, ds.reshape(ashape), dd.reshape(ashape), dbias, None, None, None


class Tok3AttnFn(torch.autograd.Function):
    """Self-attention over the 3 modality tokens; returns (ctx [B,3,H], head-averaged weights [B,3,3] fp32)."""

    @staticmethod
    def forward(ctx, qkv, heads, scale, dropout=None):
        qkv = _c(qkv)
        out, probs, avgw = K.tok3_attn_fwd(qkv, heads, scale, dropout=dropout)
        ctx.cfg = (heads, scale, dropout)
        ctx.save_for_backward(qkv, probs)
        return out, avgw

    @staticmethod
    def backward(ctx, dctx, davgw):
        qkv, probs = ctx.saved_tensors
        heads, scale, dropout = ctx.cfg
        davgw = None if davgw is None else _c(davgw.float())
        return K.tok3_attn_bwd(_c(dctx), davgw, qkv, probs, heads, scale, dropout=dropout), None, None, None


class GateMixFn(torch.autograd.Function):
    """gate = softmax(logits) (fp32), mixed = sum_m att[:,m]*gate[:,m]."""

    @staticmethod
    def forward(ctx, att, logits):
        att, logits = _c(att), _c(logits)
        gate, mixed = K.gate_mix_fwd(att, logits)
        ctx.save_for_backward(att, gate, logits)
        return mixed, gate

    @staticmethod
    def backward(ctx, dmixed, dgate):
        att, gate, logits = ctx.saved_tensors
        dgate = None if dgate is None else _c(dgate.float())
        datt, dlogits = K.gate_mix_bwd(_c(dmixed), dgate, att, gate, logits)
        return datt, dlogits


class LateCombineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lt, la, lv, w3):
        lt, la, lv = _c(lt), _c(la), _c(lv)
        fused, wsoft = K.late_combine_fwd(lt, la, lv, w3.detach())
        ctx.save_for_backward(lt, la, lv, wsoft)
        return fused, wsoft

    @staticmethod
    def backward(ctx, dfused, dwsoft):
        lt, la, lv, wsoft = ctx.saved_tensors
        dl, dw3 = K.late_combine_bwd(_c(dfused), lt, la, lv, wsoft, None if dwsoft is None else _c(dwsoft.float()))
        return dl[0], dl[1], dl[2], dw3


class RowMaskFn(torch.autograd.Function):
    """x[b, ...] * mask[b, col] for [B,L,H] sequences (modality dropout on MulT inputs)."""

    @staticmethod
    def forward(ctx, x, mask, col):
        ctx.col = col
        ctx.save_for_backward(mask)
        return K.rowmask_apply_(x.clone(), mask, col)

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        return K.rowmask_apply_(dy.clone(), mask, ctx.col), None, None


# ------------------------------------------------------------------------------------------------
# InfoNCE over the global batch
# ------------------------------------------------------------------------------------------------
def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _all_gather_rows(x: Tensor, world: int, group) -> Tensor:
    """[n, ...] -> [world*n, ...] (rank-major).  The single forward collective of the path."""
    if world == 1:
        return x
    out = torch.empty((world * x.size(0),) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


PAIRS = ((0, 1, "text_audio"), (0, 2, "text_video"), (1, 2, "audio_video"))


class InfoNCE3Fn(torch.autograd.Function):
    """The three pairwise contrastive losses of ContrastiveFusion (fusion_layers.py:345-347, 361-375) with
    negatives over the GLOBAL batch.  Forward: one all-gather of the packed embeddings; each rank computes
    the row-LSE of its rows and the column-LSE of its columns (the mirrored block), so the per-rank partial
    loss needs no further exchange (one all-reduce of 3 scalars makes every rank hold the global loss).
    Backward: one all-gather of the LSE vectors; dz_local is then exact and entirely local."""

    @staticmethod
    def forward(ctx, zt, za, zv, temperature, group):
        world, rank = _world(group)
        z = torch.stack([_c(zt), _c(za), _c(zv)])                        # [3,Bl,D] packed send buffer
        Bl, D = z.shape[1], z.shape[2]
        zg = z if world == 1 else _all_gather_rows(z.transpose(0, 1).contiguous(), world, group).transpose(0, 1).contiguous()
        Bg, off, inv_tau = Bl * world, rank * Bl, 1.0 / temperature
        lses, losses = [], []
        for (i, j, _) in PAIRS:
            lse_r, diag = K.infonce_lse(z[i], zg[j], off, inv_tau)          # rows of S owned by this rank
            lse_c, _ = K.infonce_lse(z[j], zg[i], off, inv_tau, want_diag=False)   # columns of S owned by this rank
            lses += [lse_r, lse_c]
            losses.append(0.5 * ((lse_r - diag).sum() + (lse_c - diag).sum()) / Bg)
        loss = torch.stack(losses)
        if world > 1:
            dist.all_reduce(loss, group=group)
        ctx.cfg = (world, off, inv_tau, Bg, group)
        ctx.save_for_backward(z, zg, torch.stack(lses))
        return loss[0], loss[1], loss[2]

    @staticmethod
    def backward(ctx, g0, g1, g2):
        z, zg, lses = ctx.saved_tensors
        world, off, inv_tau, Bg, group = ctx.cfg
        lses_g = lses if world == 1 else _all_gather_rows(lses.transpose(0, 1).contiguous(), world, group).transpose(0, 1).contiguous()
        dz = torch.zeros(z.shape, device=z.device, dtype=torch.float32)
        coef = inv_tau / (2.0 * Bg)
        for p, ((i, j, _), g) in enumerate(zip(PAIRS, (g0, g1, g2))):
            if g is None:
                continue
            g = g.float().reshape(1).contiguous()
            K.infonce_grad(z[i], zg[j], lses[2 * p], lses_g[2 * p + 1], coef, g, dz[i], True, off, inv_tau)
            K.infonce_grad(z[j], zg[i], lses[2 * p + 1], lses_g[2 * p], coef, g, dz[j], True, off, inv_tau)
        if z.dtype == torch.bfloat16:
            dz = K.cast_to_bf16(dz)
        return dz[0], dz[1], dz[2], None, None


def global_batch_scale(group=None) -> float:
    """1 / world_size: the factor that turns a LOCAL-mean loss (the reference trainer's CrossEntropy / MSE terms,
    training/advanced_trainer.py:136-150, or `SmoothedCrossEntropy`) into this rank's share of the GLOBAL-mean loss.
    Contract of `allreduce_gradients` / `GradBucket.all_reduce`: gradients are SUMMED over ranks, so every loss term must be
    normalised by the global batch before `backward()`.  InfoNCE3Fn already is (its negatives and its 1/B span the global
    batch); multiply every other term by this factor, or pass `scale=` to the all-reduce when NO InfoNCE term is present."""
    return 1.0 / _world(group)[0]


class GradBucket:
    """All parameter gradients of a module in ONE flat fp32 buffer, `p.grad` being views of it (the layout DDP calls
    gradient_as_bucket_view): autograd accumulates into the views, `zero()` is one fill kernel, `all_reduce()` is one NCCL call on
    the buffer in place -- no torch.cat, no per-tensor copy back -- and a fused optimizer sees one contiguous range.  Every
    parameter has a slot whether or not it received a gradient this step, so the buffer layout is identical on every rank."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise B200FusionError("GradBucket: no parameter requires a gradient")
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64            # 256-byte aligned slots
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, self.params)]
        self.attach()

    def attach(self) -> None:
        """(re)install the views as .grad -- after `optimizer.zero_grad(set_to_none=True)` or `p.grad = None`"""
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v

    def zero(self) -> None:
        self.attach()
        self.flat.zero_()

    def all_reduce(self, group=None, scale: Optional[float] = None, async_op: bool = False):
        world, _ = _world(group)
        if scale is not None and scale != 1.0:
            self.flat.mul_(scale)
        if world == 1:
            return None
        return dist.all_reduce(self.flat, group=group, async_op=async_op)


def allreduce_gradients(params, group=None, bucket_bytes: int = 256 << 20, scale: Optional[float] = None) -> None:
    """One bucketed SUM all-reduce of the parameter gradients per step, for gradients that live in separate
    tensors (see GradBucket for the copy-free layout).  SUM, not mean: every loss term must already be normalised by the GLOBAL
    batch (see `global_batch_scale`); `scale` multiplies the reduced gradients (1/world turns the sum into a mean when all loss
    terms are local means).  A parameter without a gradient on this rank contributes zeros, so all ranks reduce buffers of
    one layout (ranks with different `None` sets would otherwise mismatch or hang)."""
    world, _ = _world(group)
    params = [p for p in params if p.requires_grad]
    if world == 1:
        if scale is not None and scale != 1.0:
            for p in params:
                if p.grad is not None:
                    p.grad.mul_(scale)
        return
    bucket, size = [], 0

    def flush():
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p, dtype=torch.float32)).reshape(-1).float() for p in bucket])
        dist.all_reduce(flat, group=group)
        if scale is not None and scale != 1.0:
            flat.mul_(scale)
        o = 0
        for p in bucket:
            g = flat[o:o + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            o += p.numel()

    for p in params:
        bucket.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            flush()
            bucket, size = [], 0
    flush()


def mha_scale(H: int, heads: int) -> float:
    return 1.0 / math.sqrt(H // heads)
