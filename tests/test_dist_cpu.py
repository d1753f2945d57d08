"""N>1 host logic of the path on CPU: world_size-2 `gloo` processes run the data-parallel plumbing of
`ops.InfoNCE3Fn` (packed all-gather, mirrored row/column blocks, LSE all-gather, global-batch normalisation) and
`ops.allreduce_gradients` (bucketed SUM all-reduce).  The two CUDA entry points InfoNCE3Fn calls are replaced, IN THE
TEST ONLY, by a few lines of torch that state what `b200f_infonce_lse` / `b200f_infonce_grad` compute on one block, so
what is checked here is the sharding/exchange logic -- the kernels themselves are checked on the GPU
(tests/test_parity_gpu.py).  The expected values come from the oracle's `info_nce` on the concatenated global batch
(the reference's contrastive_loss, models/fusion_layers.py:361-375)."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _block_lse(x, y, diag_off, inv_tau, want_diag=True):
    """What b200f_infonce_lse returns for the block S = inv_tau * x y^T (rows of x against all rows of y)."""
    S = inv_tau * x.double() @ y.double().T
    lse = torch.logsumexp(S, dim=1)
    idx = torch.arange(x.size(0))
    return lse.to(torch.float32 if x.dtype != torch.float64 else torch.float64), (S[idx, diag_off + idx].to(lse.dtype) if want_diag else None)


def _block_grad(x, y, lse_x, lse_y, coef, gscale, dx, accumulate, diag_off, inv_tau):
    """What b200f_infonce_grad accumulates: dx += c * (exp(S - lse_x[i]) + exp(S - lse_y[j]) - 2*[j == off+i]) y."""
    S = inv_tau * x.double() @ y.double().T
    W = torch.exp(S - lse_x.double()[:, None]) + torch.exp(S - lse_y.double()[None, :])
    idx = torch.arange(x.size(0))
    W[idx, diag_off + idx] -= 2.0
    g = (W * (coef * float(gscale))) @ y.double()
    if accumulate:
        dx += g.to(dx.dtype)
    else:
        dx.copy_(g)


def _worker(rank, port, Bl, D, tau, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        pkg = importlib.import_module("simple-multimodal_b200")
        ops, K = pkg.ops, pkg.kernels
        K.infonce_lse, K.infonce_grad = _block_lse, _block_grad          # stand-ins for the CUDA entry points (test only)
        g = torch.Generator().manual_seed(99)
        zg = torch.nn.functional.normalize(torch.randn(3, WORLD * Bl, D, generator=g, dtype=torch.float64), dim=-1)
        z = [zg[m, rank * Bl:(rank + 1) * Bl].clone().requires_grad_(True) for m in range(3)]
        losses = ops.InfoNCE3Fn.apply(z[0], z[1], z[2], tau, None)
        w = torch.tensor([1.0, 0.5, 2.0], dtype=torch.float64)
        (losses[0] * w[0] + losses[1] * w[1] + losses[2] * w[2]).backward()
        # bucketed gradient all-reduce: tiny buckets so several flushes happen
        params = [torch.nn.Parameter(torch.full((n,), float(rank + 1))) for n in (3, 1000, 17, 5)]
        for p in params:
            p.grad = torch.arange(p.numel(), dtype=torch.float32) * (rank + 1)
        params.append(torch.nn.Parameter(torch.zeros(2)))               # no grad on either rank: zero-filled slot
        lonely = torch.nn.Parameter(torch.zeros(6))                     # a gradient on rank 1 only: rank 0 contributes zeros, so
        if rank == 1:                                                   # both ranks reduce buffers of one layout (no mismatch / hang)
            lonely.grad = torch.full((6,), 7.0)
        params.append(lonely)
        params.append(torch.nn.Parameter(torch.zeros(3), requires_grad=False))      # frozen: not part of the exchange at all
        ops.allreduce_gradients(params, bucket_bytes=2048)
        # GradBucket: the gradients live in one flat buffer, all-reduced in place
        bp = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 130, 64)]
        bucket = ops.GradBucket(bp)
        bucket.zero()
        for i, q in enumerate(bp):
            (q * float((rank + 1) * (i + 1))).sum().backward()           # autograd accumulates into the views
        same_storage = all(q.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr() for q in bp)
        bucket.all_reduce(scale=ops.global_batch_scale())
        seeds = (ops.next_drop_seed(), ops._rank_seed())
        ret[rank] = dict(losses=[float(l.detach()) for l in losses], dz=[t.grad.clone() for t in z], zg=zg,
                         grads=[p.grad.clone() for p in params[:4]], none_grad=params[4].grad.clone(), lonely=lonely.grad.clone(),
                         frozen=params[-1].grad is None, bucket=[q.grad.clone() for q in bp], same_storage=same_storage, seeds=seeds)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("Bl,D", [(5, 16), (32, 64)])
def test_infonce_global_negatives_two_ranks_gloo(Bl, D):
    from oracle import fusion_oracle as fo
    tau = 0.07
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(port, Bl, D, tau, ret), nprocs=WORLD, join=True)
        res = {r: ret[r] for r in range(WORLD)}
    zg = res[0]["zg"].clone().requires_grad_(True)
    pairs = ((0, 1), (0, 2), (1, 2))
    ref = [fo.info_nce(zg[i], zg[j], tau) for i, j in pairs]
    (ref[0] * 1.0 + ref[1] * 0.5 + ref[2] * 2.0).backward()
    for r in range(WORLD):
        # every rank holds the GLOBAL loss (sum over ranks of the per-rank partial, normalised by the global batch)
        for got, want in zip(res[r]["losses"], ref):
            assert abs(got - float(want.detach())) < 1e-10
        # dz_local is the exact slice of the global-batch gradient -- cross-rank column terms included
        for m in range(3):
            want = zg.grad[m, r * Bl:(r + 1) * Bl]
            assert float((res[r]["dz"][m] - want).abs().max()) < 1e-6 * float(want.abs().max())   # dz accumulates in fp32 by design
        for n, gr in zip((3, 1000, 17, 5), res[r]["grads"]):
            assert torch.equal(gr, torch.arange(n, dtype=torch.float32) * 3.0)      # (1 + 2) * arange: SUM over ranks
        assert torch.equal(res[r]["none_grad"], torch.zeros(2)) and res[r]["frozen"]
        assert torch.equal(res[r]["lonely"], torch.full((6,), 7.0))
        assert res[r]["same_storage"]
        for i, gr in enumerate(res[r]["bucket"]):                                    # (1 + 2) * (i + 1) summed, times 1/world
            assert torch.equal(gr, torch.full_like(gr, 3.0 * (i + 1) / WORLD))
    assert res[0]["seeds"] != res[1]["seeds"]                                        # every rank draws its own dropout masks


def test_single_process_is_identity():
    """world == 1: no process group needed, gather is the identity, all-reduce is a no-op."""
    pkg = importlib.import_module("simple-multimodal_b200")
    x = torch.randn(4, 3)
    assert pkg.ops._all_gather_rows(x, 1, None) is x
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.ones(3)
    pkg.ops.allreduce_gradients([p])
    assert torch.equal(p.grad, torch.ones(3))
    pkg.ops.allreduce_gradients([p], scale=0.5)
    assert torch.equal(p.grad, torch.full((3,), 0.5))
    assert pkg.ops.global_batch_scale() == 1.0
    b = pkg.ops.GradBucket([p])                       # adopts the parameter: .grad becomes a view of the flat buffer
    b.zero()
    (p * 2.0).sum().backward()
    b.all_reduce()
    assert torch.equal(p.grad, torch.full((3,), 2.0)) and torch.equal(b.flat[:3], p.grad)
    p.grad = None
    b.zero()                                          # re-attaches after zero_grad(set_to_none=True)
    assert p.grad is b.views[0]
