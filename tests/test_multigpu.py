"""Two ranks over NCCL (skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`):
ContrastiveFusion with global-batch negatives.  Each rank runs its half of the batch through the CUDA head; losses and
gradients must equal the fp64 oracle on the concatenated batch (the reference's contrastive_loss on B_global rows,
models/fusion_layers.py:361-375), and after `allreduce_gradients` every rank holds the full-batch parameter gradient."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, port, Bl, dtype_name, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    try:
        from oracle import fusion_oracle as fo          # seeded generators (checker side)
        from parity_util import Cfg
        pkg = importlib.import_module("simple-multimodal_b200")
        dtype = getattr(torch, dtype_name)
        P = fo.init_params("contrastive", H=512, heads=8, seed=9)
        feats = fo.synthetic_features(WORLD * Bl, (None, None, None), H=512, seed=31)
        if dtype == torch.bfloat16:
            P = {k: v.to(dtype).float() for k, v in P.items()}
            feats = [f.to(dtype).float() for f in feats]
        head = pkg.fusion_layers.ContrastiveFusion(Cfg()).cuda()
        head.load_state_dict(P, strict=True)
        head.train()
        xs = [f[rank * Bl:(rank + 1) * Bl].cuda().to(dtype).requires_grad_(True) for f in feats]
        out = head(*xs, compute_contrastive_loss=True)
        loss = (out["fused_features"].float() ** 2).sum() / (WORLD * Bl * 512) + 0.1 * sum(out["contrastive_losses"].values())
        loss.backward()
        pkg.allreduce_gradients(list(head.parameters()))
        torch.cuda.synchronize()
        ret[rank] = dict(losses={k: float(v.detach()) for k, v in out["contrastive_losses"].items()},
                         dx=[x.grad.float().cpu() for x in xs],
                         pg={k: p.grad.float().cpu() for k, p in head.named_parameters()})
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < WORLD, reason="needs 2 GPUs")
@pytest.mark.parametrize("dtype_name,Bl", [("float32", 96), ("bfloat16", 1024)])
def test_contrastive_two_ranks_nccl(dtype_name, Bl):
    from oracle import fusion_oracle as fo
    from parity_util import rel
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(port, Bl, dtype_name, ret), nprocs=WORLD, join=True)
        res = {r: ret[r] for r in range(WORLD)}
    bf16 = dtype_name == "bfloat16"
    P = fo.init_params("contrastive", H=512, heads=8, seed=9)
    feats = fo.synthetic_features(WORLD * Bl, (None, None, None), H=512, seed=31)
    if bf16:
        P = {k: v.to(torch.bfloat16).float() for k, v in P.items()}
        feats = [f.to(torch.bfloat16).float() for f in feats]
    P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
    x64 = [f.double().requires_grad_(True) for f in feats]
    ref = fo.contrastive_fusion(*x64, P64, temperature=0.07, compute_contrastive_loss=True)
    ((ref["fused_features"] ** 2).sum() / (WORLD * Bl * 512) + 0.1 * sum(ref["contrastive_losses"].values())).backward()
    tol_l, tol_g = (1e-3, 5e-2) if bf16 else (1e-5, 1e-5)
    for r in range(WORLD):
        for k, v in ref["contrastive_losses"].items():
            assert abs(res[r]["losses"][k] - float(v.detach())) <= tol_l, (r, k)
        for i in range(3):
            assert rel(res[r]["dx"][i], x64[i].grad[r * Bl:(r + 1) * Bl]) <= tol_g, (r, i)
        for k, g in P64.items():
            assert rel(res[r]["pg"][k], g.grad) <= tol_g, (r, k)


def _worker_hier(rank, port, Bl, dtype_name, ret):
    """HierarchicalFusion on [B,L,H] sequences with a modality mask, global-batch InfoNCE negatives, gradients in a GradBucket"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    try:
        from oracle import fusion_oracle as fo
        from parity_util import Cfg
        pkg = importlib.import_module("simple-multimodal_b200")
        dtype = getattr(torch, dtype_name)
        lens = (24, 16, 30)
        P = fo.init_params("hierarchical", H=512, heads=8, seed=5)
        feats = fo.synthetic_features(WORLD * Bl, lens, H=512, seed=77)
        mask = fo.modality_keep_mask(WORLD * Bl, 0.3, torch.Generator().manual_seed(3))
        if dtype == torch.bfloat16:
            P = {k: v.to(dtype).float() for k, v in P.items()}
            feats = [f.to(dtype).float() for f in feats]
        head = pkg.fusion_layers.HierarchicalFusion(Cfg()).cuda()
        head.load_state_dict(P, strict=True)
        head.train()
        head.mult_fusion.chunk_size, head.mult_fusion.graph_min_tokens = 3, 0          # chunk graphs + a ragged last chunk on every rank
        bucket = pkg.GradBucket(list(head.parameters()))
        bucket.zero()
        xs = [f[rank * Bl:(rank + 1) * Bl].cuda().to(dtype).requires_grad_(True) for f in feats]
        out = head(*xs, compute_contrastive_loss=True, mask=mask[rank * Bl:(rank + 1) * Bl].cuda())
        loss = (out["fused_features"].float() ** 2).sum() / (WORLD * Bl * 512) + 0.1 * sum(out["contrastive_losses"].values())
        loss.backward()
        bucket.all_reduce()
        torch.cuda.synchronize()
        ret[rank] = dict(losses={k: float(v.detach()) for k, v in out["contrastive_losses"].items()},
                         fused=out["fused_features"].detach().float().cpu(), dx=[x.grad.detach().float().cpu() for x in xs],
                         pg={k: p.grad.detach().float().cpu() for k, p in head.named_parameters()},
                         engine=head.mult_fusion._engine is not None)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < WORLD, reason="needs 2 GPUs")
@pytest.mark.parametrize("dtype_name", ["float32", "bfloat16"])
def test_hierarchical_two_ranks_nccl(dtype_name):
    from oracle import fusion_oracle as fo
    from parity_util import GRAD_CAP, rel
    Bl = 8
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_hier, args=(port, Bl, dtype_name, ret), nprocs=WORLD, join=True)
        res = {r: ret[r] for r in range(WORLD)}
    bf16 = dtype_name == "bfloat16"
    lens = (24, 16, 30)
    P = fo.init_params("hierarchical", H=512, heads=8, seed=5)
    feats = fo.synthetic_features(WORLD * Bl, lens, H=512, seed=77)
    mask = fo.modality_keep_mask(WORLD * Bl, 0.3, torch.Generator().manual_seed(3))
    if bf16:
        P = {k: v.to(torch.bfloat16).float() for k, v in P.items()}
        feats = [f.to(torch.bfloat16).float() for f in feats]
    P64 = {k: v.double().requires_grad_(True) for k, v in P.items()}
    x64 = [f.double().requires_grad_(True) for f in feats]
    ref = fo.hierarchical_fusion(*fo.apply_modality_mask(*x64, mask.double()), P64, heads=8, compute_contrastive_loss=True, temperature=0.07,
                                 graph_layers=3)
    ((ref["fused_features"] ** 2).sum() / (WORLD * Bl * 512) + 0.1 * sum(ref["contrastive_losses"].values())).backward()
    tol_o, tol_l, tol_g = (2e-2, 1e-3, GRAD_CAP) if bf16 else (1e-5, 1e-5, 1e-5)
    top = max(float(g.grad.norm()) for g in P64.values())
    for r in range(WORLD):
        assert res[r]["engine"] == bf16                                   # bf16 ran through the chunk graphs, fp32 eagerly
        assert rel(res[r]["fused"], ref["fused_features"][r * Bl:(r + 1) * Bl]) <= tol_o, r
        for k, v in ref["contrastive_losses"].items():
            assert abs(res[r]["losses"][k] - float(v.detach())) <= tol_l, (r, k)
        for i in range(3):
            want = x64[i].grad[r * Bl:(r + 1) * Bl]
            assert rel(res[r]["dx"][i], want) <= tol_g, (r, i, rel(res[r]["dx"][i], want))
        for k, g in P64.items():                                          # after the all-reduce: the full-batch gradient on every rank
            err = float((res[r]["pg"][k].double() - g.grad).norm()) / max(float(g.grad.norm()), 1e-4 * top)
            assert err <= tol_g, (r, k, err)
