"""FusedAdamW + device-side gradient clipping against torch.optim.AdamW + torch.nn.utils.clip_grad_norm_ (the reference trainer's
optimizer step, training/advanced_trainer.py:91-94,174-180) on the same parameters and gradients, two param groups with different
learning rates, a OneCycleLR schedule, several steps; fp32, 1e-6."""
import copy
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("simple-multimodal_b200")


def _models():
    torch.manual_seed(0)
    shapes = [(512, 1536), (512,), (1024, 2560), (7, 256), (3,), (1, 4, 64), (2048, 512), (4097,)]
    a = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.1) for s in shapes]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    return a, b


@pytest.mark.parametrize("max_norm", [1.0, 1e9], ids=["clipping", "no-clipping"])
def test_fused_adamw_matches_torch(max_norm):
    pa, pb = _models()
    groups = lambda ps: [{"params": ps[:3], "lr": 1e-4}, {"params": ps[3:], "lr": 1e-3}]
    ref = torch.optim.AdamW(groups(pa), weight_decay=0.01)
    ours = pkg.FusedAdamW(groups(pb), weight_decay=0.01)
    sr = torch.optim.lr_scheduler.OneCycleLR(ref, max_lr=[1e-4, 1e-3], total_steps=20, pct_start=0.1, anneal_strategy="cos")
    so = torch.optim.lr_scheduler.OneCycleLR(ours, max_lr=[1e-4, 1e-3], total_steps=20, pct_start=0.1, anneal_strategy="cos")
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(6):
        for x, y in zip(pa, pb):
            grad = torch.randn(x.shape, device="cuda", generator=g) * (5.0 if it % 2 else 0.05)
            x.grad, y.grad = grad.clone(), grad.clone()
        if it == 3:                                   # a parameter without a gradient is skipped, like torch
            pa[4].grad = pb[4].grad = None
        n_ref = torch.nn.utils.clip_grad_norm_(pa, max_norm)
        n_ours = ours.clip_grad_norm_(max_norm)
        ref.step(); ours.step()
        sr.step(); so.step()
        assert abs(float(n_ours) - float(n_ref)) <= 1e-5 * float(n_ref)
        for x, y in zip(pa, pb):
            err = float((x - y).abs().max()) / max(float(x.abs().max()), 1e-12)
            assert err < 2e-6, (it, tuple(x.shape), err)
    assert ours.state[pb[0]]["step"] == 6 and ours.state[pb[4]]["step"] == 5


def test_fused_adamw_refuses_cpu_and_half():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.zeros(4)
    with pytest.raises(pkg.B200FusionError):
        pkg.FusedAdamW([p]).step()


def test_fused_adamw_resumes_a_torch_adamw_checkpoint_and_bumps_versions():
    """the reference trainer checkpoints torch.optim.AdamW's state_dict (training/advanced_trainer.py): its `step` entries are
    tensors; FusedAdamW must resume from it and continue exactly like torch.  Also: the kernel writes parameters through raw pointers,
    so it must bump their version counters (operand caches such as mult_engine._Weights key on them)."""
    pa, pb = _models()
    ref = torch.optim.AdamW(pa, lr=1e-3, weight_decay=0.01)
    g = torch.Generator(device="cuda").manual_seed(2)
    for _ in range(3):
        for x in pa:
            x.grad = torch.randn(x.shape, device="cuda", generator=g)
        ref.step()
    with torch.no_grad():
        for x, y in zip(pa, pb):
            y.copy_(x)
    ours = pkg.FusedAdamW(pb, lr=1e-3, weight_decay=0.01)
    # a checkpoint goes through torch.save / torch.load; load_state_dict alone would SHARE exp_avg / exp_avg_sq with `ref`
    # (`.to()` of a tensor already on the right device and dtype returns the tensor itself) and both optimizers would update them
    ours.load_state_dict(copy.deepcopy(ref.state_dict()))
    assert all(torch.is_tensor(st["step"]) for st in ours.state.values())        # what torch saved
    v0 = [y._version for y in pb]
    for _ in range(2):
        for x, y in zip(pa, pb):
            grad = torch.randn(x.shape, device="cuda", generator=g)
            x.grad, y.grad = grad.clone(), grad.clone()
        ref.step(); ours.step()
    for x, y in zip(pa, pb):
        assert float((x - y).abs().max()) / max(float(x.abs().max()), 1e-12) < 2e-6
    assert all(st["step"] == 5 for st in ours.state.values())
    assert all(y._version > v for y, v in zip(pb, v0))
