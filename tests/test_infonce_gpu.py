"""Fused tcgen05 InfoNCE kernels (csrc/infonce_tc.cu) through the C ABI: row log-sum-exp + positive logit of a [Bl x Bg] block of
the similarity matrix and the gradient block dx += W y, against explicit fp64 torch math (the reference's contrastive_loss
arithmetic, models/fusion_layers.py:366-375, restricted to a row block as the data-parallel ranks use it) and against the
GEMM + row-kernel route (b200f_debug_set(6, 1)).  Ragged sizes, a diagonal offset (rank > 0), several key splits."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("simple-multimodal_b200")
K = pkg.kernels

CASES = [(2, 2, 256, 0), (64, 64, 256, 0), (100, 1000, 256, 37), (96, 192, 256, 96), (1024, 4096, 256, 3072), (300, 300, 128, 0),
         (4096, 4096, 256, 0), (130, 8192, 64, 5000)]


def _unit(n, d, seed, dtype):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n, d, device="cuda", generator=g), dim=-1).to(dtype)


@pytest.mark.parametrize("Bl,Bg,D,off", CASES)
def test_fused_infonce_block(Bl, Bg, D, off):
    inv_tau = 1 / 0.07
    y = _unit(Bg, D, 1, torch.bfloat16)
    x = (y[off:off + Bl].float() + 0.3 * _unit(Bl, D, 2, torch.float32)).to(torch.bfloat16)     # positives are similar, not identical
    lse, diag = K.infonce_lse(x, y, off, inv_tau)
    S = inv_tau * x.double() @ y.double().T
    idx = torch.arange(Bl, device="cuda")
    assert float((lse.double() - torch.logsumexp(S, 1)).abs().max()) < 2e-3
    assert float((diag.double() - S[idx, off + idx]).abs().max()) < 1e-3
    lse_nd, none = K.infonce_lse(x, y, off, inv_tau, want_diag=False)
    assert none is None and torch.equal(lse_nd, lse)
    # gradient block: W = c g (exp(S - lse_x) + exp(S - lse_y) - 2 I), dx += W y
    lse_y = torch.logsumexp(inv_tau * y.double() @ y.double().T, 1).float() if Bg <= 4096 else torch.randn(Bg, device="cuda") + 8.0
    coef, g = 0.37, torch.tensor([1.7], device="cuda")
    dx = torch.full((Bl, D), 0.5, device="cuda")
    K.infonce_grad(x, y, lse, lse_y, coef, g, dx, True, off, inv_tau)
    W = torch.exp(S - lse.double()[:, None]) + torch.exp(S - lse_y.double()[None, :])
    W[idx, off + idx] -= 2.0
    want = 0.5 + (W * (coef * 1.7)) @ y.double()
    err = float((dx.double() - want).norm() / (want - 0.5).norm())
    assert err < 2e-2, err
    # the GEMM + row-kernel route agrees (same bf16 operands, S through HBM)
    pkg._lib.lib().b200f_debug_set(6, 1)
    try:
        lse2, diag2 = K.infonce_lse(x, y, off, inv_tau)
        dx2 = torch.full((Bl, D), 0.5, device="cuda")
        K.infonce_grad(x, y, lse, lse_y, coef, g, dx2, True, off, inv_tau)
    finally:
        pkg._lib.lib().b200f_debug_set(6, 0)
    torch.cuda.synchronize()
    assert float((lse2 - lse).abs().max()) < 2e-3 and float((diag2 - diag).abs().max()) < 1e-3
    assert float((dx2 - dx).norm() / (dx2 - 0.5).norm()) < 2e-2


def test_fused_infonce_overwrite_mode():
    x, y = _unit(256, 256, 3, torch.bfloat16), _unit(512, 256, 4, torch.bfloat16)
    lse, _ = K.infonce_lse(x, y, 0, 10.0)
    lse_y = torch.zeros(512, device="cuda") + 6.0
    a = torch.full((256, 256), 9.0, device="cuda")
    b = torch.zeros(256, 256, device="cuda")
    K.infonce_grad(x, y, lse, lse_y, 1.0, None, a, False, 0, 10.0)       # accumulate = False: dx is overwritten
    K.infonce_grad(x, y, lse, lse_y, 1.0, None, b, True, 0, 10.0)
    torch.cuda.synchronize()
    assert float((a - b).abs().max()) < 1e-5 * float(b.abs().max()) + 1e-7
