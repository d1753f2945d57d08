"""Row kernels through the C ABI against plain torch fp32 math: LayerNorm forward / backward (TMA-staged kernels of
csrc/rownorm_tma.cu and the register-staged ones behind b200f_debug_set(9, 1)) over ragged row counts (last tile partial, fewer rows
than one tile, fewer than the TMA threshold), widths with one and two 32-lane passes, with and without the residual post-adds /
residual gradient / upstream bias-gradient column sum; mean and weighted pooling over odd shapes."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("simple-multimodal_b200")
K = pkg.kernels


def _rel(x, ref):
    return float((x.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30))


@pytest.fixture(params=[0, 1], ids=["tma", "regs"])
def ln_family(request):
    pkg._lib.lib().b200f_debug_set(9, request.param)
    yield request.param
    pkg._lib.lib().b200f_debug_set(9, 0)


@pytest.mark.parametrize("dtype,H", [(torch.bfloat16, 512), (torch.bfloat16, 256), (torch.bfloat16, 64), (torch.float32, 256), (torch.float32, 64),
                                     (torch.float32, 512)], ids=["bf16-512", "bf16-256", "bf16-64", "fp32-256", "fp32-64", "fp32-512"])
@pytest.mark.parametrize("rows", [1, 63, 64, 65, 1000, 4097])
def test_layernorm_forward_backward(ln_family, dtype, H, rows):
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + H)
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g)
    x, p1, p2, dy, dres = (mk(rows, H).to(dtype) for _ in range(5))
    gamma, beta = torch.rand(H, device="cuda", generator=g) + 0.5, mk(H)
    tol = 2e-6 if dtype == torch.float32 else 6e-3
    xf = x.float()
    mu, var = xf.mean(1, keepdim=True), xf.var(1, unbiased=False, keepdim=True)
    xhat = (xf - mu) * torch.rsqrt(var + 1e-5)
    for posts in (0, 1, 2):
        y, mean, rstd = K.layernorm_fwd(x, gamma, beta, 1e-5, *([p1, p2][:posts]))
        ref = xhat * gamma + beta + (p1.float() if posts >= 1 else 0) + (p2.float() if posts >= 2 else 0)
        assert _rel(y, ref) < tol, posts
        assert float((mean - mu.squeeze(1)).abs().max()) < 1e-5
        assert _rel(rstd, torch.rsqrt(var + 1e-5).squeeze(1)) < 1e-5
    _, mean, rstd = K.layernorm_fwd(x, gamma, beta, 1e-5)
    dyf = dy.float()
    gg = dyf * gamma
    dx_ref = rstd[:, None] * (gg - gg.mean(1, keepdim=True) - xhat * (gg * xhat).mean(1, keepdim=True))
    for with_res in (False, True):
        dg, db, dsum = (torch.zeros(H, device="cuda") for _ in range(3))
        dx = K.layernorm_bwd(dy, x, mean, rstd, gamma, dg, db, dres=dres if with_res else None, dxsum=dsum if with_res else None)
        ref = dx_ref + (dres.float() if with_res else 0)
        assert _rel(dx, ref) < tol, with_res
        assert _rel(dg, (dyf * xhat).sum(0)) < max(tol, 1e-5)
        assert _rel(db, dyf.sum(0)) < max(tol, 1e-5)
        if with_res:
            assert _rel(dsum, dx.float().sum(0)) < 1e-5                  # column sum of the STORED dx


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("B,L,H", [(1, 1, 64), (3, 30, 512), (5, 7, 48), (2, 513, 512), (4, 100, 1024)])
def test_mean_and_weighted_pool(dtype, B, L, H):
    g = torch.Generator(device="cuda").manual_seed(B * 100 + L)
    x = torch.randn(B, L, H, device="cuda", generator=g).to(dtype)
    tol = 2e-6 if dtype == torch.float32 else 6e-3
    assert _rel(K.meanpool_fwd(x), x.float().mean(1)) < tol
    w = torch.rand(B, L, device="cuda", generator=g)
    w[0, L // 2:] = 0
    w = (w / w.sum(1, keepdim=True).clamp(min=1e-9)).contiguous()
    assert _rel(K.weighted_pool_fwd(x, w), torch.einsum("bl,blh->bh", w, x.float())) < tol
    dy = torch.randn(B, H, device="cuda", generator=g).to(dtype)
    assert _rel(K.weighted_pool_bwd(dy, w), w[:, :, None] * dy.float()[:, None, :]) < tol
    assert _rel(K.meanpool_bwd(dy, L), (dy.float() / L)[:, None, :].expand(B, L, H)) < tol
