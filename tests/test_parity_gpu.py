"""Parity of the CUDA fusion heads (through the C ABI) against the CPU oracle: outputs, losses, input
gradients and parameter gradients, fp32 (rtol 1e-5 norm-wise) and bf16 (2e-2, loss 1e-3)."""
import pytest
import torch

from parity_util import Cfg, compare
from oracle import fusion_oracle as fo

pytestmark = pytest.mark.gpu
DTYPES = [torch.float32, torch.bfloat16]
ids = lambda d: "fp32" if d == torch.float32 else "bf16"


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
@pytest.mark.parametrize("B", [1, 16, 257])
def test_early(dtype, B):
    compare("early", Cfg(), B, (None, None, None), dtype)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
def test_late(dtype):
    compare("late", Cfg(), 33, (None, None, None), dtype)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
@pytest.mark.parametrize("B", [2, 130])
def test_graph(dtype, B):
    compare("graph", Cfg(), B, (None, None, None), dtype)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
def test_graph_single_layer_256(dtype):
    compare("graph", Cfg(graph_hidden=256, graph_layers=1), 9, (None, None, None), dtype)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
@pytest.mark.parametrize("B", [2, 64, 1000])
def test_contrastive(dtype, B):
    compare("contrastive", Cfg(), B, (None, None, None), dtype, flag=True)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
def test_contrastive_without_loss(dtype):
    compare("contrastive", Cfg(), 8, (None, None, None), dtype, flag=False)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
@pytest.mark.parametrize("B", [1, 70])
def test_adaptive(dtype, B):
    compare("adaptive", Cfg(), B, (None, None, None), dtype)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
def test_mult_2d(dtype):
    compare("mult", Cfg(), 6, (None, None, None), dtype)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
@pytest.mark.parametrize("lens", [(1, 1, 1), (64, 64, 30), (33, 65, 7)])
def test_mult_3d(dtype, lens):
    compare("mult", Cfg(), 3, lens, dtype)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
def test_mult_chunked_recompute_matches(dtype):
    compare("mult", Cfg(), 5, (40, 24, 30), dtype, chunk=2)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
def test_hierarchical_2d(dtype):
    compare("hierarchical", Cfg(), 12, (None, None, None), dtype, flag=True)


@pytest.mark.parametrize("dtype", DTYPES, ids=ids)
def test_hierarchical_3d_with_mask(dtype):
    mask = fo.modality_keep_mask(6, 0.4, torch.Generator().manual_seed(4321))
    compare("hierarchical", Cfg(), 6, (48, 32, 30), dtype, flag=True, mask=mask, chunk=4)
