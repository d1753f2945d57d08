"""Parity of the CUDA fusion heads (through the C ABI) against the CPU oracle on identical seeded features and
weights: outputs, losses, input gradients and parameter gradients.
fp32: 1e-5 norm-wise.  bf16: outputs 2e-2, loss 1e-3, gradients 2e-2 / reference-autocast floor (parity_util.compare)."""
import pytest
import torch

from parity_cases import CASES
from parity_util import Cfg, compare
from oracle import fusion_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("case_id", list(CASES))
def test_head_parity(case_id, dtype):
    c = CASES[case_id]
    if c["fp32_only"] and dtype != torch.float32:
        pytest.skip("fp32-only case")
    mask = fo.modality_keep_mask(c["B"], 0.4, torch.Generator().manual_seed(c["mask_seed"])) if c["mask_seed"] else None
    compare(c["kind"], Cfg(**c["cfg"]), c["B"], c["lens"], dtype, flag=c["flag"], mask=mask, chunk=c["chunk"], case_id=case_id)
