"""CUDA heads (through the C ABI, fp32 parity mode) against the committed golden fixtures themselves -- the values the
EXECUTED reference produced (oracle/make_golden.py) -- without the oracle in between: outputs, losses, input
gradients and parameter gradients, norm-wise rtol 1e-5 (north_star fp32 tolerance).  Inputs and weights are
regenerated from the seeds stored in each fixture (the same seeded generators make_golden.py used)."""
import glob
import os

import pytest
import torch

from oracle import fusion_oracle as fo          # seeded input / weight generators only
from parity_util import CLS, FL, Cfg, rel

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))
RTOL = 1e-5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_cuda_head_matches_reference_fixture(path):
    rec = torch.load(path, weights_only=True)
    m = rec["meta"]
    cfg = Cfg(H=m["H"], heads=m["heads"], graph_hidden=m["graph_hidden"], graph_layers=m["graph_layers"])
    P = fo.init_params(m["kind"], H=m["H"], heads=m["heads"], graph_hidden=m["graph_hidden"], graph_layers=m["graph_layers"],
                       seed=m["param_seed"])
    feats = fo.synthetic_features(m["B"], m["lens"], H=m["H"], seed=m["feat_seed"])
    head = getattr(FL, CLS[m["kind"]])(cfg).cuda()
    head.load_state_dict(P, strict=True)
    head.train()
    xs = [f.cuda().float().requires_grad_(True) for f in feats]
    kw = {"compute_contrastive_loss": m["flag"]} if m["kind"] in ("contrastive", "hierarchical") else {}
    out = head(*xs, **kw)
    loss = fo.objective(out)
    loss.backward()
    torch.cuda.synchronize()

    assert abs(float(loss.detach()) - float(rec["loss"])) <= RTOL * max(1.0, abs(float(rec["loss"])))
    if isinstance(out, torch.Tensor):
        assert rel(out, rec["outputs"]["__tensor__"]) <= RTOL
    else:
        for k, ref in rec["outputs"].items():
            assert rel(out[k], ref) <= RTOL, k
        for k, ref in rec["losses"].items():
            assert abs(float(out["contrastive_losses"][k].detach()) - float(ref)) <= RTOL * max(1.0, abs(float(ref))), k
    top = max(float(g.norm()) for g in rec["input_grads"])
    for i, (x, g) in enumerate(zip(xs, rec["input_grads"])):
        err = float((x.grad.double().cpu() - g).norm()) / max(float(g.norm()), 1e-4 * top)
        assert err <= RTOL, (i, err)
    grads = {k: p.grad for k, p in head.named_parameters()}
    if m["full"]:
        top = max(float(g.norm()) for g in rec["param_grads"].values())
        for k, g in rec["param_grads"].items():
            err = float((grads[k].double().cpu() - g).norm()) / max(float(g.norm()), 1e-4 * top)
            assert err <= RTOL, (k, err)
    else:                                            # H=512 fixtures store (sum, norm, first 4) digests of each gradient
        for k, dg in rec["param_grads"].items():
            f = grads[k].double().cpu().flatten()
            assert abs(float(f.norm()) - float(dg[1])) <= 2 * RTOL * max(float(dg[1]), 1e-12), k
            assert float((f[:4] - dg[2:6][:f.numel()]).abs().max()) <= 2e-5 * max(float(dg[1]), 1e-12), k
