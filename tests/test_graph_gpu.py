"""GraphedTrainStep: a whole training step captured in a CUDA graph gives the same loss and gradients as eager issue (dropout
off), and with dropout on each replay draws a new mask (device-side epoch) while staying a valid step."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("simple-multimodal_b200")
FL, ops = pkg.fusion_layers, pkg.ops


class Cfg:
    def __init__(self, p):
        self.fusion_hidden_size, self.fusion_num_heads, self.fusion_dropout = 512, 8, p
        self.num_emotions, self.graph_hidden_size, self.graph_num_layers, self.graph_dropout = 7, 512, 3, p
        self.contrastive_temperature = 0.07


def loss_fn(out):
    return (out["fused_features"].float() ** 2).mean()


def test_graphed_step_matches_eager():
    torch.manual_seed(0)
    head = FL.MultimodalTransformer(Cfg(0.0)).cuda()
    head.train()
    xs = [torch.randn(4, L, 512, device="cuda").to(torch.bfloat16).requires_grad_(True) for L in (96, 64, 30)]
    loss_fn(head(*xs)).backward()
    want_loss = float(loss_fn(head(*[x.detach() for x in xs])).detach())
    want = {k: p.grad.clone() for k, p in head.named_parameters()}
    want_dx = [x.grad.clone() for x in xs]
    step = pkg.GraphedTrainStep(head, xs, loss_fn)
    assert step.kernel_launches > 100
    for _ in range(3):                                                  # replays are idempotent: gradients are overwritten, not accumulated
        loss = step(*xs)
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - want_loss) < 1e-6
    for k, p in head.named_parameters():
        err = float((p.grad - want[k]).norm() / want[k].norm().clamp_min(1e-20))
        assert err < 1e-3 or float(want[k].norm()) < 1e-6, (k, err)      # fp32 atomics: equal up to summation order
    for g, w in zip(step.input_grads, want_dx):
        assert torch.equal(g, w)
    ys = [torch.randn_like(x) for x in xs]                              # new data goes through the static input buffers
    l2 = float(step(*ys).detach())
    assert abs(l2 - float(loss_fn(head(*ys)).detach())) < 1e-6 and abs(l2 - want_loss) > 1e-6


def test_graphed_step_draws_new_dropout_masks_per_replay():
    torch.manual_seed(0)
    ops.manual_seed(5)
    head = FL.MultimodalTransformer(Cfg(0.1)).cuda()
    head.train()
    xs = [torch.randn(4, L, 512, device="cuda").to(torch.bfloat16) for L in (96, 64, 30)]
    step = pkg.GraphedTrainStep(head, xs, loss_fn)
    losses = [float(step(*xs).detach()) for _ in range(4)]
    torch.cuda.synchronize()
    assert len(set(losses)) == 4                                        # same inputs, frozen seeds, advancing epoch: four different masks
    assert max(losses) / min(losses) < 1.5
    assert all(torch.isfinite(p.grad).all() for p in head.parameters())
    pkg.kernels.dropout_epoch(0)                                        # leave the epoch at its eager default for the other tests
