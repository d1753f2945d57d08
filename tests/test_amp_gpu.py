"""The reference trainer wraps the forward in torch.cuda.amp.autocast (fp16 by default) and scales the loss with GradScaler
(training/advanced_trainer.py:57,131,171-176).  The drop-in heads are opaque to autocast, so they must (i) pick their bf16 path when
autocast is on, whatever the input dtype (fp32 encoder features, fp16 autocast outputs), (ii) hand back finite fp32 parameter
gradients under a 2^16 loss scale that unscale to the unscaled run's gradients, (iii) keep losses in fp32."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("simple-multimodal_b200")
FL = pkg.fusion_layers


class Cfg:
    fusion_hidden_size, fusion_num_heads, fusion_dropout = 512, 8, 0.0
    num_emotions, graph_hidden_size, graph_num_layers, graph_dropout = 7, 512, 3, 0.0
    contrastive_temperature = 0.07


def _loss(out, clf, ce, target):
    logits = clf(out["fused_features"])
    return ce(logits, target) + 0.1 * sum(out["contrastive_losses"].values())


@pytest.mark.parametrize("in_dtype", [torch.float32, torch.float16], ids=["fp32-in", "fp16-in"])
def test_autocast_and_gradscaler(in_dtype):
    torch.manual_seed(0)
    head = FL.HierarchicalFusion(Cfg).cuda().train()
    clf = pkg.EmotionClassifier(Cfg).cuda().train()
    ce = pkg.SmoothedCrossEntropy(0.1)
    B = 16
    xs = [torch.randn(B, 512, device="cuda").to(in_dtype) for _ in range(3)]
    target = torch.randint(0, 7, (B,), device="cuda")
    params = list(head.parameters()) + list(clf.parameters())

    def run(scale):
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            out = head(*xs, compute_contrastive_loss=True)
            loss = _loss(out, clf, ce, target)
        assert out["fused_features"].dtype == torch.bfloat16          # the tcgen05 path ran
        assert loss.dtype == torch.float32 and all(v.dtype == torch.float32 for v in out["contrastive_losses"].values())
        (loss * scale).backward()
        torch.cuda.synchronize()
        return float(loss.detach()), [None if p.grad is None else p.grad.clone() for p in params]

    l1, g1 = run(1.0)
    l2, g2 = run(65536.0)                                             # GradScaler's initial scale
    assert l1 == l2
    n_checked = 0
    for a, b in zip(g1, g2):
        if a is None:
            assert b is None
            continue
        assert a.dtype == torch.float32 and torch.isfinite(b).all()
        na = float(a.norm())
        if na > 1e-8:
            assert float((b / 65536.0 - a).norm()) / na < 2e-2          # bf16 intermediates round differently at another scale
            n_checked += 1
    assert n_checked > 50
    # GradScaler itself drives the same tensors without complaints (unscale_, inf checks, step)
    opt = torch.optim.AdamW(params, lr=1e-4)
    scaler = torch.amp.GradScaler("cuda")
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.float16):
        loss = _loss(head(*xs, compute_contrastive_loss=True), clf, ce, target)
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    torch.nn.utils.clip_grad_norm_(params, 1.0)                        # advanced_trainer.py:174-176
    scaler.step(opt)
    scaler.update()
    assert float(scaler.get_scale()) == 65536.0                        # no inf / nan was found
