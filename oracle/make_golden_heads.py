"""Golden vectors of the heads downstream of fusion from the EXECUTED reference -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):  cd /tmp && python /root/repo/oracle/make_golden_heads.py
The reference's own `models.multimodal_model.EmotionClassifier` is imported (behind the torch_geometric stand-in of
ref_shim) and run in float64; the valence / arousal / uncertainty heads are the three nn.Linear of
MultimodalEmotionModel.__init__ (multimodal_model.py:55-60 -- the model itself cannot be constructed offline: its encoders call
from_pretrained) followed by F.softmax as in its forward (:160-164); the loss is the trainer's
nn.CrossEntropyLoss(label_smoothing=0.1) (training/advanced_trainer.py:53).  Writes tests/golden/heads/heads_h*.pt."""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import fusion_oracle as fo          # noqa: E402
from oracle import ref_shim                      # noqa: E402


def load_reference_model_module():
    tg, tgn, tgd = types.ModuleType("torch_geometric"), types.ModuleType("torch_geometric.nn"), types.ModuleType("torch_geometric.data")
    tgn.GATConv, tgn.global_mean_pool = ref_shim._EdgeListGATConv, ref_shim._global_mean_pool
    tgd.Data, tgd.Batch = getattr(ref_shim, "_Data", None), getattr(ref_shim, "_Batch", None)
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tgn, "torch_geometric.data": tgd})
    sys.path.insert(0, "/root/reference")
    return importlib.import_module("models.multimodal_model")


def main():
    mm = load_reference_model_module()
    for H, B in ((64, 9), (512, 33)):
        cfg = ref_shim.RefConfig(H=H)
        P = fo.init_head_params(H=H, num_emotions=cfg.num_emotions, seed=7)
        clf = mm.EmotionClassifier(cfg).double()
        clf.load_state_dict({k: v.double() for k, v in P.items() if k.split(".")[0] in ("classifier", "sentiment_classifier", "positive_classifier",
                                                                                        "negative_classifier")}, strict=True)
        clf.train()                                   # dropout p = 0
        heads = {n: nn.Linear(H, o).double() for n, o in (("valence_regressor", 1), ("arousal_regressor", 1), ("uncertainty_head", cfg.num_emotions))}
        for n, m in heads.items():
            m.load_state_dict({"weight": P[n + ".weight"].double(), "bias": P[n + ".bias"].double()})
        g = torch.Generator().manual_seed(1234)
        fused = torch.randn(B, H, generator=g).double().requires_grad_(True)
        target = torch.randint(0, cfg.num_emotions, (B,), generator=g)
        logits = clf(fused)
        probs = F.softmax(logits, dim=-1)
        val, aro = heads["valence_regressor"](fused), heads["arousal_regressor"](fused)
        unc = F.softmax(heads["uncertainty_head"](fused), dim=-1)
        ce = nn.CrossEntropyLoss(label_smoothing=0.1)(logits, target)
        total = ce + 0.05 * ((val ** 2).mean() + (aro ** 2).mean() + (unc ** 2).mean() + (probs ** 2).mean())
        total.backward()
        pg = {k: p.grad for k, p in clf.named_parameters() if p.grad is not None}
        for n, m in heads.items():
            pg[n + ".weight"], pg[n + ".bias"] = m.weight.grad, m.bias.grad
        rec = {"meta": {"H": H, "B": B, "num_emotions": cfg.num_emotions, "param_seed": 7, "feat_seed": 1234, "label_smoothing": 0.1},
               "fused": fused.detach(), "target": target, "logits": logits.detach(), "probs": probs.detach(), "valence": val.detach(),
               "arousal": aro.detach(), "uncertainty": unc.detach(), "ce": ce.detach(), "total": total.detach(),
               "dfused": fused.grad, "param_grads": {k: v.detach() for k, v in pg.items()}}
        path = os.path.join(ROOT, "tests", "golden", "heads", f"heads_h{H}.pt")
        torch.save(rec, path)
        print(f"{path}: ce={float(ce):.12f} ({os.path.getsize(path) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()
