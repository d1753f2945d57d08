"""The drop-in itself (SURVEY 4 "Boundary"), build container only: the UNMODIFIED reference `models/multimodal_model.py` is
imported from /root/reference with `sys.modules["models.fusion_layers"]` pointing at this repo's `fusion_layers` -- the one-line
substitution INTEGRATION.md describes -- and its own `MultimodalEmotionModel.__init__` (models/multimodal_model.py:12-60) builds
the fusion head from our classes.  Checked: the reference model constructs for every fusion_type, no torch_geometric import is
needed any more, a `state_dict()` saved by the all-reference model (tests/golden/model/*.pt, oracle/make_golden_model.py) loads
strictly, and `state_dict()` / `named_parameters()` come out in the reference's order.  Runs in a subprocess (the reference's
top-level package is called `models`); skipped where /root/reference does not exist (the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = textwrap.dedent('''
    import importlib, os, sys
    import torch
    ROOT, REF = sys.argv[1], sys.argv[2]
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    import simple_multimodal_b200 as b200
    sys.modules["models.fusion_layers"] = b200.fusion_layers            # <- the substitution (INTEGRATION.md)
    enc = importlib.import_module("models.encoders")                     # reference code, unmodified
    mm = importlib.import_module("models.multimodal_model")              # reference code, unmodified
    assert "torch_geometric" not in sys.modules                          # the PyG dependency left with the reference's fusion_layers
    assert mm.HierarchicalFusion is b200.fusion_layers.HierarchicalFusion and mm.MultimodalTransformer is b200.fusion_layers.MultimodalTransformer
    em = importlib.import_module("simple-multimodal_b200.emotion_model")  # only for the toy random-init backbones (no network)

    class Cfg:
        text_model_name = audio_model_name = video_model_name = "toy"
        fusion_hidden_size, fusion_dropout, fusion_num_heads, num_emotions = 32, 0.0, 8, 7
        graph_hidden_size, graph_num_layers, graph_dropout, contrastive_temperature = 32, 3, 0.0, 0.07
        adapter_size, prompt_length = 8, 3

    def build(fusion_type):
        torch.manual_seed(21)
        bb = em.build_backbones("tiny")
        enc.AutoModel.from_pretrained = staticmethod(lambda _n: bb["text"])
        enc.Wav2Vec2Model.from_pretrained = staticmethod(lambda _n: bb["audio"])
        enc.ViTModel.from_pretrained = staticmethod(lambda _n: bb["video"])
        cfg = Cfg()
        cfg.fusion_type = fusion_type
        return mm.MultimodalEmotionModel(cfg)

    for ft in ("early", "late", "mult", "graph", "contrastive", "adaptive", "hierarchical"):
        model = build(ft)
        assert type(model.fusion_layer).__module__.startswith("simple-multimodal_b200"), type(model.fusion_layer)
    for ft in ("hierarchical", "mult", "late"):
        rec = torch.load(os.path.join(ROOT, "tests", "golden", "model", f"model_{ft}.pt"), weights_only=True)
        model = build(ft)
        res = model.load_state_dict(rec["state_dict"], strict=True)       # checkpoint of the all-reference model
        assert not res.missing_keys and not res.unexpected_keys
        assert list(model.state_dict().keys()) == list(rec["state_dict"].keys()), ft      # same names, same ORDER
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert all(shapes[k] == tuple(v.shape) for k, v in rec["state_dict"].items())
        names = [n for n, _ in model.named_parameters()]
        assert names == [k for k in rec["state_dict"] if k in set(names)], ft             # optimizer param order is the reference's
        assert set(rec["grads"]) <= set(names)
    # and the heads refuse to run on the CPU instead of silently computing somewhere else
    try:
        model.fusion_layer(torch.zeros(2, 32), torch.zeros(2, 32), torch.zeros(2, 32))
    except b200.B200FusionError:
        print("DROPIN_OK")
''')


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "models", "multimodal_model.py")), reason="needs /root/reference (build container)")
def test_unmodified_reference_model_builds_around_the_substituted_module():
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
