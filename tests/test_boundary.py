"""Drop-in boundary, CPU-only: the module exposes the reference's class names and forward signatures,
every head registers exactly the reference's parameter names and shapes (taken from the golden vectors, whose
parameter-gradient keys/shapes were produced by the EXECUTED reference's named_parameters()), the shared library
loads and exports every symbol include/b200_fusion.h declares, and CPU tensors are refused (no fallback)."""
import ctypes
import glob
import importlib
import inspect
import os
import re

import pytest
import torch

from oracle import fusion_oracle as fo
from parity_util import CLS, Cfg

pkg = importlib.import_module("simple-multimodal_b200")
FL = pkg.fusion_layers
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.pt")))


def test_import_surface_matches_reference():
    # models/multimodal_model.py:6-9 imports these seven; MultimodalTransformer uses CrossModalTransformer (fusion_layers.py:102-107)
    for name in ("EarlyFusion", "LateFusion", "MultimodalTransformer", "CrossModalTransformer", "GraphFusion", "ContrastiveFusion",
                 "AdaptiveFusion", "HierarchicalFusion"):
        cls = getattr(FL, name)
        assert list(inspect.signature(cls.__init__).parameters)[:2] == ["self", "config"]
    for name in ("EarlyFusion", "LateFusion", "MultimodalTransformer", "GraphFusion", "AdaptiveFusion"):
        params = list(inspect.signature(getattr(FL, name).forward).parameters)
        assert params[:4] == ["self", "text_features", "audio_features", "video_features"]
    for name in ("ContrastiveFusion", "HierarchicalFusion"):
        sig = inspect.signature(getattr(FL, name).forward)
        assert list(sig.parameters)[:5] == ["self", "text_features", "audio_features", "video_features", "compute_contrastive_loss"]
        assert sig.parameters["compute_contrastive_loss"].default is False
    import simple_multimodal_b200 as alias                    # identifier-safe alias used in INTEGRATION.md
    assert alias.fusion_layers is FL


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_state_dict_layout_matches_executed_reference(path):
    rec = torch.load(path, weights_only=True)
    meta = rec["meta"]
    cfg = Cfg(H=meta["H"], heads=meta["heads"], graph_hidden=meta["graph_hidden"], graph_layers=meta["graph_layers"])
    head = getattr(FL, CLS[meta["kind"]])(cfg)
    ours = {k: tuple(v.shape) for k, v in head.named_parameters()}
    ref_keys = set(rec["param_grads"])                         # reference named_parameters() that received a gradient
    assert ref_keys <= set(ours), sorted(ref_keys - set(ours))
    if meta["full"]:
        for k, g in rec["param_grads"].items():
            assert ours[k] == tuple(g.shape), (k, ours[k], tuple(g.shape))
    # a reference-shaped state_dict (the oracle's init uses the reference's names) loads strictly
    P = fo.init_params(meta["kind"], H=meta["H"], heads=meta["heads"], graph_hidden=meta["graph_hidden"],
                       graph_layers=meta["graph_layers"], seed=1)
    res = head.load_state_dict(P, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_gat_accepts_pyg_23_parameter_names():
    cfg = Cfg(H=32, heads=4, graph_hidden=32, graph_layers=2)
    head = FL.GraphFusion(cfg)
    sd = {k.replace(".lin.weight", ".lin_src.weight"): v for k, v in head.state_dict().items()}
    for k in list(sd):
        if k.endswith("lin_src.weight"):
            sd[k.replace("lin_src", "lin_dst")] = sd[k]
    FL.GraphFusion(cfg).load_state_dict(sd, strict=True)


def test_graph_constructor_constraint_matches_reference():
    # all layers are built with in=fusion_hidden_size, so >1 layer needs graph_hidden == fusion_hidden
    with pytest.raises(pkg.B200FusionError):
        FL.GraphFusion(Cfg(H=64, graph_hidden=32, graph_layers=3))


def test_shared_library_exports_every_declared_symbol():
    text = open(os.path.join(ROOT, "include", "b200_fusion.h")).read()
    symbols = sorted(set(re.findall(r"\b(b200f_[a-z0-9_]+)\s*\(", text)))
    assert len(symbols) >= 30
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for s in symbols:
        assert hasattr(lib, s), f"{s} declared in include/b200_fusion.h but not exported"
    lib.b200f_version.restype = ctypes.c_int
    assert lib.b200f_version() >= 100


def test_cpu_tensors_are_refused_without_fallback():
    head = FL.EarlyFusion(Cfg(H=32, heads=4, graph_hidden=32))
    x = torch.randn(2, 32)
    with pytest.raises(pkg.B200FusionError):
        head(x, x, x)


def test_product_package_never_imports_the_oracle():
    src_dir = os.path.join(ROOT, "simple-multimodal_b200")
    for p in glob.glob(os.path.join(src_dir, "*.py")):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", open(p).read(), re.M), p


def test_ctypes_mirrors_have_the_size_and_field_offsets_of_the_c_structs(tmp_path):
    """The Python binding restates b200f_gemm_args / b200f_attn_args field for field; a host compiler's view of the header (plain C, gcc)
    must agree on the size and on the offset of every field -- a drifted mirror would hand the library garbage pointers."""
    import ctypes as C
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no host C compiler")
    L = importlib.import_module("simple-multimodal_b200._lib")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "b200_fusion.h"', "int main(void) {"]
    for cname, cls in (("b200f_gemm_args", L.GemmArgs), ("b200f_attn_args", L.AttnArgs)):
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in (("b200f_gemm_args", L.GemmArgs), ("b200f_attn_args", L.AttnArgs)):
        assert int(got[cname]) == C.sizeof(cls)
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
