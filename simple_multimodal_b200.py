"""Identifier-safe alias of the `simple-multimodal_b200` package (the directory name has a hyphen)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("simple-multimodal_b200")
sys.modules[__name__] = _pkg
