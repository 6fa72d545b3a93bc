"""Import selected reference modules without executing the reference's package __init__s
(metrics/__init__.py pulls open3d, model/__init__.py pulls a missing file -- SURVEY.md §0.3).

A stub package named ``metrics`` with ``__path__ = [<ref>/metrics]`` is registered, then each
file is loaded as ``metrics.<name>`` so its relative imports (``from .alignment import *``,
metrics/eval_depth.py:4) still resolve.  Nothing is copied.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF = os.environ.get("UNIGEO_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "metrics"))


def _load(pkg: str, name: str):
    full = f"{pkg}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    if pkg not in sys.modules or getattr(sys.modules[pkg], "__ug_stub__", False) is False:
        stub = types.ModuleType(pkg)
        stub.__path__ = [os.path.join(REF, pkg)]
        stub.__ug_stub__ = True
        sys.modules[pkg] = stub
    spec = importlib.util.spec_from_file_location(full, os.path.join(REF, pkg, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def metrics_eval_depth():
    _load("metrics", "alignment")
    return _load("metrics", "eval_depth")


def metrics_eval_normal():
    return _load("metrics", "eval_normal")


def utils_geometry():
    return _load("utils", "geometry_utils")


def utils_io():
    return _load("utils", "io_utils")
