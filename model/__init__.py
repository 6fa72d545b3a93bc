"""Top-level ``model`` package: what the reference's ``import_class_from_module("model", config["model_name"])``
(/root/reference/configs/config_utils.py:3-6, called at eval.py:21) imports.  With this repository on ``sys.path``
ahead of the reference's own ``model/`` directory, eval.py picks up the B200 adapters with no edit
(reference list: /root/reference/model/__init__.py:1-5)."""
from unigeo_b200.model import DepthCrafter, StableNormal

__all__ = ["DepthCrafter", "StableNormal"]
