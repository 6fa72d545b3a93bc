"""Drop-in ``model`` plugin surface of the reference (/root/reference/model/__init__.py:3-4):
``import_class_from_module("model", config["model_name"])`` finds these by name
(configs/config_utils.py:3-6, eval.py:21-22)."""
from .depthcrafter import DepthCrafter
from .stablenormal import StableNormal

__all__ = ["DepthCrafter", "StableNormal"]
