"""``depth_evaluation`` / ``normal_evaluation`` on the B200 engine.

Same names, argument names, result keys and return shapes as the reference functions eval.py calls
(/root/reference/eval.py:49, :54):
  depth_evaluation(predicted_depth_original, ground_truth_depth_original, max_depth=80, custom_mask=None,
                   align_with_lstsq=True) -> (results, error_map, aligned_prediction, valid_gt)
                                                      /root/reference/metrics/eval_depth.py:6-246
  normal_evaluation(predicted_normal_original, ground_truth_normal_original, custom_mask=None) -> dict
                                                      /root/reference/metrics/eval_normal.py:36-72
Inputs may live on the host (the plugin's CPU tensors) or already on the GPU (``forward_device`` outputs),
in which case nothing crosses PCIe but the 19 result scalars.  Only the alignment eval.py uses
(``align_with_lstsq=True``, passed explicitly at eval.py:49) is implemented.  The DEFAULT stays the reference's
(``align_with_lstsq=False`` = median scaling, eval_depth.py:13): calling without the flag selects a mode that does not
exist on the device and raises NotImplementedError rather than silently scoring with another alignment.  No CPU path.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .config import get_config
from .engine import Engine

DEPTH_KEYS = ("Abs Rel", "Sq Rel", "RMSE", "Log RMSE", "delta < 1.", "delta < 1.25", "delta < 1.25^2",
              "delta < 1.25^3", "valid_pixels")
NORMAL_KEYS = ("normal mean", "normal median", "normal rmse", "angle < 5", "angle < 7.5", "angle < 11.25",
               "angle < 22.5", "angle < 30")

_engine: Optional[Engine] = None


def _default_engine() -> Engine:
    """A weight-less context on cuda:0 (the metric kernels need only its workspace and stream plumbing)."""
    global _engine
    if _engine is None:
        _engine = Engine(get_config("tiny"), device=0)
    return _engine


def _t(x):
    return torch.from_numpy(x) if isinstance(x, np.ndarray) else x


# options of the reference signature (metrics/eval_depth.py:6-23) that select another alignment / clipping mode
_OTHER_MODES = ("post_clip_min", "post_clip_max", "pre_clip_min", "pre_clip_max", "align_with_lad", "align_with_lad2",
                "metric_scale", "align_with_scale", "disp_input")
_IGNORED = ("lr", "max_iters", "use_gpu")          # only read by the modes above / a device hint


def depth_evaluation(predicted_depth_original, ground_truth_depth_original, max_depth=80, custom_mask=None,
                     align_with_lstsq=False, engine: Optional[Engine] = None, with_maps: bool = True, **options):
    unknown = [k for k in options if k not in _OTHER_MODES + _IGNORED]
    if unknown:
        raise TypeError(f"depth_evaluation() got unexpected keyword arguments {unknown}")
    other = [k for k in _OTHER_MODES if options.get(k) not in (None, False)]
    if not align_with_lstsq or other:
        raise NotImplementedError("only align_with_lstsq=True without clipping (the mode eval.py:49 uses) runs on the "
                                  f"device; requested: {other or 'median scaling'}")
    if max_depth is None:
        max_depth = float("inf")
    eng = engine or _default_engine()
    pred, gt = _t(predicted_depth_original), _t(ground_truth_depth_original)
    vals, maps = eng.depth_metrics(pred, gt, None if custom_mask is None else _t(custom_mask), max_depth, with_maps)
    results = dict(zip(DEPTH_KEYS, vals[:9]))
    results["valid_pixels"] = int(results["valid_pixels"])
    if results["valid_pixels"] == 0:                       # eval_depth.py:217-227 returns integer zeros
        results = {k: 0 for k in DEPTH_KEYS}
    if not with_maps:
        return results, None, None, None
    shape2d = (-1, gt.shape[-1])                           # the reference flattens [Nf,H,W] to [Nf*H,W] (:47-52)
    err, aligned, gtv = (m.reshape(shape2d) if gt.dim() == 3 else m for m in maps)
    return results, err, aligned, gtv


def normal_evaluation(predicted_normal_original, ground_truth_normal_original, custom_mask=None,
                      engine: Optional[Engine] = None):
    eng = engine or _default_engine()
    vals = eng.normal_metrics(_t(predicted_normal_original), _t(ground_truth_normal_original),
                              None if custom_mask is None else _t(custom_mask))
    return dict(zip(NORMAL_KEYS, vals))
