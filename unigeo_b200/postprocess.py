"""UniGeo's per-clip post-processing around the pipeline call, on the GPU.

One C-ABI call, ``ug_depth_postprocess`` (``csrc/post.cu``), replaces, in order (reference file:line):
  disparity -> depth       model/depthcrafter.py:92-97   channel mean, clip-wide min-max, 1/(x+0.1)
  backprojection           utils/geometry_utils.py:246-253
  plane-fit normals        utils/geometry_utils.py:9-70  (5x5 un-normalised box, 1e-6 I, orientation)
  OpenCV -> OpenGL flip    model/depthcrafter.py:59
The reference runs this on the CPU at ~0.9 s/frame.  There is no torch or CPU fallback here: the
function needs the engine.
"""
from __future__ import annotations

import torch


def depth_and_normals(engine, frames: torch.Tensor, intrinsics: torch.Tensor):
    """frames [T,H,W,3] (pipeline output) + K [T,3,3] -> (pred_depths [T,H,W], pred_normals [T,H,W,3] OpenGL)."""
    return engine.depth_postprocess(frames, intrinsics)
