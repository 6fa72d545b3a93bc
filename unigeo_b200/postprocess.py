"""UniGeo's per-clip post-processing around the pipeline call, on the GPU (torch library ops).

Mirrors, in order (reference file:line):
  disparity -> depth       model/depthcrafter.py:92-97   channel mean, clip-wide min-max, 1/(x+0.1)
  backprojection           utils/geometry_utils.py:246-253
  plane-fit normals        utils/geometry_utils.py:9-70  (5x5 un-normalised box, 1e-6 I, orientation)
  OpenCV -> OpenGL flip    model/depthcrafter.py:59
The reference runs this on the CPU at ~0.9 s/frame; here the 3x3 systems are solved in closed
form (adjugate) in float64 on the device, which is the exact solution the reference's float32
``lstsq`` approximates.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def disparity_to_depth(frames: torch.Tensor) -> torch.Tensor:
    """frames [T,H,W,3] in [0,1] -> depth [T,H,W] = 1 / (minmax(mean_c) + 0.1)."""
    res = frames.sum(-1) / frames.shape[-1]
    lo, hi = res.min(), res.max()
    res = (res - lo) / (hi - lo)
    return 1.0 / (res + 0.1)


def backproject(depth: torch.Tensor, K: torch.Tensor) -> torch.Tensor:
    """depth [T,H,W], K [T,3,3] -> camera-space points [T,H,W,3] (float64 like the numpy reference)."""
    T, H, W = depth.shape
    z = depth.double()
    K = K.double()
    i = torch.arange(W, device=depth.device, dtype=torch.float64).view(1, 1, W)
    j = torch.arange(H, device=depth.device, dtype=torch.float64).view(1, H, 1)
    x = (i - K[:, 0, 2].view(T, 1, 1)) * z / K[:, 0, 0].view(T, 1, 1)
    y = (j - K[:, 1, 2].view(T, 1, 1)) * z / K[:, 1, 1].view(T, 1, 1)
    return torch.stack((x, y, z), dim=-1)


def surface_normals(pts: torch.Tensor, patch_size: int = 5) -> torch.Tensor:
    """pts [T,H,W,3] float32 -> unit normals [T,H,W,3], flipped to face the camera."""
    p = pts.permute(0, 3, 1, 2).double()                       # box sums of fp32 inputs, accumulated in fp64
    T, _, H, W = p.shape
    x, y, z = p[:, 0:1], p[:, 1:2], p[:, 2:3]
    feats = torch.cat([x * x, x * y, x * z, y * y, y * z, z * z, x, y, z], dim=1)     # [T,9,H,W]
    w = torch.ones((9, 1, patch_size, patch_size), device=p.device, dtype=p.dtype)
    s = F.conv2d(feats, w, padding=patch_size // 2, groups=9)
    a, b, c, d, e, f = (s[:, k] for k in range(6))              # [[a,b,c],[b,d,e],[c,e,f]] + 1e-6 I
    a, d, f = a + 1e-6, d + 1e-6, f + 1e-6
    r0, r1, r2 = s[:, 6], s[:, 7], s[:, 8]
    # adjugate solve of the symmetric 3x3 system
    c00, c01, c02 = d * f - e * e, c * e - b * f, b * e - c * d
    c11, c12, c22 = a * f - c * c, b * c - a * e, a * d - b * b
    det = a * c00 + b * c01 + c * c02
    nx = (c00 * r0 + c01 * r1 + c02 * r2) / det
    ny = (c01 * r0 + c11 * r1 + c12 * r2) / det
    nz = (c02 * r0 + c12 * r1 + c22 * r2) / det
    n = torch.stack((nx, ny, nz), dim=-1)
    n = n / torch.sqrt((n * n).sum(-1, keepdim=True))
    flip = (n * pts.double()).sum(-1) > 0
    n = torch.where(flip.unsqueeze(-1), -n, n)
    return n.float()


def depth_and_normals(frames: torch.Tensor, intrinsics: torch.Tensor):
    """frames [T,H,W,3] (pipeline output) + K [T,3,3] -> (pred_depths [T,H,W], pred_normals [T,H,W,3] OpenGL)."""
    depth = disparity_to_depth(frames.float())
    pts = backproject(depth, intrinsics).float()
    n = surface_normals(pts)
    n = n * torch.tensor([1.0, -1.0, -1.0], device=n.device)
    return depth, n
