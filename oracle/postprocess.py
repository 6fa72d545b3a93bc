"""UniGeo's in-tree post-processing around the pipeline call -- oracle.

These are restated from files that ARE in /root/reference, and are pinned against
the unmodified reference functions by tests/golden/ (make_golden.py):

  prepare_input        model/depthcrafter.py:39-45   uint8 truncation, /255
  disparity_to_depth   model/depthcrafter.py:92-97   channel mean, clip-wide min-max, 1/(x+0.1)
  backproject          utils/geometry_utils.py:246-253
  surface_normal       utils/geometry_utils.py:9-70  5x5 box plane fit, lstsq, orientation
  prepare_output       model/depthcrafter.py:48-69   yz flip to OpenGL, stacking
  stablenormal_post    model/stablenormal.py:41-50   uint8 x-negation wraparound, /255*2-1

Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def prepare_input(images) -> np.ndarray:
    """model/depthcrafter.py:43-44: list of [3,H,W] 0..255 -> [Nf,H,W,3] float32 in [0,1]."""
    frames = [np.asarray(x).transpose(1, 2, 0).astype(np.uint8) for x in images]
    return np.stack(frames, axis=0).astype(np.float32) / 255.0


def disparity_to_depth(res: np.ndarray) -> np.ndarray:
    """model/depthcrafter.py:93-97: [Nf,H,W,3] -> [Nf,H,W] depth in [1/1.1, 10]."""
    res = res.sum(-1) / res.shape[-1]
    res = (res - res.min()) / (res.max() - res.min())
    return np.stack([1 / (x + 0.1) for x in res], axis=0)


def backproject(depth: np.ndarray, K: np.ndarray) -> np.ndarray:
    """utils/geometry_utils.py:246-253 (int64 grid x float32 intrinsics: float64 when depth is)."""
    h, w = depth.shape
    i, j = np.meshgrid(np.arange(w), np.arange(h), indexing="xy")
    z = depth
    x = (i - K[0, 2]) * z / K[0, 0]
    y = (j - K[1, 2]) * z / K[1, 1]
    return np.stack((x, y, z), axis=-1).reshape(h, w, 3)


def surface_normal(xyz: torch.Tensor, patch_size: int = 5) -> torch.Tensor:
    """utils/geometry_utils.py:9-70 on xyz [H,W,3] float32 -> unit normals [H,W,3].

    Per-pixel 3x3 systems (A^T A + 1e-6 I) n = A^T 1 over an un-normalised 5x5 box,
    solved with torch.linalg.lstsq in the reference's 4x4-tile sweep; H and W must be
    divisible by 4 (App. B.7: otherwise the reference leaves random values behind).
    Normals are flipped to face the camera (n.p <= 0).
    """
    p = xyz.permute(2, 0, 1)[None].float()                    # [1,3,H,W]
    x, y, z = p[:, 0:1], p[:, 1:2], p[:, 2:3]
    w = torch.ones((1, 1, patch_size, patch_size))
    pad = patch_size // 2

    def box(t):
        return F.conv2d(t, w, padding=pad)[0, 0]

    xx, yy, zz, xy, xz, yz = box(x * x), box(y * y), box(z * z), box(x * y), box(x * z), box(y * z)
    ata = torch.stack([xx, xy, xz, xy, yy, yz, xz, yz, zz], dim=-1).reshape(*xx.shape, 3, 3)
    ata = ata + 1e-6 * torch.eye(3)
    at1 = torch.stack([box(x), box(y), box(z)], dim=-1)[..., None]
    # :42-62 -- the reference sweeps a 4x4 grid of tiles (each grown by patch//2+1 pixels towards its
    # neighbours) and keeps the tile interiors; LAPACK's batched driver is sensitive to the batch
    # it is handed, so the sweep is restated as is to stay bit-identical.
    H, W = xx.shape
    tiles = 4
    th, tw = H // tiles, W // tiles
    grow = patch_size // 2 + 1
    n = torch.empty((H, W, 3))
    for ty in range(tiles):
        for tx in range(tiles):
            gy0 = grow if ty > 0 else 0
            gx0 = grow if tx > 0 else 0
            gy1 = grow if ty < tiles - 1 else 0
            gx1 = grow if tx < tiles - 1 else 0
            ys = slice(ty * th - gy0, (ty + 1) * th + gy1)
            xs = slice(tx * tw - gx0, (tx + 1) * tw + gx1)
            sol = torch.linalg.lstsq(ata[ys, xs], at1[ys, xs]).solution[..., 0]
            n[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw] = sol[gy0:gy0 + th, gx0:gx0 + tw]
    n = n / torch.sqrt(torch.sum(n ** 2, dim=2, keepdim=True))
    flip = torch.sum(n * xyz.float(), dim=2) > 0
    n[flip] *= -1
    return n


def prepare_output(depths: np.ndarray, intrinsics) -> dict:
    """model/depthcrafter.py:48-69: depth [Nf,H,W] + K list -> pred_depths / pred_normals (OpenGL)."""
    normals = []
    for d, K in zip(depths, intrinsics):
        pts = torch.from_numpy(backproject(d, np.asarray(K))).float()
        n = surface_normal(pts)
        n[:, :, 1:] = -n[:, :, 1:]
        normals.append(n)
    return {
        "pred_depths": torch.stack([torch.from_numpy(np.asarray(d)).float() for d in depths], 0),
        "pred_normals": torch.stack(normals, 0),
    }


def stablenormal_post(normals_u8) -> dict:
    """model/stablenormal.py:41-50: list of uint8 [H,W,3] -> pred_normals, zero depths.
    The x-flip negates a uint8 array, i.e. v -> (256 - v) mod 256 (App. B.10)."""
    out = []
    for n in normals_u8:
        n = np.array(n, dtype=np.uint8)
        n[:, :, 0] = (256 - n[:, :, 0].astype(np.int32)) % 256
        out.append(torch.from_numpy(n / 255.0 * 2 - 1).float())
    pn = torch.stack(out, 0)
    return {"pred_normals": pn, "pred_depths": torch.zeros_like(pn[..., 0])}
