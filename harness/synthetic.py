"""Synthetic ``data`` dicts in the reference's Unified Data Format (dataset/Readme.md:22-33):
an analytic scene (tilted plane + three spheres, camera translating 1 cm/frame) so that GT
depth / normals / mask exist for the metric functions (SURVEY.md §8(d) "Synthetic inputs")."""
from __future__ import annotations

import numpy as np


def make_clip(num_frames: int, H: int, W: int, seed: int = 1234, scene_name: str = "synthetic_000") -> dict:
    rng = np.random.default_rng(seed)
    fx = fy = 0.9 * W
    cx, cy = W / 2.0, H / 2.0
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float32)
    jj, ii = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    dx, dy = (ii - cx) / fx, (jj - cy) / fy                    # ray directions (OpenCV: x right, y down, z fwd)
    rays = np.stack([dx, dy, np.ones_like(dx)], -1)
    centers = np.array([[-0.8, 0.2, 3.0], [0.6, -0.3, 2.2], [0.1, 0.5, 4.0]])
    radii = np.array([0.6, 0.4, 0.8])
    n_plane = np.array([0.15, -0.25, -1.0]); n_plane /= np.linalg.norm(n_plane)
    d_plane = 6.0
    data = {k: [] for k in ("images", "intrinsics", "extrinsics", "cam_coord", "cam_normal", "world_coord", "mask")}
    data["scene_name"] = scene_name
    for t in range(num_frames):
        cam = np.array([0.01 * t, 0.0, 0.0])                   # camera centre in world == cv-camera axes
        # plane: n.(cam + z*ray) + d = 0
        z = -(n_plane @ cam + d_plane) / (rays @ n_plane)
        nrm = np.broadcast_to(n_plane, rays.shape).copy()
        for c, r in zip(centers, radii):
            oc = cam - c
            a = (rays * rays).sum(-1); b = 2 * (rays @ oc); cc = oc @ oc - r * r
            disc = b * b - 4 * a * cc
            hit = disc > 0
            zs = np.where(hit, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
            closer = hit & (zs > 0.1) & (zs < z)
            z = np.where(closer, zs, z)
            p = cam + np.where(hit, zs, 0.0)[..., None] * rays
            ns = (p - c) / r
            nrm = np.where(closer[..., None], ns, nrm)
        z = np.clip(z, 0.5, 8.0)
        pts_cv = z[..., None] * rays                           # camera coords, OpenCV
        # Unified format stores OpenGL camera coords (y, z flipped): utils/io_utils.py:27-28 flips back
        pts_gl = pts_cv * np.array([1.0, -1.0, -1.0])
        nrm = nrm / np.linalg.norm(nrm, axis=-1, keepdims=True)
        nrm = np.where(((nrm * pts_cv).sum(-1) > 0)[..., None], -nrm, nrm)
        nrm_gl = nrm * np.array([1.0, -1.0, -1.0])
        tex = 127 + 60 * np.sin(3 * pts_cv[..., 0] + 0.5 * t * 0.01) + 40 * np.cos(2 * pts_cv[..., 1]) + 25 * np.sin(z)
        img = np.stack([tex, np.roll(tex, 3, 1) * 0.9 + 10, 255 - tex * 0.8], 0) + rng.normal(0, 3, (3, H, W))
        ext = np.eye(4, dtype=np.float32); ext[0, 3] = -cam[0]
        mask = rng.random((H, W)) > 0.05
        data["images"].append(np.clip(img, 0, 255).astype(np.float32))
        data["intrinsics"].append(K.copy())
        data["extrinsics"].append(ext)
        data["cam_coord"].append(pts_gl.transpose(2, 0, 1).astype(np.float32))
        data["cam_normal"].append(nrm_gl.transpose(2, 0, 1).astype(np.float32))
        data["world_coord"].append((pts_gl + cam * np.array([1.0, -1.0, -1.0])).transpose(2, 0, 1).astype(np.float32))
        data["mask"].append(mask)
    data["keyview_idx"] = 0
    return data


def gt_label(data: dict) -> dict:
    """What utils/io_utils.py:4-45 (prepare_gt_label) produces, restricted to the keys eval.py:49-54 reads."""
    import torch
    depths, normals, masks = [], [], []
    for i in range(len(data["images"])):
        cam = data["cam_coord"][i].astype(np.float32).copy()
        cam[1:] *= -1                                           # opengl -> opencv (:27-28)
        depths.append(torch.from_numpy(cam).permute(1, 2, 0)[..., -1])
        normals.append(torch.from_numpy(data["cam_normal"][i]).permute(1, 2, 0))
        masks.append(torch.from_numpy(data["mask"][i]).bool())
    return {"gt_depths": torch.stack(depths), "gt_normals": torch.stack(normals), "gt_masks": torch.stack(masks)}


def correlated_gt(ref_depth, amplitude: float = 0.15, offset: float = 0.1):
    """A ground truth that DEPENDS on a prediction: depth' = depth * (1 + amplitude * smooth field) + offset.

    Scoring two arms against the synthetic scene says little when the network is randomly initialised -- its output is
    uncorrelated with the scene, the least-squares alignment (metrics/alignment.py:150-167) then returns scale ~ 0 and
    both arms collapse onto the same constant, whatever they predicted.  Against THIS label (built from the oracle arm's
    own prediction) the alignment keeps scale ~ 1, Abs Rel sits near amplitude / 2, and any deviation of the other arm's
    depths moves its score: the parity number is sensitive to the thing it is meant to compare."""
    import torch
    d = torch.as_tensor(ref_depth, dtype=torch.float32)
    T, H, W = d.shape
    yy = torch.linspace(0, 1, H)[:, None]
    xx = torch.linspace(0, 1, W)[None, :]
    field = torch.sin(6.2831853 * 3 * xx) * torch.cos(6.2831853 * 2 * yy)
    tt = 1.0 + 0.25 * torch.cos(torch.arange(T, dtype=torch.float32) * 0.7)[:, None, None]
    gt = d * (1.0 + amplitude * field[None] * tt) + offset
    return {"gt_depths": gt, "gt_masks": torch.ones_like(gt, dtype=torch.bool)}
