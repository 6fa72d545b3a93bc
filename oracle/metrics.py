"""Restatement of the two metric functions eval.py applies to the hot path's outputs -- oracle.

  depth_evaluation(..., custom_mask, align_with_lstsq=True)   metrics/eval_depth.py:6-246
      + align_with_lstsq_torch                                 metrics/alignment.py:150-167
  normal_evaluation / compute_normal_metrics                   metrics/eval_normal.py:4-72
as called from eval.py:49 and eval.py:54.  Pinned against the unmodified reference functions
by tests/golden/metrics_kat.npz (tests/golden/make_golden.py).  The GPU box has no
/root/reference, so parity tests score both arms with these.
Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
import torch


def depth_evaluation(pred, gt, custom_mask=None, max_depth=80):
    """eval.py:49 call shape: [Nf,H,W] tensors, lstsq scale/shift alignment. Returns the metrics dict."""
    pred = torch.as_tensor(pred)
    gt = torch.as_tensor(gt)
    if pred.dim() == 3:                                   # :47-52
        w = pred.shape[-1]
        pred, gt = pred.reshape(-1, w), gt.reshape(-1, w)
        if custom_mask is not None:
            custom_mask = torch.as_tensor(custom_mask).reshape(-1, w)
    mask = (gt > 0) & (gt < max_depth)                    # :60-67
    p, g = pred[mask], gt[mask]
    A = np.hstack([p.numpy().reshape(-1, 1), np.ones((p.numel(), 1), dtype=p.numpy().dtype)])   # alignment.py:153-160
    sol = np.linalg.lstsq(A, g.numpy().reshape(-1, 1), rcond=None)[0]
    s, t = torch.tensor(sol[0]), torch.tensor(sol[1])
    p = s * p + t                                         # :86
    if custom_mask is not None:                           # :134-138
        mm = custom_mask[mask]
        p, g = p[mm], g[mm]
    abs_rel = torch.mean(torch.abs(p - g) / g).item()     # :141-143
    sq_rel = torch.mean(((p - g) ** 2) / g).item()
    rmse = torch.sqrt(torch.mean((p - g) ** 2)).item()
    p = torch.clamp(p, min=1e-5)                          # :152
    log_rmse = torch.sqrt(torch.mean((torch.log(p) - torch.log(g)) ** 2)).item()
    ratio = torch.maximum(p / g, g / p)
    n_valid = int(mask.sum().item() if custom_mask is None else mm.sum().item())
    res = {
        "Abs Rel": abs_rel, "Sq Rel": sq_rel, "RMSE": rmse, "Log RMSE": log_rmse,
        "delta < 1.": torch.mean((ratio < 1.0).float()).item(),
        "delta < 1.25": torch.mean((ratio < 1.25).float()).item(),
        "delta < 1.25^2": torch.mean((ratio < 1.25 ** 2).float()).item(),
        "delta < 1.25^3": torch.mean((ratio < 1.25 ** 3).float()).item(),
        "valid_pixels": n_valid,
    }
    if n_valid == 0:                                      # :217-227
        for k in list(res)[:-1]:
            res[k] = 0
    return res


def depth_maps(pred, gt, max_depth=80):
    """The three full-size maps depth_evaluation returns beside the dict (metrics/eval_depth.py:166-213,
    align_with_lstsq branch): |s p + t - gt| / gt over valid pixels (0 elsewhere), s p + t everywhere, gt over valid
    pixels (0 elsewhere); [Nf,H,W] inputs come back flattened to [Nf*H, W] like the reference's (:47-52)."""
    pred = torch.as_tensor(pred)
    gt = torch.as_tensor(gt)
    if pred.dim() == 3:
        w = pred.shape[-1]
        pred, gt = pred.reshape(-1, w), gt.reshape(-1, w)
    mask = (gt > 0) & (gt < max_depth)
    p, g = pred[mask], gt[mask]
    A = np.hstack([p.numpy().reshape(-1, 1), np.ones((p.numel(), 1), dtype=p.numpy().dtype)])   # alignment.py:153-160
    sol = np.linalg.lstsq(A, g.numpy().reshape(-1, 1), rcond=None)[0]
    s, t = torch.tensor(sol[0]), torch.tensor(sol[1])
    aligned = pred * s + t                                # :175
    err = torch.abs(aligned - gt) / gt                    # :178-181
    zeros = torch.zeros_like(gt)
    return torch.where(mask, err, zeros), aligned, torch.where(mask, gt, zeros)


def normal_evaluation(pred, gt, custom_mask=None):
    """eval.py:54 call shape: [Nf,H,W,3] tensors + bool mask [Nf,H,W]."""
    pred = torch.as_tensor(pred).permute(0, 3, 1, 2)      # eval_normal.py:63-64
    gt = torch.as_tensor(gt).permute(0, 3, 1, 2)
    mask = torch.as_tensor(custom_mask)
    dot = (pred * gt).sum(dim=1)
    err = dot / (torch.norm(pred, dim=1) * torch.norm(gt, dim=1) + 1e-6)       # :12-15
    err = torch.arccos(torch.clamp(err, -1.0, 1.0)) * 180.0 / np.pi             # :17-18
    e = err[mask]
    n = e.shape[0]
    out = {
        "normal mean": torch.mean(e),
        "normal median": torch.median(e),                 # lower of the two middle values (App. B.15)
        "normal rmse": torch.sqrt(torch.sum(e * e) / n),
        "angle < 5": 100.0 * (torch.sum(e < 5) / n),
        "angle < 7.5": 100.0 * (torch.sum(e < 7.5) / n),
        "angle < 11.25": 100.0 * (torch.sum(e < 11.25) / n),
        "angle < 22.5": 100.0 * (torch.sum(e < 22.5) / n),
        "angle < 30": 100.0 * (torch.sum(e < 30) / n),
    }
    return {k: v.item() for k, v in out.items()}
