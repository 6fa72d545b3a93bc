"""Clip-sharded data parallelism over the clips of one scene.

The reference scores every clip independently (eval.py:33-99; clips built by
dataset/scannetpp/scannetpp.py:42-48 with ``clip_overlap`` shared frames), so clips shard across
ranks with NO collective on the denoising path.  The only exchange is the optional overlap
stitch: every clip is min-max normalised on its own (model/depthcrafter.py:95), so consecutive
clips disagree by an affine map; one all-gather of the overlap frames lets every rank solve the
2-parameter scale/shift chain and ramp the shared frames.  The stitched depth is an ADDITIONAL
output (``pred_depths_stitched``); per-clip ``pred_depths`` stay as the reference produces them.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def clips_of_rank(num_clips: int, rank: int, world: int) -> List[int]:
    """Round-robin: clip i -> rank i % world (weights are replicated, clips are the unit of work)."""
    return list(range(rank, num_clips, world))


def clip_starts(num_frames: int, clip_length: int, clip_overlap: int) -> List[int]:
    """Start frames as dataset/scannetpp/scannetpp.py:44 builds them: range(0, N, length - overlap)."""
    return list(range(0, num_frames, clip_length - clip_overlap))


def fit_scale_shift(src: torch.Tensor, dst: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Least-squares (s, t) with s*src + t ~= dst, from the 2x2 normal equations in float64."""
    x, y = src.double().flatten(), dst.double().flatten()
    n = x.numel()
    sx, sy, sxx, sxy = x.sum(), y.sum(), (x * x).sum(), (x * y).sum()
    det = n * sxx - sx * sx
    s = (n * sxy - sx * sy) / det
    t = (sy - s * sx) / n
    return s, t


def _gather(t: torch.Tensor, world: int) -> List[torch.Tensor]:
    if world == 1 or not dist.is_initialized():
        return [t]
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous())          # NCCL over NVLink on GPUs, gloo in the CPU tests
    return out


def stitch_scene(local_depths: Sequence[torch.Tensor], local_ids: Sequence[int], num_clips: int, overlap: int,
                 rank: int = 0, world: int = 1) -> List[torch.Tensor]:
    """local_depths[j]: [T,H,W] depth of clip local_ids[j].  Returns, for every local clip, its frames
    mapped into clip 0's scale, with its first ``overlap`` frames ramped from the previous clip's
    (aligned) last ``overlap`` frames with weights linspace(0,1,overlap)."""
    if overlap <= 0 or num_clips == 1:
        return [d.clone() for d in local_depths]
    T, H, W = local_depths[0].shape
    dev, dt = local_depths[0].device, local_depths[0].dtype
    per_rank = (num_clips + world - 1) // world
    # [per_rank, 2, overlap, H, W]: head (first frames) and tail (last frames) of every local clip
    buf = torch.zeros((per_rank, 2, overlap, H, W), device=dev, dtype=dt)
    for j, d in enumerate(local_depths):
        buf[j, 0] = d[:overlap]
        buf[j, 1] = d[-overlap:]
    gathered = _gather(buf, world)
    head = lambda k: gathered[k % world][k // world, 0]
    tail = lambda k: gathered[k % world][k // world, 1]
    # chain of affine maps into clip 0's frame: depth_k_global = S[k] * depth_k + Tt[k]
    S = [torch.ones((), dtype=torch.float64, device=dev)]
    Tt = [torch.zeros((), dtype=torch.float64, device=dev)]
    for k in range(1, num_clips):
        s, t = fit_scale_shift(head(k), S[k - 1] * tail(k - 1).double() + Tt[k - 1])
        S.append(s)
        Tt.append(t)
    ramp = torch.linspace(0.0, 1.0, overlap, device=dev, dtype=torch.float64).view(overlap, 1, 1)
    out = []
    for d, k in zip(local_depths, local_ids):
        g = S[k] * d.double() + Tt[k]
        if k > 0:
            prev = S[k - 1] * tail(k - 1).double() + Tt[k - 1]
            g[:overlap] = (1.0 - ramp) * prev + ramp * g[:overlap]
        out.append(g.to(dt))
    return out


def assemble_scene(stitched: Sequence[torch.Tensor], starts: Sequence[int], num_frames: int, overlap: int):
    """Concatenate stitched clips (all of them, in clip order) into one [num_frames,H,W] video: clip k
    contributes its frames except the last ``overlap`` ones, which clip k+1's ramped head supersedes."""
    frames = []
    for k, (d, s0) in enumerate(zip(stitched, starts)):
        last = k == len(stitched) - 1
        keep = d if last else d[: d.shape[0] - overlap]
        frames.append(keep)
    return torch.cat(frames, 0)[:num_frames]


def gather_metric_rows(local_ids: Sequence[int], local_rows: torch.Tensor, num_clips: int, rank: int = 0,
                       world: int = 1) -> torch.Tensor:
    """Per-clip metric rows (the 9 + 8 floats eval.py:49-57 puts in one CSV line) from every rank
    to every rank: local_rows [len(local_ids), K] float64 -> [num_clips, K] in clip order.  One fixed-size
    all-gather (NCCL on GPUs, gloo in the CPU test); ranks with fewer clips pad with NaN rows."""
    local_rows = local_rows.to(torch.float64)
    K = local_rows.shape[1]
    per_rank = (num_clips + world - 1) // world
    buf = torch.full((per_rank, K), float("nan"), dtype=torch.float64, device=local_rows.device)
    buf[: local_rows.shape[0]] = local_rows
    gathered = _gather(buf, world)
    out = torch.full((num_clips, K), float("nan"), dtype=torch.float64, device=local_rows.device)
    for r in range(world):
        ids = clips_of_rank(num_clips, r, world)
        if ids:
            out[torch.as_tensor(ids, device=out.device)] = gathered[r][: len(ids)]
    return out


def average_row(rows: torch.Tensor) -> torch.Tensor:
    """The 'Average' line metrics/save_utils.py:51-63 computes: column means over the clips, NaNs skipped
    (``mean(skipna=True)``; a column of NaNs only stays NaN)."""
    return torch.nanmean(rows, dim=0)


def export_metric_csv(path: str, seq_names: Sequence[str], rows: torch.Tensor, metric_names: Sequence[str]) -> None:
    """The table metrics/save_utils.py:65-90 writes (``MetricsManager.export_to_csv``): one line per sequence, an
    'Average' line, floats as %.5f, NaN as an empty field, header ',<metric>,...'.  Plain text, no pandas."""
    import os
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    body = torch.cat([rows.to(torch.float64).cpu(), average_row(rows.to(torch.float64).cpu())[None]], 0)

    def fmt(v: float) -> str:
        return "" if v != v else "%.5f" % v

    with open(path, "w") as f:
        f.write("," + ",".join(metric_names) + "\n")
        for name, row in zip(list(seq_names) + ["Average"], body.tolist()):
            f.write(str(name) + "," + ",".join(fmt(v) for v in row) + "\n")
