"""Clip-sharded data parallelism over the clips of one scene.

The reference scores every clip independently (eval.py:33-99; clips built by
dataset/scannetpp/scannetpp.py:42-48 with ``clip_overlap`` shared frames), so clips shard across
ranks with NO collective on the denoising path.  The only exchange is the optional overlap
stitch: every clip's disparity is min-max normalised on its own and turned into depth = 1/(x+0.1)
(model/depthcrafter.py:95-96), so consecutive clips disagree by an affine map IN x (a projective one in depth);
one all-gather of the overlap frames lets every rank solve the 2-parameter chain on x and ramp the shared frames.  The stitched depth is an ADDITIONAL
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
    """Least-squares (s, t) with s*src + t ~= dst, from the 2x2 normal equations in float64.  A degenerate system
    (constant ``src``: det <= 1e-12 n sxx) keeps the scale, s = 1, and matches the means, t = mean(dst) - mean(src)."""
    x, y = src.double().flatten(), dst.double().flatten()
    n = x.numel()
    sx, sy, sxx, sxy = x.sum(), y.sum(), (x * x).sum(), (x * y).sum()
    det = n * sxx - sx * sx
    ok = det > 1e-12 * n * sxx
    safe = torch.where(ok, det, torch.ones_like(det))
    s = torch.where(ok, (n * sxy - sx * sy) / safe, torch.ones_like(det))
    t = (sy - s * sx) / n
    return s, t


def to_fit_space(d: torch.Tensor, disparity: bool, offset: float) -> torch.Tensor:
    """The adapter's depth is 1 / (x + 0.1) of the clip's min-max normalised disparity x (model/depthcrafter.py:95-96);
    two clips' x on shared frames differ by an affine map, their depths by a projective one -- so fit on x."""
    return 1.0 / d.double() - offset if disparity else d.double()


def from_fit_space(x: torch.Tensor, disparity: bool, offset: float) -> torch.Tensor:
    return 1.0 / (x + offset).clamp(min=1e-3) if disparity else x


def _gather(t: torch.Tensor, world: int) -> torch.Tensor:
    """[...] per rank -> [world, ...] on every rank: ONE all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    t = t.contiguous()
    if world == 1 or not dist.is_initialized():
        return t[None]
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather(list(out.unbind(0)), t)
    return out


def stitch_scene(local_depths: Sequence[torch.Tensor], local_ids: Sequence[int], num_clips: int, overlap: int,
                 rank: int = 0, world: int = 1, engine=None, space: str = "disparity", offset: float = 0.1,
                 device=None) -> List[torch.Tensor]:
    """local_depths[j]: [T,H,W] depth of clip local_ids[j].  Returns, for every local clip, its frames mapped into
    clip 0's frame, with its first ``overlap`` frames ramped from the previous clip's (mapped) last ``overlap`` frames
    with weights linspace(0,1,overlap).  ``space="disparity"`` fits the 2-parameter map on 1/depth - offset, where it
    is affine (see ``to_fit_space``); ``"depth"`` on the depths as given.

    Every rank enters the collectives, also one that holds no clip (``num_clips < world``): the frame size then comes
    from a 2-int all-reduce and the rank contributes zeros.  With ``engine`` (CUDA tensors) the fits, the chain and
    the ramp are the library's kernels (ug_stitch_fit / ug_stitch_apply); without one (CPU tensors: the gloo tests of
    the N > 1 host logic) the same arithmetic in torch float64."""
    if space not in ("disparity", "depth"):
        raise ValueError("space must be 'disparity' or 'depth'")
    disp = space == "disparity"
    if overlap <= 0 or num_clips == 1:
        return [d.clone() for d in local_depths]
    if local_depths:
        dev, dt = local_depths[0].device, local_depths[0].dtype
    else:
        dev, dt = torch.device(device) if device is not None else (engine.device if engine is not None else torch.device("cpu")), torch.float32
    if dev.type == "cuda" and engine is None:
        raise ValueError("stitch_scene on CUDA tensors needs the engine (no eager fallback on the device path)")
    hw = torch.tensor(list(local_depths[0].shape[-2:]) if local_depths else [0, 0], dtype=torch.int64, device=dev)
    if world > 1 and dist.is_initialized():
        dist.all_reduce(hw, op=dist.ReduceOp.MAX)
    H, W = int(hw[0]), int(hw[1])
    per_rank = (num_clips + world - 1) // world
    # [per_rank, 2, overlap, H, W]: head (first frames) and tail (last frames) of every local clip
    buf = torch.ones((per_rank, 2, overlap, H, W), device=dev, dtype=torch.float32)
    for j, d in enumerate(local_depths):
        buf[j, 0] = d[:overlap]
        buf[j, 1] = d[-overlap:]
    gathered = _gather(buf, world)                        # [world, per_rank, 2, overlap, H, W]
    tail = lambda k: gathered[k % world, k // world, 1]
    if engine is not None:
        chain = engine.stitch_fit(gathered, num_clips, disp, offset)
        return [engine.stitch_apply(d, tail(k - 1) if k > 0 else None, chain, k, overlap, disp, offset).to(dt)
                for d, k in zip(local_depths, local_ids)]
    head = lambda k: gathered[k % world, k // world, 0]
    # chain of 2-parameter maps into clip 0's frame: x_k_global = S[k] * x_k + Tt[k]; the K - 1 fits are independent
    # (clip k's head against clip k-1's RAW tail, composed afterwards -- the same least-squares solution)
    S = [torch.ones((), dtype=torch.float64, device=dev)]
    Tt = [torch.zeros((), dtype=torch.float64, device=dev)]
    for k in range(1, num_clips):
        a, b = fit_scale_shift(to_fit_space(head(k), disp, offset), to_fit_space(tail(k - 1), disp, offset))
        Tt.append(S[k - 1] * b + Tt[k - 1])
        S.append(S[k - 1] * a)
    ramp = torch.linspace(0.0, 1.0, overlap, device=dev, dtype=torch.float64).view(overlap, 1, 1)
    out = []
    for d, k in zip(local_depths, local_ids):
        g = S[k] * to_fit_space(d, disp, offset) + Tt[k]
        if k > 0:
            prev = S[k - 1] * to_fit_space(tail(k - 1), disp, offset) + Tt[k - 1]
            g[:overlap] = (1.0 - ramp) * prev + ramp * g[:overlap]
        out.append(from_fit_space(g, disp, offset).to(dt))
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
    if local_rows.dim() != 2:                      # a rank without clips (num_clips < world): [0] -> [0, K]
        local_rows = local_rows.reshape(0, 0)
    kk = torch.tensor([local_rows.shape[1]], dtype=torch.int64, device=local_rows.device)
    if world > 1 and dist.is_initialized():
        dist.all_reduce(kk, op=dist.ReduceOp.MAX)  # the row width from the ranks that have clips
    K = int(kk[0])
    per_rank = (num_clips + world - 1) // world
    buf = torch.full((per_rank, K), float("nan"), dtype=torch.float64, device=local_rows.device)
    if local_rows.shape[0]:
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
