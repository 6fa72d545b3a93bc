"""BASELINE cfg4: one scene = 8 overlapping 25-frame clips (stride 20, overlap 5 -- the reference's
clip_length / clip_overlap, configs/depthcrafter_scannetpp.yaml:5-6, dataset/scannetpp/scannetpp.py:44) sharded
across the ranks; every rank runs the plugin call on its clips with NO communication, then ONE NCCL all-gather
of the overlap frames feeds the scale/shift chain + ramp (unigeo_b200/sharding.py), and the per-clip metric
rows (device metric kernels on device-resident outputs) are all-gathered.  Prints one JSON line on rank 0.

    python tools/run_scene.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_scene.py
options: --clips 8 --frames 25 --overlap 5 --height 384 --width 512 --steps 5 --config full
Timing: barrier + synchronize on both sides, device-independent wall clock of the slowest rank (the plugin call
is synchronous and includes host work), max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness.synthetic import gt_label, make_clip  # noqa: E402
from unigeo_b200 import sharding as sh  # noqa: E402
from unigeo_b200.model import DepthCrafter  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--frames", type=int, default=25)
ap.add_argument("--overlap", type=int, default=5)
ap.add_argument("--height", type=int, default=384)
ap.add_argument("--width", type=int, default=512)
ap.add_argument("--steps", type=int, default=5)           # the reference hard-codes 5 (model/depthcrafter.py:86)
ap.add_argument("--config", default="full")
ap.add_argument("--csv", default=None)                    # the per-clip metric table + 'Average' line (save_utils format)
a = ap.parse_args()

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

stride = a.frames - a.overlap
num_frames = stride * (a.clips - 1) + a.frames
scene = make_clip(num_frames, a.height, a.width, seed=4321, scene_name="synthetic_scene")
starts = [k * stride for k in range(a.clips)]


def clip_of(k):
    s = starts[k]
    return {key: (v[s:s + a.frames] if isinstance(v, list) else v) for key, v in scene.items()}


plug = DepthCrafter(config=a.config, dtype="fp16", weights="synthetic", device_weights=True,
                    num_inference_steps=a.steps, seed=99, device=local)
mine = sh.clips_of_rank(a.clips, rank, world)
plug.forward(clip_of(mine[0]))                                   # warm-up: workspace sizing, CLIP autotune


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


sync()
t0 = time.perf_counter()
# forward_device: depths / normals stay on the GPU -- the stitch and the metric kernels read them there, and only the
# 17 metric scalars per clip come back (eval.py would move 2 x 59 MB per clip to the host and score there)
outs = [plug.forward_device(clip_of(k)) for k in mine]
torch.cuda.synchronize()
t_clips = time.perf_counter() - t0
depths = [o["pred_depths"] for o in outs]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
stitched = sh.stitch_scene(depths, mine, a.clips, a.overlap, rank, world)
e1.record()
sync()
t_total = time.perf_counter() - t0
stitch_ms = e0.elapsed_time(e1)

# per-clip metric rows on the device (ug_depth_metrics / ug_normal_metrics), gathered with ONE fixed-size all-gather
from unigeo_b200 import metrics as DM  # noqa: E402
t1 = time.perf_counter()
local_rows = []
for k, o in zip(mine, outs):
    gt = gt_label(clip_of(k))
    d = DM.depth_evaluation(o["pred_depths"], gt["gt_depths"], custom_mask=gt["gt_masks"], align_with_lstsq=True,
                            engine=plug.engine, with_maps=False)[0]
    n = DM.normal_evaluation(o["pred_normals"], gt["gt_normals"], custom_mask=gt["gt_masks"], engine=plug.engine)
    local_rows.append([float(d[key]) for key in DM.DEPTH_KEYS] + [float(n[key]) for key in DM.NORMAL_KEYS])
table = sh.gather_metric_rows(mine, torch.tensor(local_rows, dtype=torch.float64, device=plug.device), a.clips, rank,
                              world)
torch.cuda.synchronize()
t_metrics = time.perf_counter() - t1
avg = sh.average_row(table)
if rank == 0 and a.csv:
    sh.export_metric_csv(a.csv, [f"synthetic_scene_clip{k}" for k in range(a.clips)], table,
                         list(DM.DEPTH_KEYS + DM.NORMAL_KEYS))
rows = [(k, float(table[k, 0])) for k in range(a.clips)]
t = torch.tensor([t_clips, t_total, stitch_ms, t_metrics], device=plug.device, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# seam check: after stitching, the ramped head of clip k starts exactly at clip k-1's (aligned) tail
seam = None
if world == 1:
    seam = max(float((stitched[k][0] - stitched[k - 1][-a.overlap]).abs().max()) for k in range(1, a.clips))
if rank == 0:
    tc, tt, sm, tm = t.tolist()
    print(json.dumps({
        "workload": f"scene of {a.clips} clips x {a.frames} frames (overlap {a.overlap}, {num_frames} distinct frames) at "
                    f"{a.height}x{a.width}, {a.steps} denoising steps per clip, clip-sharded over {world} GPU(s)",
        "n_gpus": world, "clips": a.clips, "clips_per_rank_max": len(sh.clips_of_rank(a.clips, 0, world)),
        "seconds_clips": tc, "seconds_total_incl_stitch": tt, "stitch_ms_device": sm,
        "clips_per_s": a.clips / tt, "denoising_steps_per_s": a.clips * a.steps / tt,
        "collective": "one all_gather of [clips/rank, 2, overlap, H, W] fp32 overlap frames (NCCL)" if world > 1 else "none",
        "abs_rel_per_clip": [round(v, 5) for _, v in rows], "seam_max_abs_after_stitch": seam,
        "metrics_seconds_all_local_clips": tm,
        "metrics": "device kernels (ug_depth_metrics / ug_normal_metrics) on device-resident outputs; rows all-gathered",
        "average_row": {key: round(float(v), 5) for key, v in zip(DM.DEPTH_KEYS + DM.NORMAL_KEYS, avg.tolist())}}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
