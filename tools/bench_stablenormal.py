"""BASELINE cfg3: StableNormal, 512x512 frames, 10 refinement steps, full SD-2.1 class UNet + ControlNet
(865.9 M + 363.1 M parameters, seeded random-init), fp16, one B200.  Prints one JSON line: refinement
steps/s (CUDA events, inputs resident), per-family kernel table of one instrumented step, and the wall
clock of the plugin call with host buffers.  Development / profiling aid (run under gpurun):
    python tools/bench_stablenormal.py [frames] [height] [width]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from unigeo_b200.model import StableNormal  # noqa: E402

F_ = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
W = int(sys.argv[3]) if len(sys.argv) > 3 else 512
STEPS = 10
plug = StableNormal(config="full", dtype="fp16", weights="synthetic", device_weights=True, num_inference_steps=STEPS,
                    seed=1)
eng = plug.engine
g = torch.Generator().manual_seed(0)
h, w = H // 8, W // 8
il = torch.randn(F_, 4, h, w, generator=g).cuda()
lat = torch.randn(F_, 4, h, w, generator=g).cuda()
for _ in range(3):                        # eager, capture, replay (the loop is a CUDA graph from its second call on)
    eng.refine_2d("unet2d", "controlnet", il, lat, STEPS)
torch.cuda.synchronize()
eng.launch_count(reset=True)
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.refine_2d("unet2d", "controlnet", il, lat, STEPS); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
launches = eng.launch_count() // 3
ms = sorted(ts)[1]
assert torch.isfinite(out).all()
eng.profile(True, by_shape=True)
eng.refine_2d("unet2d", "controlnet", il, lat, 1)
rows = eng.profile_read(cap=2048)
eng.profile(False)
# per-kernel rooflines of the step (BASELINE cfg3): tensor-bound kernels against the measured cuBLAS peak, the rest
# against the measured copy bandwidth; one frame of 64 x 64 latents is far too small to fill 148 SMs (M = 4096 ..
# 64 rows per GEMM), so launch latency, not either roof, bounds most launches
import os  # noqa: E402
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
shape_rows = []
for r in sorted(rows, key=lambda r: -r["ms"])[:25]:
    tensor = r["name"].startswith(("tapgemm", "fmha"))
    us = 1e3 * r["ms"] / max(r["launches"], 1)
    ach = (r["flops"] / 1e12 if tensor else r["bytes"] / 1e9) / (r["ms"] * 1e-3) if r["ms"] else 0.0
    peak = peaks["bf16_tflops_sustained"] if tensor else peaks["hbm_gbs"]
    shape_rows.append({"kernel": r["name"], "launches": r["launches"], "us_per_launch": round(us, 2), "bound": "tensor" if tensor else "hbm",
                       "achieved": round(ach, 1), "unit": "TFLOP/s" if tensor else "GB/s", "frac": round(ach / peak, 4)})
for r in rows:
    r["name"] = r["name"].split(" ")[0]
fam = {}
for r in rows:
    k = r["name"].split(".")[0]
    d = fam.setdefault(k, {"ms": 0.0, "flops": 0.0, "launches": 0})
    d["ms"] += r["ms"]; d["flops"] += r["flops"]; d["launches"] += r["launches"]
tot = sum(d["ms"] for d in fam.values())
flops = sum(d["flops"] for d in fam.values())
data = {"images": [np.random.default_rng(i).random((3, H, W)).astype(np.float32) * 255 for i in range(F_)]}
for _ in range(2):
    plug.forward(data)
torch.cuda.synchronize()
t0 = time.perf_counter(); o = plug.forward(data); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(json.dumps({
    "workload": f"StableNormal {F_}x{H}x{W}, {STEPS} DDIM refinement steps, UNet2D + ControlNet (full config), fp16",
    "refine_ms_per_step": ms / STEPS, "refine_steps_per_s": F_ * STEPS / (ms * 1e-3) / F_, "frames": F_,
    "launches_per_step": launches / STEPS, "instrumented_step_ms": tot, "step_tflop_algorithmic": flops / 1e12,
    "step_tflops_achieved": flops / 1e12 / (ms / STEPS * 1e-3),
    "kernels": {k: {"ms": round(d["ms"], 3), "launches": d["launches"],
                    "tflops": round(d["flops"] / 1e9 / d["ms"], 1) if d["ms"] else 0} for k, d in
                sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
    "top_kernels_by_shape": shape_rows, "peaks": {"tensor_tflops": peaks["bf16_tflops_sustained"], "hbm_gbs": peaks["hbm_gbs"]},
    "plugin_forward_seconds": dt, "plugin_normals_shape": list(o["pred_normals"].shape)}))
