#!/usr/bin/env python
"""Headline benchmark: denoising-steps/sec on 25x384x512 clips (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W             our arm (libunigeo_b200.so on B200)
    python bench.py --impl reference --steps K --warmup W     CPU arm (oracle restatement, host cores)

A "step" is one denoising step of the hot path on one clip: scale_model_input + concat of the
conditioning latents, one full SVD-XT spatio-temporal UNet forward (1.52 B params, 23.16 TFLOP at
25x384x512) and the Euler update -- exactly what ug_denoise_clip runs per step.  Weights are seeded
random-init of the real architecture, inputs synthetic (no datasets/checkpoints offline).

Timing: W untimed steps, then K steps bracketed by barrier + cuda.synchronize, CUDA events on the
launching stream, max over ranks.  One step streams ~3 GB of weights and >10 GB of activations, far
beyond the 126 MB L2, so no extra flush is needed between steps ("inputs larger than L2").
`e2e` is the same metric through the plugin call DepthCrafter.forward(data) with HOST buffers
(H2D of the frames, CLIP, VAE encode, the denoising loop, VAE decode, post-processing, D2H).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET_TFLOP_PER_STEP = {(25, 384, 512): 23.16, (49, 576, 1024): 156.6}       # SURVEY.md §8(d) algorithmic counts
# algorithmic work of the VAE stages per clip (TFLOP, GB of minimal 16-bit traffic)
VAE_WORK = {(25, 384, 512): {"decode": (56.9, 68.5), "encode": (20.8, 23.3)},
            (49, 576, 1024): {"decode": (340.0, 402.8), "encode": (128.0, 137.1)}}
# per-launch DRAM bytes from an ncu launch list of THIS build (tools/ncu_traffic.py stamps the source digest; a stale
# file -- kernels changed since the capture -- is reported as traffic: null)
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")


def source_digest():
    """sha256 over the CUDA sources + nvcc flags: the same digest unigeo_b200/build.py stamps the library with."""
    from unigeo_b200 import build as B
    deps = [os.path.join(B.CSRC, f) for f in os.listdir(B.CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(B.HERE, "..", "include", "unigeo_b200.h"))
    return B._digest(deps)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="full", choices=["full", "tiny"])
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--height", type=int, default=384)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--e2e-steps", type=int, default=25, help="num_inference_steps of the e2e plugin call (cfg2: 25)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (fp32 oracle on the GPU, rank 0)")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the torch-eager fp16 arm on the same GPU")
    ap.add_argument("--no-scene", action="store_true", help="skip the cfg4 scene (8 overlapping clips, stitch all-gather)")
    ap.add_argument("--scene-clips", type=int, default=8)
    ap.add_argument("--scene-overlap", type=int, default=5)
    ap.add_argument("--scene-steps", type=int, default=5, help="the reference's shipped num_inference_steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-frames", type=int, default=2)
    ap.add_argument("--profile-out", default=None, help="write the per-kernel launch table (json) here")
    return ap.parse_args()


def workload_name(a):
    return f"DepthCrafter {a.frames}x{a.height}x{a.width} clip, SVD-XT UNet denoising step ({a.config} config)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        busy = [c for c in sm if c > 0]
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_unet_sample(a, steps, warmup):
    """Times the oracle restatement of ONE denoising step on the host cores, on a bounded sample:
    `cpu_sample_frames` of the clip's frames at full resolution, full SVD-XT config, fp32.  Cost is
    linear in T for every layer except temporal attention (0.1 % of the FLOPs), so steps/s for the
    whole clip = measured / (T / sample_frames)."""
    import torch
    from oracle import scheduler as S
    from oracle.pipeline import added_time_ids
    from oracle.unet_st import unet_forward
    from unigeo_b200.config import get_config
    from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)          # torchrun pins OMP_NUM_THREADS=1; the CPU arm uses every core
    cfg = get_config(a.config)
    Ts = max(1, min(a.cpu_sample_frames, a.frames))
    h, w = a.height // 8, a.width // 8
    sd = synthetic_state_dict(unet_param_shapes(cfg.unet), 1000)
    g = torch.Generator().manual_seed(1234)
    lat = torch.randn(1, Ts, 4, h, w, generator=g)
    cond = torch.randn(1, Ts, 4, h, w, generator=g)
    enc = torch.randn(1, Ts, cfg.clip_embed_dim, generator=g)
    sig = S.karras_sigmas(25, cfg.sigma_min, cfg.sigma_max, cfg.rho)
    ids = added_time_ids(cfg)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            x = torch.cat([S.scale_model_input(lat, sig[i % 24]), cond], dim=2)
            v = unet_forward(sd, cfg.unet, x, 0.25 * math.log(sig[i % 24]), enc, ids)
            lat = S.euler_step(v, lat, sig[i % 24], sig[i % 24 + 1])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    per_step_clip = (sum(times) / len(times)) * (a.frames / Ts)
    sample = (f"oracle (torch fp32, {cores} threads) full denoising step on {Ts} of {a.frames} frames at "
              f"{a.height}x{a.width}, x{a.frames / Ts:g} (cost linear in T), mean of {len(times)} steps")
    return 1.0 / per_step_clip, per_step_clip * 1e3, cores, sample


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, ms, cores, sample = cpu_unet_sample(a, max(1, a.steps), max(0, a.warmup))
    print(json.dumps({
        "impl": "reference", "metric": "denoising-steps/sec", "value": val, "unit": "steps/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames": a.frames, "height": a.height, "width": a.width},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "extrapolated": f"UNet step on {min(a.cpu_sample_frames, a.frames)} of {a.frames} frames x "
                                f"{a.frames / max(1, min(a.cpu_sample_frames, a.frames)):g}; CLIP / VAE stages not included"},
        "note": "the reference's own diffusers path is not installable offline (no diffusers/weights); "
                "this is the oracle port of the same algorithm (parity unpinned, see DESIGN.md)",
    }))


# ----------------------------------------------------------------------------- B200 arm
def run_b200(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    from unigeo_b200.config import get_config
    from unigeo_b200.engine import Engine
    from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes, vae_param_shapes
    cfg = get_config(a.config)
    T, h, w = a.frames, a.height // 8, a.width // 8
    eng = Engine(cfg, dtype=a.dtype, device=local)
    # seeded random-init weights drawn directly on the device (values are irrelevant to the timing)
    eng.load_state_dict("unet", synthetic_state_dict(unet_param_shapes(cfg.unet), 1000, torch.float16, eng.device))
    eng.load_state_dict("vae", synthetic_state_dict(vae_param_shapes(cfg.vae), 2000, torch.float16, eng.device))
    if not a.no_e2e:
        from unigeo_b200.clip_embed import ClipEmbedder
        ClipEmbedder(eng, device_weights=True)                     # ViT-H/14 image encoder for the e2e plugin call
    eng.finalize()
    eng.prepare(T, h, w)
    dev = eng.device
    g = torch.Generator().manual_seed(1234 + rank)
    cond = torch.randn(T, 4, h, w, generator=g).to(dev)
    noise = torch.randn(T, 4, h, w, generator=g).to(dev)
    enc = torch.randn(T, cfg.clip_embed_dim, generator=g).to(dev)
    ids = [cfg.fps_id, cfg.motion_bucket_id, cfg.noise_aug_strength]
    eng.set_clip_context(enc)
    out = torch.empty_like(cond)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    eng.denoise(cond, noise, ids, max(a.warmup, 1), out=out)         # warm-up steps (untimed)
    sync()
    # the loop becomes a CUDA graph on its second call with one signature: run the K-step signature twice untimed
    # (eager, then capture) so that the timed region is a pure replay -- what a clip loop sees from clip 3 on
    graph_warm = 0 if os.environ.get("UG_NO_GRAPH") else 2
    for _ in range(graph_warm):
        eng.denoise(cond, noise, ids, a.steps, out=out)
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    eng.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    eng.denoise(cond, noise, ids, a.steps, out=out)                  # EXACTLY K timed steps
    e1.record()
    sync()
    launches = eng.launch_count()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    clocks = sampler.stop() if rank == 0 else None
    assert torch.isfinite(out).all(), "non-finite latents"
    ms_per_step = ms_total / a.steps
    value = world * a.steps / (ms_total / 1e3)

    # ---- roofline of the dominant kernel: one extra instrumented step (events after every launch)
    roof, table = None, None
    if rank == 0:
        eng.profile(True)
        eng.denoise(cond, noise, ids, 1, out=out)
        table = eng.profile_read()
        eng.profile(False)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)      # kernel timed inside a long step
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md ~1.4 PF sustained)"
        tg = [r for r in table if r["name"].startswith("tapgemm")]
        tot_ms = sum(r["ms"] for r in table)
        tg_ms, tg_fl, tg_n = sum(r["ms"] for r in tg), sum(r["flops"] for r in tg), sum(r["launches"] for r in tg)
        ach = tg_fl / (tg_ms * 1e-3) / 1e12 if tg_ms > 0 else 0.0
        traffic = {}
        try:
            traffic = json.load(open(TRAFFIC_JSON))
            if traffic.get("source_digest") != source_digest():
                traffic = {"source": f"{os.path.basename(TRAFFIC_JSON)} is STALE (kernels changed since that ncu "
                                     "capture): traffic not reported"}
        except (OSError, ValueError):
            pass
        roof = {"bound": "tensor", "kernel": "tapgemm_kernel<CTAS, LN> (tcgen05 implicit GEMM: conv3x3 / temporal conv / linear / attention GEMMs)",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic.get("tapgemm", {}).get("dram_bytes_per_launch"),
                "traffic_source": traffic.get("source"),
                "algorithmic_bytes_per_launch_avg": sum(r["bytes"] for r in tg) / max(tg_n, 1),
                "peak_source": peak_src, "launches_per_step": tg_n, "flops_per_launch_avg": tg_fl / max(tg_n, 1),
                "avg_launch_us": 1e3 * tg_ms / max(tg_n, 1), "share_of_step": tg_ms / tot_ms if tot_ms else None,
                "step_tflops_algorithmic": UNET_TFLOP_PER_STEP.get((a.frames, a.height, a.width)),
                "how": "CUDA events after every launch of one extra step on the launching stream (ug_ctx_profile)"}
        # the other kernel families of the step, each against the bound that applies (north_star: tensor-pipe on the
        # attention path, HBM GB/s on the normalisation / VAE path)
        hbm = peaks.get("hbm_gbs", 6500.0)
        fam = {}
        for r in table:
            k = r["name"].split(".")[0].split(" ")[0]
            f_ = fam.setdefault(k, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            for q in ("ms", "flops", "bytes", "launches"):
                f_[q] += r[q]
        others = []
        for k, f_ in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            if k == "tapgemm" or f_["ms"] <= 0:
                continue
            tensor = k.startswith("fmha")
            achv = (f_["flops"] / 1e12 if tensor else f_["bytes"] / 1e9) / (f_["ms"] * 1e-3)
            others.append({"kernel": k, "bound": "tensor" if tensor else "hbm", "achieved": achv,
                           "unit": "TFLOP/s" if tensor else "GB/s", "peak": peak_tf if tensor else hbm,
                           "frac": achv / (peak_tf if tensor else hbm), "launches_per_step": f_["launches"],
                           "share_of_step": f_["ms"] / tot_ms,
                           "traffic": traffic.get(k, {}).get("dram_bytes_per_launch")})
        roof["other_kernels"] = others[:6]
        roof["stages"] = vae_stage_rooflines(a, eng, cfg, peak_tf, hbm)
        if a.profile_out:
            with open(a.profile_out, "w") as f:
                json.dump({"ms_per_step_timed": ms_per_step, "instrumented_step_ms": tot_ms, "kernels": sorted(table, key=lambda r: -r["ms"])}, f, indent=1)

    # ---- e2e through the plugin call with host buffers
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, eng, cfg, world, rank, dev)

    scene = None
    if not a.no_scene and not a.no_e2e and a.config == "full":
        scene = run_scene(a, eng, cfg, world, rank, dev)
    parity = lib_base = None
    if rank == 0 and world == 1 and a.config == "full":
        if not a.no_parity and not a.no_e2e:
            parity = run_parity(a, eng, cfg, dev)
        if not a.no_library_baseline:
            lib_base = run_library_baseline(a, eng, cfg, dev)

    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        del eng
        torch.cuda.empty_cache()
        v, ms, cores, sample = cpu_unet_sample(a, 1, 0)
        cpu = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "denoising-steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": workload_name(a), "frames": T, "height": a.height, "width": a.width,
                       "clips_per_gpu": 1, "l2": "inputs larger than L2 (3 GB weights + >10 GB activations per step)",
                       "weights": "seeded random-init, SVD-XT architecture (1.52 B params)",
                       "launch": "CUDA graph replay of the K-step loop" if graph_warm else "eager launches"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "scene": scene, "parity": parity, "library_baseline": lib_base,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def vae_stage_rooflines(a, eng, cfg, peak_tf, hbm):
    """VAE encode / temporal decode of one clip through the C ABI, CUDA events, inputs resident: the whole stage
    against both bounds (report achieved GB/s vs the minimal traffic AND TFLOP/s)."""
    import torch
    work = VAE_WORK.get((a.frames, a.height, a.width))
    if work is None:
        return None
    T, H, W = a.frames, a.height, a.width
    g = torch.Generator().manual_seed(7)
    frames = torch.rand(T, H, W, 3, generator=g).to(eng.device)
    lat = (torch.randn(T, 4, H // 8, W // 8, generator=g) * 0.5).to(eng.device)
    out = {}
    for name, fn in (("encode", lambda: eng.vae_encode_frames(frames)),
                     ("decode", lambda: eng.vae_decode_frames(lat, cfg.decode_chunk_size))):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        tf, gb = work[name]
        out["vae_" + name] = {"ms_per_clip": ms, "tflops": tf / (ms * 1e-3), "frac_tensor": tf / (ms * 1e-3) / peak_tf,
                              "gbs_algorithmic": gb / (ms * 1e-3), "frac_hbm": gb / (ms * 1e-3) / hbm,
                              "work": {"tflop": tf, "gb": gb}}
    return out


def run_e2e(a, eng, cfg, world, rank, dev):
    """steps/s through the reference-facing plugin call with HOST buffers, every rank one clip."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from harness.synthetic import make_clip
    from unigeo_b200.clip_embed import ClipEmbedder
    from unigeo_b200.model.depthcrafter import DepthCrafter
    from unigeo_b200.pipeline import DepthCrafterPipelineB200
    plug = _plugin(a, eng, cfg, dev, a.e2e_steps, 1234 + rank)
    data = make_clip(a.frames, a.height, a.width, seed=1234 + rank)
    for _ in range(2):                                         # warm-up: workspace sizing, then the CUDA-graph capture
        plug.forward(data)                                     # of the denoising loop (2nd call with one signature)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = plug.forward(data)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = t.item()
    h2d = a.frames * a.height * a.width * 3 * 4
    d2h = out["pred_depths"].numel() * 4 + out["pred_normals"].numel() * 4
    return {"value": world * a.e2e_steps / dt, "unit": "steps/s", "h2d_bytes_per_step": h2d / a.e2e_steps,
            "d2h_bytes_per_step": d2h / a.e2e_steps, "clip_seconds": dt, "num_inference_steps": a.e2e_steps,
            "call": "unigeo_b200.model.DepthCrafter.forward(data) (H2D images + prepare_input + CLIP ViT-H + VAE encode + denoise + "
                    "VAE decode + depth/normal post-processing + D2H)"}


def _plugin(a, eng, cfg, dev, steps, seed):
    from unigeo_b200.clip_embed import ClipEmbedder
    from unigeo_b200.model.depthcrafter import DepthCrafter
    from unigeo_b200.pipeline import DepthCrafterPipelineB200
    plug = object.__new__(DepthCrafter)                       # reuse the already-loaded engine (one copy of weights)
    plug.device, plug.cfg, plug.dtype, plug.engine = dev, cfg, a.dtype, eng
    plug.num_inference_steps, plug.seed, plug._stage = steps, seed, None
    plug.pipeline = DepthCrafterPipelineB200(cfg, eng, ClipEmbedder.__new__(ClipEmbedder))
    plug.pipeline.clip.engine = eng
    return plug


def run_scene(a, eng, cfg, world, rank, dev):
    """BASELINE cfg4 on a measured path: ONE scene = `scene_clips` overlapping clips (stride frames - overlap, as
    dataset/scannetpp/scannetpp.py:44 builds them) sharded clip i -> rank i % world; every rank runs the plugin call on
    its clips (no communication), then the ONE exchange of the path -- an all-gather of every clip's first / last
    `overlap` depth frames (NCCL over NVLink) -- feeds the stitch kernels (ug_stitch_fit / ug_stitch_apply), the
    metric kernels score the device-resident outputs and one all-gather collects the 17-float rows.  Total work is
    fixed (strong scaling over N); timed on the host clock of the slowest rank around barriers (the plugin call is
    synchronous), stitch + gathers also with CUDA events."""
    import torch
    import torch.distributed as dist
    from harness.synthetic import gt_label, make_clip
    from unigeo_b200 import metrics as DM
    from unigeo_b200 import sharding as sh
    K, ov, T = a.scene_clips, a.scene_overlap, a.frames
    stride = T - ov
    n_frames = stride * (K - 1) + T
    sc = make_clip(n_frames, a.height, a.width, seed=4321, scene_name="synthetic_scene")
    clip_of = lambda k: {key: (v[k * stride:k * stride + T] if isinstance(v, list) else v) for key, v in sc.items()}
    plug = _plugin(a, eng, cfg, dev, a.scene_steps, 99)
    mine = sh.clips_of_rank(K, rank, world)
    # GT shaping and its upload are dataset work (utils/io_utils.py:4-45), not the path: resident before the clock starts
    gts = {k: {n: v.to(dev) for n, v in gt_label(clip_of(k)).items()} for k in mine}
    plug.forward_device(clip_of(mine[0] if mine else 0))      # warm-up: workspace / graph of this step count
    dummy = [torch.full((T, a.height, a.width), 1.0 + 0.1 * k, device=dev) for k in mine]
    sh.stitch_scene(dummy, mine, K, ov, rank, world, engine=eng, device=dev)   # warm-up: workspace sizing, NCCL channel
    del dummy

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sync()
    t0 = time.perf_counter()
    outs = [plug.forward_device(clip_of(k)) for k in mine]
    torch.cuda.synchronize()
    t_clips = time.perf_counter() - t0
    if world > 1:
        dist.barrier()                                         # rank skew belongs to seconds_total, not to stitch_ms
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    stitched = sh.stitch_scene([o["pred_depths"] for o in outs], mine, K, ov, rank, world, engine=eng, device=dev)
    e1.record()
    rows = []
    for k, o in zip(mine, outs):
        d = DM.depth_evaluation(o["pred_depths"], gts[k]["gt_depths"], custom_mask=gts[k]["gt_masks"],
                                align_with_lstsq=True, engine=eng, with_maps=False)[0]
        n = DM.normal_evaluation(o["pred_normals"], gts[k]["gt_normals"], custom_mask=gts[k]["gt_masks"], engine=eng)
        rows.append([float(d[key]) for key in DM.DEPTH_KEYS] + [float(n[key]) for key in DM.NORMAL_KEYS])
    table = sh.gather_metric_rows(mine, torch.tensor(rows, dtype=torch.float64, device=dev), K, rank, world)
    e2.record()
    sync()
    t_total = time.perf_counter() - t0
    t = torch.tensor([t_clips, t_total, e0.elapsed_time(e1), e1.elapsed_time(e2)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tc, tt, stitch_ms, score_ms = t.tolist()
    assert all(torch.isfinite(s_).all() for s_ in stitched)
    return {"workload": f"scene of {K} clips x {T} frames (overlap {ov}, {n_frames} distinct frames) at {a.height}x{a.width}, "
                        f"{a.scene_steps} denoising steps per clip, clips sharded over {world} GPU(s)",
            "scaling": "strong (fixed scene)", "value": K * a.scene_steps / tt, "unit": "steps/s", "clips_per_s": K / tt,
            "seconds_total": tt, "seconds_clips": tc, "stitch_ms": stitch_ms, "score_and_gather_ms": score_ms,
            "collective": ("all_gather of [clips/rank, 2, overlap, H, W] fp32 overlap frames + all_gather of the metric rows (NCCL)"
                           if world > 1 else "none (1 rank)"),
            "stitch": "ug_stitch_fit / ug_stitch_apply (fp64 normal equations in disparity space, chained, ramped)",
            "abs_rel_per_clip": [round(float(v), 5) for v in table[:, 0].tolist()]}


def run_parity(a, eng, cfg, dev):
    """Outside every timed region: the reference's shipped call (5 steps) through the plugin adapter against the fp32
    oracle run on this GPU (torch eager, TF32 off, cuDNN off -- see tests/test_parity_cfg2_gpu.py) on the SAME weights
    (the engine's fp16 values, upcast) and the same noise draws / CLIP embeddings; both scored by the reference metric
    restatement (oracle/metrics.py, bit-exact against /root/reference/metrics/eval_depth.py on the golden fixtures)."""
    import numpy as np
    import torch
    from harness.synthetic import gt_label, make_clip
    from oracle import metrics as OM
    from oracle import postprocess as OP
    from oracle.pipeline import depthcrafter_pipeline
    from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes, vae_param_shapes
    T, H, W = a.frames, a.height, a.width
    if T * H * W > 25 * 384 * 512:
        return {"skipped": "parity block runs at cfg2 size or smaller (the fp32 oracle needs minutes beyond it)"}
    steps = 5
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.enabled)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.enabled = False
    try:
        usd = {k: v.float() for k, v in synthetic_state_dict(unet_param_shapes(cfg.unet), 1000, torch.float16, dev).items()}
        vsd = {k: v.float() for k, v in synthetic_state_dict(vae_param_shapes(cfg.vae), 2000, torch.float16, dev).items()}
        data = make_clip(T, H, W, seed=77)
        plug = _plugin(a, eng, cfg, dev, steps, None)
        g = torch.Generator(device=dev).manual_seed(5)
        enc = torch.randn(T, cfg.clip_embed_dim, generator=g, device=dev)
        aug = torch.randn(T, 3, H, W, generator=g, device=dev)
        init = torch.randn(T, 4, H // 8, W // 8, generator=g, device=dev)
        out = plug.forward(data, enc=enc, aug_noise=aug, init_noise=init)
        frames = torch.from_numpy(OP.prepare_input(data["images"])).to(dev)
        t0 = time.perf_counter()
        with torch.no_grad():
            ref_frames = depthcrafter_pipeline(usd, vsd, cfg, frames, enc[None], aug, init[None], steps).cpu().numpy()
        t_oracle = time.perf_counter() - t0
        ref_depth = torch.from_numpy(np.asarray(OP.disparity_to_depth(ref_frames), dtype=np.float32))
        gt = gt_label(data)
        m_ref = OM.depth_evaluation(ref_depth, gt["gt_depths"], gt["gt_masks"])
        m_got = OM.depth_evaluation(out["pred_depths"], gt["gt_depths"], gt["gt_masks"])
        keys = ("Abs Rel", "delta < 1.25", "delta < 1.25^2", "delta < 1.25^3")
        # second label, built from the oracle arm's own prediction: the scene label cannot tell two randomly initialised
        # arms apart (least-squares alignment -> scale ~ 0 for both), this one can (harness.synthetic.correlated_gt)
        from harness.synthetic import correlated_gt
        cg = correlated_gt(ref_depth)
        s_ref = OM.depth_evaluation(ref_depth, cg["gt_depths"], cg["gt_masks"])
        s_got = OM.depth_evaluation(out["pred_depths"], cg["gt_depths"], cg["gt_masks"])
        sens_ok = bool(abs(s_got["Abs Rel"] - s_ref["Abs Rel"]) <= 1e-3 and
                       all(abs(s_got[k] - s_ref[k]) <= 2e-3 for k in keys[1:]))
        sens = {"label": "oracle depth x (1 + 0.15 smooth field) + 0.1 (prediction-correlated; alignment scale ~ 1)",
                "abs_rel": {"b200": s_got["Abs Rel"], "oracle": s_ref["Abs Rel"],
                            "abs_diff": abs(s_got["Abs Rel"] - s_ref["Abs Rel"])},
                "delta_abs_diff": {k: abs(s_got[k] - s_ref[k]) for k in keys[1:]}, "pass": sens_ok}
        return {"what": f"DepthCrafter.forward, {steps} steps, {T}x{H}x{W}, {a.dtype} kernels vs fp32 oracle on the same GPU",
                "abs_rel": {"b200": m_got["Abs Rel"], "oracle": m_ref["Abs Rel"], "abs_diff": abs(m_got["Abs Rel"] - m_ref["Abs Rel"])},
                "delta_abs_diff": {k: abs(m_got[k] - m_ref[k]) for k in keys[1:]},
                "valid_pixels_equal": m_got["valid_pixels"] == m_ref["valid_pixels"],
                "max_abs_depth_diff": float((out["pred_depths"] - ref_depth).abs().max()),
                "mean_rel_depth_diff": float(((out["pred_depths"] - ref_depth).abs() / ref_depth).mean()),
                "prediction_correlated_label": sens,
                "tolerance": {"abs_rel": 1e-3, "delta": 2e-3}, "oracle_seconds": t_oracle,
                "pass": bool(sens_ok and abs(m_got["Abs Rel"] - m_ref["Abs Rel"]) <= 1e-3 and
                             all(abs(m_got[k] - m_ref[k]) <= 2e-3 for k in keys[1:]))}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.enabled = tf32
        torch.cuda.empty_cache()


def run_library_baseline(a, eng, cfg, dev):
    """The stack the reference actually dispatches to, on this same B200 ("the practical bar to beat"):
    the oracle restatement of one denoising step in torch eager fp16 -- cuDNN convolutions, cuBLASLt linears, SDPA flash
    attention -- on the bench's inputs and weights.  Timed with CUDA events (median of 3 after a warm-up); its output is
    compared with the engine's for the same inputs so that a mis-computing library arm cannot pass as a baseline."""
    import torch
    from oracle import scheduler as S
    from oracle import unet_st as U
    from oracle.pipeline import added_time_ids
    from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes
    T, h, w = a.frames, a.height // 8, a.width // 8
    dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    sd = {k: v.to(dt) for k, v in synthetic_state_dict(unet_param_shapes(cfg.unet), 1000, torch.float16, dev).items()}
    g = torch.Generator(device=dev).manual_seed(4321)
    x = torch.randn(1, T, 8, h, w, generator=g, device=dev)
    enc = torch.randn(1, T, cfg.clip_embed_dim, generator=g, device=dev)
    ids = added_time_ids(cfg, dev)
    U.USE_SDPA = True
    try:
        with torch.no_grad():
            step = lambda: U.unet_forward(sd, cfg.unet, x.to(dt), 0.9, enc.to(dt), ids.to(dt))
            ref = step()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); step(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
    finally:
        U.USE_SDPA = False
    ms = sorted(ts)[1]
    eng.prepare(T, h, w)
    eng.set_clip_context(enc[0])
    ours = eng.unet_forward(x, 0.9, ids[0].tolist())
    rel = ((ours.double() - ref.double()).norm() / ref.double().norm()).item()
    del sd
    torch.cuda.empty_cache()
    return {"value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "dtype": a.dtype,
            "what": "oracle restatement of one UNet denoising step in torch eager (cuDNN / cuBLASLt / SDPA flash), same GPU, "
                    "same weights and inputs; Euler update excluded (negligible)",
            "rel_l2_vs_b200_kernels": rel, "tflops": UNET_TFLOP_PER_STEP.get((a.frames, a.height, a.width), 0) / ms * 1e3}


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
