"""BASELINE cfg1 on the FULL architecture: one 8-frame 256x256 clip, 5 denoising steps, through the eval.py-shaped loop
(adapter forward -> depth_evaluation -> normal_evaluation), once on the host cores (the fp32 oracle restatement of the
pipeline the reference calls at /root/reference/model/depthcrafter.py:80-90) and once through the plugin
on the B200, same seeded weights / frames / noise draws / image embeddings, and the two scored side by side.

    python tools/run_cfg1.py [--out profiles/r02_cfg1_full_arch.json] [--skip-b200]

Development / reporting aid (run under gpurun); the CPU arm alone also runs in the build container."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/cfg1_full_arch.json")
    ap.add_argument("--skip-b200", action="store_true")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    from harness.synthetic import correlated_gt, gt_label, make_clip
    from oracle import metrics as OM
    from oracle import postprocess as OP
    from oracle.pipeline import depthcrafter_pipeline
    from unigeo_b200.config import full_config
    from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes, vae_param_shapes
    cfg = full_config()
    T, H, W = a.frames, a.size, a.size
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    data = make_clip(T, H, W, seed=1)
    g = torch.Generator().manual_seed(9)
    enc = torch.randn(T, cfg.clip_embed_dim, generator=g)
    aug = torch.randn(T, 3, H, W, generator=g)
    init = torch.randn(T, 4, H // 8, W // 8, generator=g)
    # fp16-representable weights, so that both arms hold exactly the same values
    usd16 = synthetic_state_dict(unet_param_shapes(cfg.unet), 1000, torch.float16)
    vsd16 = synthetic_state_dict(vae_param_shapes(cfg.vae), 2000, torch.float16)
    usd = {k: v.float() for k, v in usd16.items()}
    vsd = {k: v.float() for k, v in vsd16.items()}
    gt = gt_label(data)
    rep = {"config": f"DepthCrafter {T}x{H}x{W}, {a.steps} denoising steps, full SVD-XT architecture (1.52 B + 97 M params), "
                     "seeded random-init weights, synthetic scene", "host_cores": cores}

    # ---- CPU arm: the eval.py-shaped loop on the host
    t0 = time.perf_counter()
    frames = torch.from_numpy(OP.prepare_input(data["images"]))
    with torch.no_grad():
        ref_frames = depthcrafter_pipeline(usd, vsd, cfg, frames, enc[None], aug, init[None], a.steps).numpy()
    t_pipe = time.perf_counter() - t0
    ref_depth = np.asarray(OP.disparity_to_depth(ref_frames), dtype=np.float32)
    t1 = time.perf_counter()
    ref_out = OP.prepare_output(ref_depth, data["intrinsics"])
    t_post = time.perf_counter() - t1
    t2 = time.perf_counter()
    ref_d = torch.from_numpy(ref_depth)
    m_ref = OM.depth_evaluation(ref_d, gt["gt_depths"], gt["gt_masks"])
    n_ref = OM.normal_evaluation(ref_out["pred_normals"], gt["gt_normals"], gt["gt_masks"])
    t_metric = time.perf_counter() - t2
    rep["cpu"] = {"what": "oracle restatement (torch fp32, all host threads): VAE encode + 5 UNet steps + temporal VAE decode, "
                          "depth / plane-fit-normal post-processing, depth_evaluation + normal_evaluation",
                  "seconds_pipeline": t_pipe, "seconds_postprocess": t_post, "seconds_metrics": t_metric,
                  "steps_per_s": a.steps / t_pipe,
                  "Abs Rel": m_ref["Abs Rel"], "delta < 1.25": m_ref["delta < 1.25"], "normal mean": n_ref["normal mean"]}
    print("cpu arm", json.dumps(rep["cpu"]), flush=True)
    cg = correlated_gt(ref_d)
    s_ref = OM.depth_evaluation(ref_d, cg["gt_depths"], cg["gt_masks"])

    if not a.skip_b200 and torch.cuda.is_available():
        from unigeo_b200.engine import Engine
        from unigeo_b200.model.depthcrafter import DepthCrafter
        from unigeo_b200.pipeline import DepthCrafterPipelineB200
        dev = torch.device("cuda", 0)
        eng = Engine(cfg, dtype="fp16", device=0)
        eng.load_state_dict("unet", {k: v.to(dev) for k, v in usd16.items()})
        eng.load_state_dict("vae", {k: v.to(dev) for k, v in vsd16.items()})
        eng.finalize()
        plug = object.__new__(DepthCrafter)
        plug.device, plug.cfg, plug.dtype, plug.engine = dev, cfg, "fp16", eng
        plug.num_inference_steps, plug.seed, plug._stage = a.steps, None, None
        plug.pipeline = DepthCrafterPipelineB200(cfg, eng, None)
        kw = dict(enc=enc.to(dev), aug_noise=aug.to(dev), init_noise=init.to(dev))
        out = plug.forward(data, **kw)                         # first call: workspace sizing, graph capture
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = plug.forward(data, **kw)
        torch.cuda.synchronize()
        t_gpu = time.perf_counter() - t0
        m_got = OM.depth_evaluation(out["pred_depths"], gt["gt_depths"], gt["gt_masks"])
        n_got = OM.normal_evaluation(out["pred_normals"], gt["gt_normals"], gt["gt_masks"])
        s_got = OM.depth_evaluation(out["pred_depths"], cg["gt_depths"], cg["gt_masks"])
        rep["b200"] = {"what": "unigeo_b200.model.DepthCrafter.forward(data), fp16 kernels, host buffers in and out",
                       "seconds_forward": t_gpu, "steps_per_s": a.steps / t_gpu,
                       "Abs Rel": m_got["Abs Rel"], "delta < 1.25": m_got["delta < 1.25"], "normal mean": n_got["normal mean"]}
        rep["parity"] = {"abs_rel_diff": abs(m_got["Abs Rel"] - m_ref["Abs Rel"]),
                         "delta125_diff": abs(m_got["delta < 1.25"] - m_ref["delta < 1.25"]),
                         "normal_mean_diff_deg": abs(n_got["normal mean"] - n_ref["normal mean"]),
                         "max_abs_depth_diff": float((out["pred_depths"] - ref_d).abs().max()),
                         "prediction_correlated_label": {"abs_rel_oracle": s_ref["Abs Rel"], "abs_rel_b200": s_got["Abs Rel"],
                                                         "abs_diff": abs(s_got["Abs Rel"] - s_ref["Abs Rel"])},
                         "tolerance": {"abs_rel": 1e-3, "delta": 2e-3, "normal_mean_deg": 0.1}}
        rep["speedup_forward_over_cpu_pipeline"] = t_pipe / t_gpu
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
