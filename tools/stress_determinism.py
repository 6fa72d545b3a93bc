"""Run-to-run determinism stress of one denoising step (development aid, gpurun): N calls of ug_denoise_clip(steps=1) on
fixed inputs, every result compared bit for bit with the first.
    python tools/stress_determinism.py [T h w [dtype]]        DBG_RUNS=200 DBG_NO_PROFILE=1 UG_NO_PDL_K=fmha,...
This is the tool that found (and bisected, with UG_NO_PDL_K and no-trigger build variants) the programmatic-dependent-
launch race recorded in profiles/r02_pdl_race.txt."""
import os, sys, torch
sys.path.insert(0, ".")
from unigeo_b200.config import full_config
from unigeo_b200.engine import Engine
from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes
T, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (49, 72, 128)))
dt = sys.argv[4] if len(sys.argv) > 4 else "bf16"
cfg = full_config()
e = Engine(cfg, dtype=dt, device=0)
e.load_state_dict("unet", synthetic_state_dict(unet_param_shapes(cfg.unet), 1000, torch.float16, e.device))
e.finalize()
g = torch.Generator(device="cuda").manual_seed(3)
cond = torch.randn(T, 4, h, w, generator=g, device="cuda"); noise = torch.randn(T, 4, h, w, generator=g, device="cuda")
e.prepare(T, h, w)
e.set_clip_context(torch.randn(T, cfg.clip_embed_dim, generator=g, device="cuda"))
ids = [cfg.fps_id, cfg.motion_bucket_id, cfg.noise_aug_strength]
N = int(os.environ.get("DBG_RUNS", "40"))
ref = e.denoise(cond, noise, ids, 1).clone()
bad = []
for i in range(N):
    o = e.denoise(cond, noise, ids, 1)
    if not torch.equal(ref, o):
        bad.append((i, float((ref - o).abs().max())))
torch.cuda.synchronize()
print("env", {k: v for k, v in os.environ.items() if k.startswith("UG_")}, f"{len(bad)} of {N} runs differ from run 0", bad[:6], flush=True)
if os.environ.get("DBG_NO_PROFILE"):
    sys.exit(0)
e.profile(True, by_shape=True)
e.denoise(cond, noise, ids, 1)
for r in e.profile_read(cap=1024):
    if "splitk" in r["name"]:
        print("  ", r["launches"], round(1e3 * r["ms"] / r["launches"], 1), "us", r["name"])
