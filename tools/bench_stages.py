"""Stage timings through the C ABI (CUDA events): VAE encode / decode at cfg2, one denoising step at cfg5.
Development aid for gpurun; prints JSON lines."""
import json
import sys

import torch

sys.path.insert(0, ".")
from unigeo_b200.config import get_config  # noqa: E402
from unigeo_b200.engine import Engine  # noqa: E402
from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes, vae_param_shapes  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
cfg = get_config("full")
dev = torch.device("cuda", 0)


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


if which in ("all", "vae"):
    eng = Engine(cfg, dtype="fp16")
    eng.load_state_dict("vae", synthetic_state_dict(vae_param_shapes(cfg.vae), 2000, torch.float16, dev))
    eng.finalize()
    T, H, W = 25, 384, 512
    img = torch.rand(T, 3, H, W, device=dev) * 2 - 1
    lat = torch.randn(T, 4, H // 8, W // 8, device=dev) * 0.5
    eng.vae_encode(img)
    eng.profile(True, by_shape=True)
    timed(lambda: eng.vae_encode(img), iters=1)
    erows = sorted(eng.profile_read(), key=lambda r: -r["ms"])
    eng.profile(False)
    ms = timed(lambda: eng.vae_encode(img))
    print(json.dumps({"stage": "vae_encode 25x384x512", "ms": ms, "tflops": 20.8 / ms * 1e3, "gbs_alg": 23.3 / ms * 1e3}))
    for r in erows[:40]:
        tf = r["flops"] / (r["ms"] / 2) / 1e9 if r["ms"] else 0
        gb = r["bytes"] / (r["ms"] / 2) / 1e6 if r["ms"] else 0
        print(f"   {r['ms']/2:8.3f} ms n={r['launches']//2:3d} {tf/2:7.0f} TF/s {gb/2:7.0f} GB/s  {r['name']}")
    json.dump(erows, open("gpurun_out/vae_encode_by_shape.json", "w"), indent=1)
    eng.profile(True, by_shape=True)
    ms = timed(lambda: eng.vae_decode(lat, 8), iters=1)
    rows = sorted(eng.profile_read(), key=lambda r: -r["ms"])
    eng.profile(False)
    ms = timed(lambda: eng.vae_decode(lat, 8))
    print(json.dumps({"stage": "vae_decode 25x384x512 chunk 8", "ms": ms, "tflops": 56.9 / ms * 1e3, "gbs_alg": 68.5 / ms * 1e3,
                      "workspace_gb": eng.workspace_bytes() / 2 ** 30}))
    tot = sum(r["ms"] for r in rows) / 2
    for r in rows[:60]:
        tf = r["flops"] / (r["ms"] / 2) / 1e9 if r["ms"] else 0
        gb = r["bytes"] / (r["ms"] / 2) / 1e6 if r["ms"] else 0
        print(f"   {r['ms']/2:8.3f} ms n={r['launches']//2:3d} {tf/2:7.0f} TF/s {gb/2:7.0f} GB/s  {r['name']}")
    json.dump(rows, open("gpurun_out/vae_decode_by_shape.json", "w"), indent=1)
    eng.close()
    del eng
    torch.cuda.empty_cache()

if which in ("all", "cfg5"):
    eng = Engine(cfg, dtype="bf16")
    eng.load_state_dict("unet", synthetic_state_dict(unet_param_shapes(cfg.unet), 1000, torch.float16, dev))
    eng.finalize()
    T, h, w = 49, 72, 128
    eng.prepare(T, h, w)
    eng.set_clip_context(torch.randn(T, 1024, device=dev))
    cond, noise = torch.randn(T, 4, h, w, device=dev), torch.randn(T, 4, h, w, device=dev)
    ids = [7.0, 127.0, 0.02]
    ms = timed(lambda: eng.denoise(cond, noise, ids, 2), iters=2) / 2
    out = eng.denoise(cond, noise, ids, 2)
    print(json.dumps({"stage": "denoise step 49x576x1024 bf16 (cfg5)", "ms_per_step": ms, "steps_per_s": 1e3 / ms,
                      "tflops": 156.6 / ms * 1e3, "finite": bool(torch.isfinite(out).all()),
                      "workspace_gb": eng.workspace_bytes() / 2 ** 30}))
