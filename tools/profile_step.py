"""Per-shape kernel table of one denoising step at cfg2 (development aid, run under gpurun)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from unigeo_b200.config import get_config  # noqa: E402
from unigeo_b200.engine import Engine  # noqa: E402
from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes  # noqa: E402

cfg = get_config("full")
T, h, w = 25, 48, 64
eng = Engine(cfg, dtype="fp16", device=0)
eng.load_state_dict("unet", synthetic_state_dict(unet_param_shapes(cfg.unet), 1000, torch.float16))
eng.finalize()
eng.prepare(T, h, w)
g = torch.Generator().manual_seed(0)
cond, noise = torch.randn(T, 4, h, w, generator=g).cuda(), torch.randn(T, 4, h, w, generator=g).cuda()
eng.set_clip_context(torch.randn(T, 1024, generator=g).cuda())
ids = [7.0, 127.0, 0.02]
eng.denoise(cond, noise, ids, 2)
torch.cuda.synchronize()
eng.profile(True, by_shape=True)
eng.denoise(cond, noise, ids, 1)
rows = sorted(eng.profile_read(), key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in rows)
print(f"total {tot:.2f} ms")
for r in rows[:int(sys.argv[1]) if len(sys.argv) > 1 else 60]:
    tf = r["flops"] / r["ms"] / 1e9 if r["ms"] else 0
    print(f"{r['ms']:7.3f} ms  n={r['launches']:3d}  {r['ms']/r['launches']*1e3:7.1f} us  {tf:7.0f} TF/s  {r['name']}")
json.dump(rows, open("gpurun_out/step_by_shape.json", "w"), indent=1)
