"""Development aid (gpurun): full-width temporal VAE against the fp32 oracle on the GPU at a ladder of sizes, with
per-frame errors -- finds the size / frame at which a parity break starts."""
import sys

import torch

sys.path.insert(0, ".")
from oracle.vae import vae_decode, vae_encode  # noqa: E402
from unigeo_b200.config import full_config  # noqa: E402
from unigeo_b200.engine import Engine  # noqa: E402
from unigeo_b200.weights import synthetic_state_dict, vae_param_shapes  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda", 0)
cfg = full_config()
vsd = synthetic_state_dict(vae_param_shapes(cfg.vae), 2000, torch.float32, dev)
e = Engine(cfg, dtype=sys.argv[1] if len(sys.argv) > 1 else "fp16", device=0)
e.load_state_dict("vae", vsd)
e.finalize()


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


g = torch.Generator(device=dev).manual_seed(22)
import os  # noqa: E402
SIZES = [(2, 8, 16), (8, 8, 16), (3, 16, 32), (8, 24, 32), (1, 48, 64), (2, 48, 64), (8, 48, 64), (9, 48, 64)]
if os.environ.get("BISECT_SIZES"):
    SIZES = [tuple(int(v) for v in s_.split("x")) for s_ in os.environ["BISECT_SIZES"].split(",")]
for (T, h, w) in SIZES:
    lat = torch.randn(T, 4, h, w, generator=g, device=dev) * 0.5
    with torch.no_grad():
        ref = vae_decode(vsd, cfg.vae, lat, 8)
    got = e.vae_decode(lat, chunk=8)
    torch.cuda.synchronize()
    pf = [rel(got[t], ref[t]) for t in range(T)]
    print(f"decode T{T} {h}x{w}: rel-L2 {rel(got, ref):.3e}  finite {bool(torch.isfinite(got).all())}  ref absmax "
          f"{ref.abs().max().item():.3f} got absmax {got.abs().max().item():.3f}  per-frame {['%.1e' % v for v in pf]}", flush=True)
for (N, H, W) in ([] if os.environ.get("BISECT_SIZES") else [(2, 64, 128), (1, 384, 512), (4, 384, 512)]):
    img = torch.rand(N, 3, H, W, generator=g, device=dev) * 2 - 1
    with torch.no_grad():
        ref = vae_encode(vsd, cfg.vae, img)
    got = e.vae_encode(img)
    torch.cuda.synchronize()
    print(f"encode N{N} {H}x{W}: rel-L2 {rel(got, ref):.3e} finite {bool(torch.isfinite(got).all())} ref absmax {ref.abs().max().item():.3f}", flush=True)
