"""A/B of the balanced tile lists of tapgemm (UG_SCHED, read once per process): for the cfg2 layer shapes whose last
N tile is ragged, prints per shape the median time (CUDA events, L2 flushed between iterations) and a SHA-256 of the
output bytes.  Run once with UG_SCHED=0 and once with UG_SCHED=1: the hashes must be identical (tiles are computed
independently of which CTA runs them), the times show the gain.  `python tools/ab_sched.py [--out file.json]`."""
import hashlib
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unigeo_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
dt = torch.float16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
quick = "--quick" in sys.argv


def timeit(fn, iters=15, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def r(*s, scale=1.0, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    return (torch.randn(*s, device=dev, generator=g) * scale).to(dt)


def digest(t):
    return hashlib.sha256(t.cpu().contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()[:16]


rows = []
T = 25
lin = [("L1 to_out +res", T * 768, 640, 640, True), ("L1 ff2 +res", T * 768, 2560, 640, True),
       ("L0 to_out +res", T * 3072, 320, 320, True), ("L0 ff2 +res", T * 3072, 1280, 320, True),
       ("L0 qkv", T * 3072, 320, 960, False), ("L1 qkv", T * 768, 640, 1920, False)]
if quick:
    lin = lin[:3]
for name, M, K, N, with_res in lin:
    x, W = r(M, K, seed=1), r(N, K, scale=1 / math.sqrt(K), seed=2)
    res = r(M, N, seed=3) if with_res else None
    b = torch.linspace(-1, 1, N, device=dev)
    y = ops.linear(x, W, bias=b, res=res)
    rows.append(dict(op=f"linear {name} M{M} N{N} K{K}", sha=digest(y),
                     us=None if quick else 1e3 * timeit(lambda: ops.linear(x, W, bias=b, res=res))))
conv = [("L0 320", 48, 64, 320, 320), ("L1 640", 24, 32, 640, 640)]
for name, Hh, Ww, C, Co in conv:
    x, W = r(T, Hh, Ww, C, seed=4), r(9, Co, C, scale=1 / math.sqrt(9 * C), seed=5)
    y = ops.conv3x3(x, W)
    rows.append(dict(op=f"conv3x3 {name}", sha=digest(y), us=None if quick else 1e3 * timeit(lambda: ops.conv3x3(x, W))))
x, W = r(T, 768, 640, seed=6), r(3, 640, 640, scale=1 / math.sqrt(3 * 640), seed=7)
y = ops.tconv3(x, W)
rows.append(dict(op="tconv3 L1 640", sha=digest(y), us=None if quick else 1e3 * timeit(lambda: ops.tconv3(x, W))))
out = {"UG_SCHED": os.environ.get("UG_SCHED", "(default on)"), "rows": rows}
line = json.dumps(out)
print(line)
for a in sys.argv:
    if a.startswith("--out="):
        os.makedirs(os.path.dirname(a[6:]) or ".", exist_ok=True)
        open(a[6:], "w").write(line + "\n")
