"""A/B of the cluster GroupNorm (UG_GN_CLUSTER, read once per process) at the UNet's per-frame GroupNorm shapes:
median time (CUDA events, L2 flushed) and max-abs error against torch's fp32 group_norm(+SiLU) per shape.
`UG_GN_CLUSTER=0 python tools/ab_gn.py; UG_GN_CLUSTER=1 python tools/ab_gn.py [--out=file.json]`."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unigeo_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


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


rows_out = []
shapes = [(25, 48, 1280, 0), (25, 48, 1280, 1280), (25, 192, 1280, 0), (25, 192, 1280, 1280), (25, 768, 640, 0),
          (25, 768, 320, 640), (25, 768, 640, 640), (25, 3072, 320, 0), (1, 1200, 1280, 0)]
for sets, rps, C1, C2 in shapes:
    g = torch.Generator(device=dev).manual_seed(sets * 1000 + rps + C1 + C2)
    rows, C = sets * rps, C1 + C2
    x1 = (torch.randn(rows, C1, device=dev, generator=g) * 2 + 0.5).half()
    x2 = torch.randn(rows, C2, device=dev, generator=g).half() if C2 else None
    ga = torch.randn(C, device=dev, generator=g) * 0.1 + 1
    be = torch.randn(C, device=dev, generator=g) * 0.1
    xc = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], 1)
    ref = F.silu(F.group_norm(xc.view(sets, rps, C).permute(0, 2, 1), 32, ga, be, 1e-5).permute(0, 2, 1).reshape(rows, C))
    y = ops.groupnorm(x1, ga, be, rps, 32, 1e-5, True, x2)
    y2 = ops.groupnorm(x1, ga, be, rps, 32, 1e-5, True, x2)
    rows_out.append(dict(shape=f"sets{sets} rows/set{rps} C{C1}+{C2}", max_abs_err=float((y.float() - ref).abs().max()),
                         rerun_identical=bool(torch.equal(y, y2)),
                         us=1e3 * timeit(lambda: ops.groupnorm(x1, ga, be, rps, 32, 1e-5, True, x2))))
out = {"UG_GN_CLUSTER": os.environ.get("UG_GN_CLUSTER", "(default on)"), "rows": rows_out}
line = json.dumps(out)
print(line)
for a in sys.argv:
    if a.startswith("--out="):
        os.makedirs(os.path.dirname(a[6:]) or ".", exist_ok=True)
        open(a[6:], "w").write(line + "\n")
