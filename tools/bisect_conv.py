"""Development aid (gpurun): error map of one conv3x3 shape against torch."""
import math
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_ops_gpu import conv_ref, rel_l2, rnd  # noqa: E402
from unigeo_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
F_, H, W, C, Co = (int(v) for v in sys.argv[1:6])
dt = torch.float16
torch.backends.cudnn.allow_tf32 = False
x = rnd((F_, H, W, C), dt, dev, 1)
Wt = rnd((9, Co, C), dt, dev, 2, 1 / math.sqrt(9 * C))
b = rnd((Co,), torch.float32, dev, 3)
ref = conv_ref(x, Wt, b, 1, 0)
got = ops.conv3x3(x, Wt, bias=b).float()
got2 = ops.conv3x3(x, Wt, bias=b).float()
print(f"conv [{F_},{H},{W},{C}]->{Co}: rel-L2 {rel_l2(got, ref):.3e}; rerun identical {bool(torch.equal(got, got2))}")
err = (got - ref).abs()
bad = err > 0.05
print("bad fraction", bad.float().mean().item())
pf = bad.float().mean(dim=(1, 2, 3))
print("bad fraction per frame", [round(v, 4) for v in pf.tolist()])
rows = bad.float().mean(dim=(2, 3))            # [F, H]
for f in range(F_):
    r = rows[f]
    idx = (r > 0).nonzero().flatten().tolist()
    if idx:
        print(f" frame {f}: bad rows {idx[:40]}{'...' if len(idx) > 40 else ''} (n={len(idx)})")
cols = bad.float().mean(dim=(0, 1, 3))
print("bad x positions", (cols > 0).nonzero().flatten().tolist()[:64])
ch = bad.float().mean(dim=(0, 1, 2))
ci = (ch > 0).nonzero().flatten().tolist()
print("bad channels", ci[:16], "...", ci[-16:], "n=", len(ci))
if bad.any():
    i = bad.nonzero()[0].tolist()
    print("first bad", i, "got", got[tuple(i)].item(), "ref", ref[tuple(i)].item())
    f, y = i[0], i[1]
    print("row sample got", got[f, y, :4, :4].tolist(), "ref", ref[f, y, :4, :4].tolist())
# independent fp64 reference of frame 0, row 1 (unfold + matmul), all channels
xp = torch.nn.functional.pad(x[0].double(), (0, 0, 1, 1, 1, 1))          # [H+2, W+2, C]
yy = 1
acc = torch.zeros(W, Co, dtype=torch.float64, device=dev)
for ky in range(3):
    for kx in range(3):
        acc += xp[yy + ky, kx:kx + W, :] @ Wt[ky * 3 + kx].double().t()
acc += b.double()
print("fp64 check, frame 0 row 1: |ours - fp64| max", (got[0, yy].double() - acc).abs().max().item(),
      " |torch ref - fp64| max", (ref[0, yy].double() - acc).abs().max().item())
print(" at ch 254: ours", got[0, yy, :3, 254].tolist(), "torch", ref[0, yy, :3, 254].tolist(), "fp64", acc[:3, 254].tolist())
