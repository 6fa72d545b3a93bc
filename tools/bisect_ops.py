"""Development aid (gpurun): op-level parity at the exact shapes of a failing VAE decode (F frames of h x w latents)."""
import math
import sys

import torch
import torch.nn.functional as Fn

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_ops_gpu import conv_ref, rel_l2, rnd, tconv_ref  # noqa: E402
from unigeo_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
F_, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8, 24, 32)))
dt = torch.float16
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def show(name, got, ref):
    print(f"{name:60s} rel-L2 {rel_l2(got, ref):.3e} finite {bool(torch.isfinite(got.float()).all())}", flush=True)


for lvl, (C, Co) in enumerate([(512, 512), (512, 512), (512, 256), (256, 128)]):
    H, W = h << lvl, w << lvl
    P = H * W
    x = rnd((F_, H, W, C), dt, dev, 1)
    Wt = rnd((9, Co, C), dt, dev, 2, 1 / math.sqrt(9 * C))
    b = rnd((Co,), torch.float32, dev, 3)
    show(f"conv3x3 [{F_},{H},{W},{C}]->{Co}", ops.conv3x3(x, Wt, bias=b), conv_ref(x, Wt, b, 1, 0))
    r = rnd((F_, H, W, Co), dt, dev, 4)
    show(f"conv3x3+res [{F_},{H},{W},{C}]->{Co}", ops.conv3x3(x, Wt, bias=b, res=r), conv_ref(x, Wt, b, 1, 0) + r.float())
    xt = rnd((F_, P, Co), dt, dev, 5)
    Wtt = rnd((3, Co, Co), dt, dev, 6, 1 / math.sqrt(3 * Co))
    bt = rnd((Co,), torch.float32, dev, 7)
    ref = tconv_ref(xt, Wtt, bt, F_)
    show(f"tconv3 [{F_},{P},{Co}]", ops.tconv3(xt, Wtt, bias=bt, chunk=F_), ref)
    s = rnd((F_, P, Co), dt, dev, 8)
    show(f"tconv3+res+blend [{F_},{P},{Co}]", ops.tconv3(xt, Wtt, bias=bt, res=s, blend=s, alpha=0.3, chunk=F_),
         0.3 * s.float() + 0.7 * (ref + s.float()))
    gam, bet = rnd((Co,), torch.float32, dev, 9) * 0.1 + 1, rnd((Co,), torch.float32, dev, 10) * 0.1
    x2 = xt.reshape(F_ * P, Co)
    for rps, nm in ((P, "per-frame"), (F_ * P, "whole-chunk")):
        xr = x2.float().reshape(F_ * P // rps, rps, Co).permute(0, 2, 1)
        ref_g = Fn.silu(Fn.group_norm(xr, 32, gam, bet, 1e-6)).permute(0, 2, 1).reshape(F_ * P, Co)
        show(f"groupnorm rows {F_ * P} C {Co} {nm}", ops.groupnorm(x2, gam, bet, rps, eps=1e-6), ref_g)
    if Co != C:
        Ws = rnd((Co, C), dt, dev, 11, 1 / math.sqrt(C))
        show(f"linear shortcut M {F_ * P} {C}->{Co}", ops.linear(x.reshape(-1, C), Ws, bias=b), x.reshape(-1, C).float() @ Ws.float().t() + b)
    if lvl == 0:
        N = P
        qkv = rnd((F_ * N, 3 * C), dt, dev, 12)
        q, k, v = (t.float().reshape(F_, N, C) for t in qkv.split(C, dim=1))
        ref_a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), -1) @ v
        show(f"spatial_attention F{F_} N{N} d{C}", ops.spatial_attention(qkv, F_, N, C, head_dim=C), ref_a.reshape(F_ * N, C))
    if lvl < 3:
        # nearest x2 upsample + conv (the decoder's up path)
        pass
