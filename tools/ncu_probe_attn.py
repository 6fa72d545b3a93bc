"""Runs the L0 spatial attention (fmha_d64, 25 frames x 3072 tokens x 5 heads) and the L0 GroupNorm / LayerNorm launches
a few times for an ncu capture (development aid, gpurun)."""
import sys, torch
sys.path.insert(0, ".")
from unigeo_b200 import ops
dev = torch.device("cuda", 0)
F_, N, C = 25, 3072, 320
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.randn(F_ * N, 3 * C, generator=g, device=dev).half()
x = torch.randn(F_ * N, C, generator=g, device=dev).half()
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(3):
    ops.spatial_attention(qkv, F_, N, C, head_dim=64)
    ops.groupnorm(x, gam, bet, N)              # per-frame sets
    ops.groupnorm(x, gam, bet, F_ * N)         # one set over the clip
    ops.layernorm(x, gam, bet)
torch.cuda.synchronize()
