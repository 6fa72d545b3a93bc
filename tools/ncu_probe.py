"""Runs a few single tapgemm launches (L0 linear shapes) for an ncu capture."""
import math, sys, torch
sys.path.insert(0, ".")
from unigeo_b200 import ops
dev = torch.device("cuda", 0)
M, K, N = 76800, 320, 320
x = (torch.randn(M, K, device=dev)).half()
W = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
b = torch.zeros(N, device=dev)
r = torch.randn(M, N, device=dev).half()
W3 = (torch.randn(960, K, device=dev) / math.sqrt(K)).half()
Wg = (torch.randn(2560, K, device=dev) / math.sqrt(K)).half()
bg = torch.zeros(2560, device=dev)
for _ in range(2):
    ops.linear(x, W, bias=b)             # plain
    ops.linear(x, W, bias=b, res=r)      # +res
    ops.linear(x, W3)                    # qkv
    ops.linear(x, Wg, bias=bg, geglu=True)
torch.cuda.synchronize()
