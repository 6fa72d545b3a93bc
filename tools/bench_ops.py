"""Micro-benchmarks of the hot kernels at cfg2 (25x384x512) layer shapes: CUDA events, L2 flushed
between iterations, prints achieved TFLOP/s / GB/s.  Development aid (gpurun), not the bench."""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
from unigeo_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
dt = torch.bfloat16 if "bf16" in sys.argv else torch.float16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
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


def r(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(dt)


out = []
T = 25
for name, M, K, N in [("L0 qkv", T * 3072, 320, 960), ("L0 ff2", T * 3072, 1280, 320), ("L1 qkv", T * 768, 640, 1920),
                      ("L1 ff2", T * 768, 2560, 640), ("L2 qkv", T * 192, 1280, 3840), ("L2 ff2", T * 192, 5120, 1280),
                      ("big sq", 8192, 8192, 8192)]:
    x, W = r(M, K), r(N, K, scale=1 / math.sqrt(K))
    ms = timeit(lambda: ops.linear(x, W))
    out.append(dict(op="linear " + name, ms=ms, tflops=2 * M * K * N / ms / 1e9))
for name, M, K, Hh in [("L0 geglu", T * 3072, 320, 1280), ("L2 geglu", T * 192, 1280, 5120)]:
    x, W = r(M, K), r(2 * Hh, K, scale=1 / math.sqrt(K))
    b = torch.zeros(2 * Hh, device=dev)
    ms = timeit(lambda: ops.linear(x, W, bias=b, geglu=True))
    out.append(dict(op="linear " + name, ms=ms, tflops=2 * M * K * 2 * Hh / ms / 1e9))
for name, Hh, Ww, C, Co in [("L0 320", 48, 64, 320, 320), ("L1 640", 24, 32, 640, 640), ("L2 1280", 12, 16, 1280, 1280),
                            ("L3 1280", 6, 8, 1280, 1280), ("up3 960->320", 48, 64, 960, 320)]:
    x, W = r(T, Hh, Ww, C), r(9, Co, C, scale=1 / math.sqrt(9 * C))
    ms = timeit(lambda: ops.conv3x3(x, W))
    out.append(dict(op="conv3x3 " + name, ms=ms, tflops=2 * T * Hh * Ww * 9 * C * Co / ms / 1e9))
for name, F, Hh, Ww, C, Co in [("vae 128@384x512 x8", 8, 384, 512, 128, 128), ("vae 256@192x256 x8", 8, 192, 256, 256, 256),
                               ("vae 512@96x128 x8", 8, 96, 128, 512, 512)]:
    x, W = r(F, Hh, Ww, C), r(9, Co, C, scale=1 / math.sqrt(9 * C))
    ms = timeit(lambda: ops.conv3x3(x, W), iters=5)
    out.append(dict(op="conv3x3 " + name, ms=ms, tflops=2 * F * Hh * Ww * 9 * C * Co / ms / 1e9,
                    gbs=(x.numel() + F * Hh * Ww * Co) * 2 / ms / 1e6))
for name, P, C in [("L0", 3072, 320), ("L2", 192, 1280)]:
    x, W = r(T, P, C), r(3, C, C, scale=1 / math.sqrt(3 * C))
    ms = timeit(lambda: ops.tconv3(x, W))
    out.append(dict(op="tconv3 " + name, ms=ms, tflops=2 * T * P * 3 * C * C / ms / 1e9))
for name, N, C in [("L0 N=3072 h5", 3072, 320), ("L1 N=768 h10", 768, 640), ("L2 N=192 h20", 192, 1280)]:
    qkv = r(T * N, 3 * C)
    ms = timeit(lambda: ops.spatial_attention(qkv, T, N, C, 64), iters=5)
    out.append(dict(op="attn " + name, ms=ms, tflops=4 * T * (C // 64) * N * N * 64 / ms / 1e9))
for name, P, C in [("L0", 3072, 320), ("L2", 192, 1280)]:
    qkv = r(T, P, 3 * C)
    ms = timeit(lambda: ops.temporal_attention(qkv, T, P, C))
    out.append(dict(op="tattn " + name, ms=ms, gbs=qkv.numel() * 2 * 4 / 3 / ms / 1e6))
for name, rows, C in [("L0 320", T * 3072, 320), ("vae 128@384x512x8", 8 * 384 * 512, 128)]:
    x = r(rows, C)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    ms = timeit(lambda: ops.groupnorm(x, g, b, rows // (T if rows % T == 0 else 8)))
    out.append(dict(op="groupnorm " + name, ms=ms, gbs=x.numel() * 2 * 3 / ms / 1e6))
    ms = timeit(lambda: ops.layernorm(x, g, b))
    out.append(dict(op="layernorm " + name, ms=ms, gbs=x.numel() * 2 * 2 / ms / 1e6))
for o in out:
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in o.items()}))
