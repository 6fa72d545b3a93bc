"""Where do the joules of a power-capped denoising step go?  Runs ONE op of the cfg2 step (L0 shapes, fp16) back to back
for ~1.2 s, samples nvidia-smi (board power, SM clock) every 20 ms, and reports per launch: time (CUDA events over the
loop), mean board power and SM clock of the loaded window, and the dynamic energy (power above the idle board) --
next to the op's algorithmic flops, so that pJ per flop of the tensor-bound ops can be compared with cuBLAS.
    python tools/energy_by_kernel.py            (gpurun; development / reporting aid)"""
import json
import math
import subprocess
import sys
import time

import torch

sys.path.insert(0, ".")
from unigeo_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
T, h, w, C = 25, 48, 64, 320
M = T * h * w


def rnd(*s, sc=1.0, dt=torch.float16):
    return (torch.randn(*s, generator=g, device=dev) * sc).to(dt)


x = rnd(M, C)
x4 = rnd(M, 4 * C)
xc = rnd(T, h, w, C)
res = rnd(M, C)
W_sq = rnd(C, C, sc=1 / math.sqrt(C))
W_qkv = rnd(3 * C, C, sc=1 / math.sqrt(C))
W_ff2 = rnd(C, 4 * C, sc=1 / math.sqrt(4 * C))
Wg, bg = ops.geglu_interleave(rnd(8 * C, C, sc=1 / math.sqrt(C)), torch.zeros(8 * C, device=dev))
W_conv = rnd(9, C, C, sc=1 / math.sqrt(9 * C))
b = torch.zeros(C, device=dev)
qkv = rnd(M, 3 * C)
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
A8k = rnd(8192, 8192, dt=torch.bfloat16)
B8k = rnd(8192, 8192, dt=torch.bfloat16)

CASES = [
    ("cublas bf16 8192^3 (the peak's own workload)", lambda: torch.matmul(A8k, B8k), 2 * 8192 ** 3),
    ("conv3x3 L0 320->320 +res", lambda: ops.conv3x3(xc, W_conv, bias=b, res=xc), 2 * M * C * C * 9),
    ("linear_geglu L0 K320 N2560", lambda: ops.linear(x, Wg, bias=bg, geglu=True), 2 * M * C * 8 * C),
    ("linear qkv L0 K320 N960", lambda: ops.linear(x, W_qkv), 2 * M * C * 3 * C),
    ("linear L0 K320 N320 +res", lambda: ops.linear(x, W_sq, bias=b, res=res), 2 * M * C * C),
    ("linear L0 K1280 N320 +res", lambda: ops.linear(x4, W_ff2, bias=b, res=res), 2 * M * C * 4 * C),
    ("fmha_d64 L0 25x3072 tokens, 5 heads", lambda: ops.spatial_attention(qkv, T, h * w, C, head_dim=64), 4 * T * (h * w) ** 2 * C),
    ("groupnorm L0 per frame (+SiLU)", lambda: ops.groupnorm(x, gam, bet, h * w), 0),
    ("layernorm L0", lambda: ops.layernorm(x, gam, bet), 0),
    ("temporal_attention L0", lambda: ops.temporal_attention(qkv.view(T, h * w, 3 * C), T, h * w, C), 4 * T * T * 64 * (h * w) * (C // 64)),
]


def sample(fn, seconds=1.2):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    n = max(20, int(seconds * 1e3 / (e0.elapsed_time(e1) / 20)))
    time.sleep(1.0)
    mon = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits", "-lms", "20"],
                           stdout=subprocess.PIPE, text=True)
    time.sleep(0.25)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    time.sleep(0.1)
    mon.terminate()
    rows = [tuple(float(v) for v in ln.split(",")) for ln in mon.stdout.read().strip().splitlines() if "," in ln]
    idle = min(r[0] for r in rows)
    peak = max(r[0] for r in rows)
    loaded = [r for r in rows if r[0] > idle + 0.5 * (peak - idle)]
    loaded = loaded[len(loaded) // 4:] or loaded                       # drop the ramp
    p = sum(r[0] for r in loaded) / len(loaded)
    clk = sum(r[1] for r in loaded) / len(loaded)
    return ms, p, clk, idle, n


out = []
for name, fn, flops in CASES:
    ms, p, clk, idle, n = sample(fn)
    dyn_mj = (p - idle) * ms                                        # W * ms = mJ
    row = {"op": name, "us_per_launch": round(ms * 1e3, 1), "launches": n, "board_w": round(p, 1), "idle_w": round(idle, 1),
           "sm_mhz": round(clk), "dynamic_mj_per_launch": round(dyn_mj, 2),
           "tflops": round(flops / ms / 1e9, 1) if flops else None,
           "pj_per_flop_dynamic": round(dyn_mj * 1e9 / flops, 3) if flops else None}
    out.append(row)
    print(json.dumps(row), flush=True)
json.dump(out, open("gpurun_out/energy_by_kernel.json", "w"), indent=1)
