"""SM clock / board power / throttle reasons sampled every 50 ms while the cfg2 denoising loop runs (development aid,
gpurun): the evidence behind "every sample sits under the software power cap" in DESIGN.md 8.
    python tools/power_trace.py [steps]"""
import subprocess, sys, time, torch
sys.path.insert(0, ".")
from unigeo_b200.config import get_config
from unigeo_b200.engine import Engine
from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
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
for _ in range(2):
    eng.denoise(cond, noise, ids, steps)           # eager, then captured
torch.cuda.synchronize()
time.sleep(2.0)                                    # let the board idle: the trace starts from rest
mon = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,power.draw,power.limit,clocks_throttle_reasons.active,temperature.gpu",
                        "--format=csv,noheader", "-lms", "50"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
ev0.record()
eng.denoise(cond, noise, ids, steps)
ev1.record()
torch.cuda.synchronize()
t1 = time.time()
time.sleep(0.3)
mon.terminate()
lines = mon.stdout.read().strip().splitlines()
print(f"{steps} steps in {ev0.elapsed_time(ev1):.1f} ms = {ev0.elapsed_time(ev1) / steps:.3f} ms per step (wall {t1 - t0:.2f} s)")
print("timestamp, sm MHz, power W, limit W, throttle reasons (bitmask: 0x4 = sw power cap), temp C")
for ln in lines:
    print(ln)
