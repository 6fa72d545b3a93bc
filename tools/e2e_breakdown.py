"""Wall-clock breakdown of DepthCrafter.forward(data) at cfg2 (development aid, gpurun)."""
import sys, time, torch
sys.path.insert(0, ".")
from harness.synthetic import make_clip
from unigeo_b200.model import DepthCrafter
import unigeo_b200.pipeline as P
import unigeo_b200.postprocess as PP

T, H, W = 25, 384, 512
plug = DepthCrafter(config="full", dtype="fp16", weights="synthetic", num_inference_steps=25, seed=1, device_weights=True)
data = make_clip(T, H, W, seed=1)
for _ in range(2):                 # workspace sizing, then the CUDA-graph capture of the denoising loop
    plug.forward(data)
torch.cuda.synchronize()
eng = plug.engine
marks = {}
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); marks[label] = marks.get(label, 0) + time.perf_counter() - t
        return r
    setattr(obj, name, g)
wrap(plug.pipeline, "clip", "clip")
wrap(eng, "vae_encode_frames", "vae_encode"); wrap(eng, "denoise", "denoise"); wrap(eng, "vae_decode_frames", "vae_decode")
wrap(eng, "depth_postprocess", "  of which depth_postprocess kernel")
wrap(eng, "set_clip_context", "clip_context")
wrap(plug, "prepare_input_device", "prepare_input(stack + H2D + kernel)"); wrap(plug, "prepare_output", "prepare_output(gpu post + D2H)")
t0 = time.perf_counter(); plug.forward(data); torch.cuda.synchronize(); tot = time.perf_counter() - t0
for k, v in marks.items(): print(f"{k:34s} {v*1e3:8.1f} ms")
print(f"{'total':34s} {tot*1e3:8.1f} ms  (unaccounted {1e3*(tot-sum(marks.values())):.1f})")
