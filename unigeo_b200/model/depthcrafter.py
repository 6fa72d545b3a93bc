"""``DepthCrafter`` plugin adapter on the B200 engine.

Same contract as the reference adapter (/root/reference/model/depthcrafter.py):
  ctor   DepthCrafter(model_dir, unet_path, pre_train_path, **kwargs)     :8
  call   forward(data) -> {'pred_depths' [Nf,H,W], 'pred_normals' [Nf,H,W,3]} CPU float32  :73-99
  input  data['images'] list of [3,H,W] float 0..255 (uint8-truncated, /255)      :39-45
         data['intrinsics'] list of [3,3]                                          :49
Extra, optional ``model_params`` (ride in **kwargs like the reference tolerates):
  config="full"|"tiny", dtype="fp16"|"bf16", num_inference_steps=5, seed=None,
  weights="pretrained" (default) | "synthetic", device=0, clip="random"|"none",
  vae_encode_dtype="bf16" (the VAE encoder alone in bf16: upstream upcasts it to fp32 for range).
Like the reference (from_pretrained at :18-29 raises on a missing checkpoint), a missing ``unet_path`` /
``pre_train_path`` raises FileNotFoundError; seeded random weights load ONLY on an explicit weights="synthetic".
There is no CPU path: constructing this class without a B200 raises.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from ..clip_embed import ClipEmbedder
from ..config import get_config
from ..engine import Engine
from ..pipeline import DepthCrafterPipelineB200
from ..postprocess import depth_and_normals
from ..weights import load_diffusers_dir, synthetic_state_dict, unet_param_shapes, vae_param_shapes


class DepthCrafter:
    def __init__(self, model_dir: Optional[str] = None, unet_path: Optional[str] = None,
                 pre_train_path: Optional[str] = None, **kwargs):
        self.device = torch.device("cuda", int(kwargs.get("device", 0)))
        self.cfg = get_config(kwargs.get("config", "full"))
        self.dtype = kwargs.get("dtype", "fp16")                 # reference: torch_dtype=float16 (:21,:27)
        self.num_inference_steps = int(kwargs.get("num_inference_steps", 5))   # reference hard-codes 5 (:86)
        self.seed = kwargs.get("seed")
        weights = kwargs.get("weights", "pretrained")
        if weights not in ("pretrained", "synthetic"):
            raise ValueError(f"weights must be 'pretrained' or 'synthetic', got {weights!r}")
        if weights == "pretrained":              # never fall back to random weights silently (garbage metrics.csv)
            for what, pth, sub in (("unet_path", unet_path, ""), ("pre_train_path", pre_train_path, "vae")):
                if not pth or not os.path.isdir(os.path.join(pth, sub)):
                    raise FileNotFoundError(f"DepthCrafter: {what}={pth!r} is not a checkpoint directory"
                                            f"{' with a ' + sub + '/ sub-directory' if sub else ''} "
                                            "(pass weights='synthetic' for seeded random weights)")
        self.engine = Engine(self.cfg, dtype=self.dtype, device=self.device.index,
                             vae_encode_dtype=kwargs.get("vae_encode_dtype"))
        self._stage = None                                       # pinned staging buffer of prepare_input_device
        clip_dir = None
        if weights == "pretrained":
            self.engine.load_state_dict("unet", load_diffusers_dir(unet_path))
            self.engine.load_state_dict("vae", load_diffusers_dir(os.path.join(pre_train_path, "vae")))
            clip_dir = os.path.join(pre_train_path, "image_encoder")
        else:
            s = int(kwargs.get("weight_seed", 0))
            # device_weights=True draws the synthetic tensors on the GPU (fast, different values than the CPU draw)
            wdev = self.device if kwargs.get("device_weights") else "cpu"
            wdt = torch.float16 if kwargs.get("device_weights") else torch.float32
            self.engine.load_state_dict("unet", synthetic_state_dict(unet_param_shapes(self.cfg.unet), 1000 + s, wdt, wdev))
            self.engine.load_state_dict("vae", synthetic_state_dict(vae_param_shapes(self.cfg.vae), 2000 + s, wdt, wdev))
        clip = None
        if kwargs.get("clip", "random") != "none":       # CLIP image encoder weights ride in the same context
            clip = ClipEmbedder(self.engine, pretrained=clip_dir if clip_dir and os.path.isdir(clip_dir) else None,
                                device_weights=bool(kwargs.get("device_weights")))
        self.engine.finalize()
        self.pipeline = DepthCrafterPipelineB200(self.cfg, self.engine, clip)
        self.weight_source = (f"pretrained: unet {unet_path}, vae/image_encoder {pre_train_path}" if weights == "pretrained"
                              else f"synthetic (seeded random, weight_seed={int(kwargs.get('weight_seed', 0))})")
        print(f"Using device: {self.device}; weights: {self.weight_source}")

    def prepare_input(self, data):
        """reference :39-45 -- uint8 TRUNCATION (astype), then /255."""
        frames = [np.asarray(x).transpose(1, 2, 0).astype(np.uint8) for x in data["images"]]
        return np.stack(frames, axis=0).astype(np.float32) / 255.0

    def prepare_input_device(self, data) -> torch.Tensor:
        """``prepare_input`` on the device (bit-identical): the raw float images are stacked straight into a
        pinned staging buffer, copied once, truncated / scaled / transposed by one kernel."""
        imgs = data["images"]
        T = len(imgs)
        shape = (T,) + tuple(np.shape(imgs[0]))
        if self._stage is None or tuple(self._stage.shape) != shape:
            self._stage = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        np.stack([np.asarray(x, dtype=np.float32) for x in imgs], axis=0, out=self._stage.numpy())
        return self.engine.prepare_frames(self._stage.to(self.device, non_blocking=True))

    def prepare_output_device(self, frames: torch.Tensor, data):
        """reference :92-97 + :48-69 on the device; the tensors stay there (scoring with
        ``unigeo_b200.metrics`` then moves 19 scalars instead of 2 x 59 MB)."""
        K = torch.from_numpy(np.stack([np.asarray(k, dtype=np.float32) for k in data["intrinsics"]], 0))
        depth, normals = depth_and_normals(self.engine, frames, K)
        return {"pred_depths": depth, "pred_normals": normals}

    def prepare_output(self, frames: torch.Tensor, data):
        """``prepare_output_device`` + the copy to CPU float32 tensors the reference contract asks for (pinned
        pages, so the device->host copy runs at PCIe speed; torch's host allocator recycles them)."""
        dev = self.prepare_output_device(frames, data)
        out = {k: torch.empty(v.shape, dtype=torch.float32, pin_memory=True) for k, v in dev.items()}
        for k, v in dev.items():
            out[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out

    def _run(self, data, **debug_inputs):
        frames = self.prepare_input_device(data)
        gen = None
        if self.seed is not None:
            gen = torch.Generator(device=self.device).manual_seed(int(self.seed))
        return self.pipeline(frames, num_inference_steps=self.num_inference_steps, generator=gen,
                             output_type="pt", **debug_inputs)

    def forward(self, data, **debug_inputs):
        """``debug_inputs``: enc / aug_noise / init_noise tensors to pin the random draws (parity tests)."""
        return self.prepare_output(self._run(data, **debug_inputs), data)

    def forward_device(self, data, **debug_inputs):
        """``forward`` without the final device->host copy: same values, CUDA tensors."""
        return self.prepare_output_device(self._run(data, **debug_inputs), data)
