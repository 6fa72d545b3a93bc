"""DepthCrafter pipeline over the B200 engine: what ``self.pipeline(frames, ...)`` does at
/root/reference/model/depthcrafter.py:80-90 with the argument set UniGeo fixes
(guidance_scale=1.0, window_size=len(frames) => one window, output_type="np").

Host side only moves data: CLIP embeddings (torch library code), then three C-ABI calls --
``ug_vae_encode_frames`` -> ``ug_denoise_clip`` -> ``ug_vae_decode_frames`` -- on the current stream
(x*2-1, the noise augmentation and postprocess_video's clamp / layout ride in the layout kernels).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .config import PipelineConfig
from .engine import Engine


class DepthCrafterPipelineB200:
    def __init__(self, cfg: PipelineConfig, engine: Engine, clip=None):
        self.cfg, self.engine, self.clip = cfg, engine, clip
        self.device = engine.device

    def added_time_ids(self):
        return [self.cfg.fps_id, self.cfg.motion_bucket_id, self.cfg.noise_aug_strength]

    @torch.no_grad()
    def __call__(self, frames, num_inference_steps: int = 5, enc: Optional[torch.Tensor] = None,
                 aug_noise: Optional[torch.Tensor] = None, init_noise: Optional[torch.Tensor] = None,
                 generator: Optional[torch.Generator] = None, output_type: str = "np"):
        """frames: [T,H,W,3] float32 in [0,1] (numpy, or a pinned/CPU/CUDA tensor).
        Returns ``.frames[0]`` of the upstream pipeline: [T,H,W,3] float32 in [0,1]
        (numpy for output_type="np", a CUDA tensor for "pt")."""
        e, cfg = self.engine, self.cfg
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(frames)
        T, H, W, _ = frames.shape
        if H % 64 or W % 64:
            raise ValueError("height and width must be multiples of 64 (three stride-2 levels below /8 latents)")
        h, w = H // 8, W // 8
        frames = frames.to(self.device, non_blocking=True)
        if enc is None and self.clip is None:
            raise RuntimeError("no CLIP embedder configured and no `enc` given")
        # the two random draws of the upstream pipeline (noise augmentation, initial latents); drawn on the
        # generator's device -- a CUDA generator keeps 15 M normals off the host (70 ms per clip on CPU)
        gdev = generator.device if generator is not None else self.device
        if aug_noise is None:
            aug_noise = torch.randn((T, 3, H, W), generator=generator, device=gdev).to(self.device)
        if init_noise is None:
            init_noise = torch.randn((T, 4, h, w), generator=generator, device=gdev).to(self.device)
        e.prepare(T, h, w)
        if enc is None:
            cond, video = e.vae_encode_frames(frames, aug_noise, cfg.noise_aug_strength, want_video=True)
            enc = self.clip(video)
        else:
            cond = e.vae_encode_frames(frames, aug_noise, cfg.noise_aug_strength)
        e.set_clip_context(enc.to(self.device).float().reshape(T, -1))
        lat = e.denoise(cond, init_noise.reshape(T, 4, h, w), self.added_time_ids(), num_inference_steps)
        out = e.vae_decode_frames(lat, cfg.decode_chunk_size)
        if output_type == "pt":
            return out
        return out.cpu().numpy()
