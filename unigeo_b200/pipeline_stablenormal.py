"""StableNormal predictor over the B200 engine: what ``self.predictor(image)`` does at
/root/reference/model/stablenormal.py:39 (hub model ``Stable-X/StableNormal``),
for a batch of frames at once.

Host side only moves data; four C-ABI calls on the current stream --
``ug_vae2d_encode`` -> [``ug_unet2d_forward`` (YOSO start latent)] -> ``ug_refine_frames_2d`` ->
``ug_vae2d_decode`` (which also emits the 8-bit normal image the hub predictor returns).
The DINOv2 semantic prior of the upstream pipeline is not part of this path.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .config import StableNormalConfig
from .engine import Engine


class StableNormalPipelineB200:
    UNET, CONTROLNET, YOSO_UNET, YOSO_CONTROLNET = "unet2d", "controlnet", "yoso_unet", "yoso_controlnet"

    def __init__(self, sn_cfg: StableNormalConfig, engine: Engine, prompt_embeds: torch.Tensor,
                 controlnet: bool = True, yoso: bool = False):
        self.cfg, self.engine = sn_cfg, engine
        self.device = engine.device
        self.controlnet = self.CONTROLNET if controlnet else None
        self.yoso = yoso
        nets = [self.UNET] + ([self.CONTROLNET] if controlnet else [])
        if yoso:
            nets += [self.YOSO_UNET] + ([self.YOSO_CONTROLNET] if controlnet else [])
        for net in nets:                       # K | V of every cross-attention: once per prompt
            engine.set_text_context(net, prompt_embeds)

    @torch.no_grad()
    def __call__(self, frames_u8, num_inference_steps: Optional[int] = None,
                 init_noise: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None):
        """frames_u8: uint8 [F,H,W,3] (numpy or tensor), H and W multiples of 64.
        Returns the 8-bit normal maps [F,H,W,3] as a CUDA uint8 tensor."""
        e, cfg = self.engine, self.cfg
        steps = int(num_inference_steps or cfg.num_inference_steps)
        if isinstance(frames_u8, np.ndarray):
            frames_u8 = torch.from_numpy(frames_u8)
        F_, H, W, _ = frames_u8.shape
        if H % 64 or W % 64:
            raise ValueError("height and width must be multiples of 64")
        h, w = H // 8, W // 8
        img = frames_u8.to(self.device, non_blocking=True).float().permute(0, 3, 1, 2).contiguous() / 255.0 * 2.0 - 1.0
        image_latent = e.vae2d_encode(img, cfg.vae2d.scaling_factor)
        if init_noise is None:
            gdev = generator.device if generator is not None else self.device
            init_noise = torch.randn((F_, 4, h, w), generator=generator, device=gdev)
        lat = init_noise.to(self.device).float()
        if self.yoso:                          # one-step initialiser: its sample prediction is the start latent
            lat = e.unet2d_forward(self.YOSO_UNET, lat, float(cfg.num_train_timesteps - 1),
                                   self.YOSO_CONTROLNET if self.controlnet else None,
                                   image_latent if self.controlnet else None)
        lat = e.refine_2d(self.UNET, self.controlnet, image_latent, lat, steps)
        _, normals = e.vae2d_decode(lat, want_image=False, want_normals_u8=True)
        return normals
