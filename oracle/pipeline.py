"""DepthCrafterPipeline.__call__ as UniGeo invokes it -- oracle.

Restates SURVEY.md App. A.1 for the argument set fixed at the reference call site
model/depthcrafter.py:80-90: ``guidance_scale=1.0`` (one UNet pass per step, no
CFG), ``window_size=len(frames)`` (one window, overlap forced to 0),
``output_type="np"``.  Random draws (noise augmentation, initial latents) are
explicit inputs -- the reference passes no generator and is non-deterministic
(App. B.4), so both arms are fed the same tensors.
Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import torch

from . import scheduler as S
from .unet_st import unet_forward
from .vae import vae_decode, vae_encode


def added_time_ids(cfg, device=None) -> torch.Tensor:
    return torch.tensor([[cfg.fps_id, cfg.motion_bucket_id, cfg.noise_aug_strength]],
                        dtype=torch.float32, device=device)


def encode_cond_latents(vae_sd, cfg, frames: torch.Tensor, aug_noise: torch.Tensor) -> torch.Tensor:
    """frames [T,H,W,3] in [0,1], aug_noise [T,3,H,W] -> cond latents [1,T,4,h,w] (unscaled mode)."""
    video = frames.permute(0, 3, 1, 2) * 2.0 - 1.0
    video = video + cfg.noise_aug_strength * aug_noise
    return vae_encode(vae_sd, cfg.vae, video).unsqueeze(0)


def denoise(unet_sd, cfg, cond_latents, enc, init_noise, num_steps, trace=None) -> torch.Tensor:
    """Euler/Karras loop over [1,T,4,h,w] latents; returns the final latents (fp32)."""
    sig = S.karras_sigmas(num_steps, cfg.sigma_min, cfg.sigma_max, cfg.rho)
    ts = S.timesteps_from_sigmas(sig)
    ids = added_time_ids(cfg, cond_latents.device)
    lat = init_noise.float() * S.init_noise_sigma(sig)
    for i in range(num_steps):
        x_in = torch.cat([S.scale_model_input(lat, sig[i]).to(cond_latents.dtype), cond_latents], dim=2)
        v = unet_forward(unet_sd, cfg.unet, x_in, ts[i], enc, ids)
        lat = S.euler_step(v, lat, sig[i], sig[i + 1])
        if trace is not None:
            trace.append(lat.clone())
    return lat


def decode_frames(vae_sd, cfg, latents: torch.Tensor) -> torch.Tensor:
    """latents [1,T,4,h,w] -> frames [T,H,W,3] in [0,1] (postprocess_video 'np' layout)."""
    x = vae_decode(vae_sd, cfg.vae, latents[0], cfg.decode_chunk_size).float()
    x = (x / 2.0 + 0.5).clamp(0.0, 1.0)
    return x.permute(0, 2, 3, 1).contiguous()


def depthcrafter_pipeline(unet_sd, vae_sd, cfg, frames, enc, aug_noise, init_noise, num_steps):
    """frames [T,H,W,3] in [0,1]; enc [1,T,D]; -> ``.frames[0]`` equivalent [T,H,W,3] float32."""
    cond = encode_cond_latents(vae_sd, cfg, frames, aug_noise)
    lat = denoise(unet_sd, cfg, cond, enc, init_noise, num_steps)
    return decode_frames(vae_sd, cfg, lat)
