"""StableNormal refinement loop (SD-2.1 class UNet + ControlNet, DDIM "sample" prediction) -- oracle.

Restates SURVEY.md App. A.5 for what the hub predictor does per frame behind
``self.predictor(image)`` (reference call site model/stablenormal.py:39): VAE-encode the RGB
frame, iterate ``UNet(latents, t, text_ctx, ControlNet(image_latent, t, text_ctx))`` under a
DDIM scheduler with ``prediction_type="sample"`` ([UPSTREAM] diffusers
``schedulers/scheduling_ddim.py`` arithmetic, eta = 0, scaled-linear betas 0.00085 -> 0.012,
"trailing" spacing), VAE-decode, renormalise to unit normals, map to 8-bit.  The one-step YOSO
initialiser is the same graph with its own weights: its sample prediction is the start latent.
The DINOv2 semantic prior of the upstream pipeline is NOT restated (no architecture available
offline; SURVEY.md App. A.5 "lower confidence").  PARITY UNPINNED.
Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from .unet_2d import controlnet_forward, unet2d_forward
from .vae import vae_decode_2d, vae_encode


def alphas_cumprod(num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> np.ndarray:
    """scaled_linear betas -> cumulative alpha products (float64 on the host)."""
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=np.float64) ** 2
    return np.cumprod(1.0 - betas)


def ddim_timesteps(num_steps: int, num_train: int = 1000, t_start: Optional[int] = None) -> List[int]:
    """'trailing' spacing from ``t_start`` (default num_train - 1) down: round(T' - k T'/N) - 1."""
    top = num_train if t_start is None else t_start + 1
    return [int(round(top - k * top / num_steps)) - 1 for k in range(num_steps)]


def ddim_step_sample(x0_pred: torch.Tensor, x: torch.Tensor, a_t: float, a_prev: float) -> torch.Tensor:
    """prediction_type="sample", eta=0: eps = (x - sqrt(a_t) x0) / sqrt(1 - a_t);
    x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps."""
    x, x0 = x.float(), x0_pred.float()
    eps = (x - math.sqrt(a_t) * x0) / math.sqrt(1.0 - a_t)
    return math.sqrt(a_prev) * x0 + math.sqrt(1.0 - a_prev) * eps


def refine(unet_sd, ctrl_sd, cfg, image_latent, ctx, latents, num_steps: int, t_start: Optional[int] = None,
           trace=None) -> torch.Tensor:
    """DDIM loop; image_latent / latents [F,4,h,w] (scaled), ctx [1 or F,L,D]."""
    ac = alphas_cumprod(cfg.num_train_timesteps, cfg.beta_start, cfg.beta_end)
    ts = ddim_timesteps(num_steps, cfg.num_train_timesteps, t_start)
    x = latents.float()
    for i, t in enumerate(ts):
        down = mid = None
        if ctrl_sd is not None:
            down, mid = controlnet_forward(ctrl_sd, cfg.unet2d, image_latent, float(t), ctx)
        x0 = unet2d_forward(unet_sd, cfg.unet2d, x.to(image_latent.dtype), float(t), ctx, down, mid)
        t_prev = ts[i + 1] if i + 1 < len(ts) else -1
        a_prev = float(ac[t_prev]) if t_prev >= 0 else 1.0      # set_alpha_to_one-style final step
        x = ddim_step_sample(x0, x, float(ac[t]), a_prev)
        if trace is not None:
            trace.append(x.clone())
    return x


def normals_to_u8(decoded: torch.Tensor) -> np.ndarray:
    """decoded [F,3,H,W] -> unit normals -> ((clip(n,-1,1)+1)/2*255) uint8 [F,H,W,3] (truncating cast)."""
    n = decoded.float()
    n = n / n.norm(dim=1, keepdim=True).clamp_min(1e-6)
    n = ((n.clamp(-1.0, 1.0) + 1.0) * 0.5 * 255.0).permute(0, 2, 3, 1)
    return n.cpu().numpy().astype(np.uint8)


def stablenormal_predict(unet_sd, ctrl_sd, vae_sd, cfg, frames_u8: np.ndarray, ctx, init_noise, num_steps: int,
                         yoso_unet_sd=None, yoso_ctrl_sd=None) -> np.ndarray:
    """frames_u8 [F,H,W,3] uint8 -> normal maps uint8 [F,H,W,3] (what ``np.array(predictor(image))`` holds)."""
    img = torch.from_numpy(frames_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2) * 2.0 - 1.0
    image_latent = vae_encode(vae_sd, cfg.vae2d, img) * cfg.vae2d.scaling_factor
    lat = init_noise.float()
    if yoso_unet_sd is not None:
        t0 = cfg.num_train_timesteps - 1
        down = mid = None
        if yoso_ctrl_sd is not None:
            down, mid = controlnet_forward(yoso_ctrl_sd, cfg.unet2d, image_latent, float(t0), ctx)
        lat = unet2d_forward(yoso_unet_sd, cfg.unet2d, lat, float(t0), ctx, down, mid).float()
    lat = refine(unet_sd, ctrl_sd, cfg, image_latent, ctx, lat, num_steps)
    return normals_to_u8(vae_decode_2d(vae_sd, cfg.vae2d, lat))
