"""EulerDiscreteScheduler, SVD configuration -- oracle copy.

[UPSTREAM] diffusers ``schedulers/scheduling_euler_discrete.py`` with
``prediction_type="v_prediction", use_karras_sigmas=True, sigma_min=0.002,
sigma_max=700, timestep_type="continuous", timestep_spacing="leading"``.
Reference call site: model/depthcrafter.py:80-90 (``num_inference_steps=5``).
Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch


def karras_sigmas(num_steps: int, sigma_min: float = 0.002, sigma_max: float = 700.0,
                  rho: float = 7.0) -> List[float]:
    """sigma_i = (smax^(1/rho) + r_i (smin^(1/rho) - smax^(1/rho)))^rho, r=linspace(0,1,N); append 0."""
    lo = sigma_min ** (1.0 / rho)
    hi = sigma_max ** (1.0 / rho)
    out = []
    for i in range(num_steps):
        r = i / (num_steps - 1) if num_steps > 1 else 0.0
        out.append((hi + r * (lo - hi)) ** rho)
    out.append(0.0)
    return out


def timesteps_from_sigmas(sigmas: List[float]) -> List[float]:
    """Continuous timestep fed to the UNet: 0.25 * ln(sigma)."""
    return [0.25 * math.log(s) for s in sigmas[:-1]]


def init_noise_sigma(sigmas: List[float]) -> float:
    """timestep_spacing='leading' => sqrt(sigma_max^2 + 1)."""
    return math.sqrt(sigmas[0] ** 2 + 1.0)


def scale_model_input(x: torch.Tensor, sigma: float) -> torch.Tensor:
    return x / math.sqrt(sigma * sigma + 1.0)


def euler_step(v: torch.Tensor, x: torch.Tensor, sigma: float, sigma_next: float) -> torch.Tensor:
    """v-prediction Euler step, no churn, fp32 state."""
    x = x.float()
    v = v.float()
    x0 = v * (-sigma / math.sqrt(sigma * sigma + 1.0)) + x / (sigma * sigma + 1.0)
    d = (x - x0) / sigma
    return x + d * (sigma_next - sigma)
