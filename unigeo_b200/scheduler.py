"""Karras-sigma Euler schedule of the SVD configuration (host-side numbers only).

The denoising loop itself runs inside ``ug_denoise_clip``; these helpers expose the same
schedule to Python callers that drive ``ug_unet_st_forward`` step by step (tests, profiling).
[UPSTREAM] EulerDiscreteScheduler(use_karras_sigmas, v_prediction, "leading"), SURVEY.md App. A.2.
"""
from __future__ import annotations

import math
from typing import List


def karras_sigmas(num_steps: int, sigma_min: float = 0.002, sigma_max: float = 700.0, rho: float = 7.0) -> List[float]:
    lo, hi = sigma_min ** (1.0 / rho), sigma_max ** (1.0 / rho)
    sig = [(hi + (i / (num_steps - 1) if num_steps > 1 else 0.0) * (lo - hi)) ** rho for i in range(num_steps)]
    return sig + [0.0]


def unet_timesteps(sigmas: List[float]) -> List[float]:
    return [0.25 * math.log(s) for s in sigmas[:-1]]


def init_noise_sigma(sigmas: List[float]) -> float:
    return math.sqrt(sigmas[0] ** 2 + 1.0)
