"""2-D conditional UNet + ControlNet (SD-2.1 class) as StableNormal runs it -- oracle.

Functional restatement over diffusers-keyed state dicts of [UPSTREAM] diffusers
``UNet2DConditionModel`` (``unets/unet_2d_condition.py``, ``unets/unet_2d_blocks.py``:
CrossAttnDownBlock2D x3 + DownBlock2D, UNetMidBlock2DCrossAttn, UpBlock2D + CrossAttnUpBlock2D x3),
``Transformer2DModel(use_linear_projection=True)`` and ``ControlNetModel`` without a
conditioning embedding (StableNormal's ControlNet variant takes the RGB latent as its sample),
as summarised in SURVEY.md App. A.5.  Reference call site: model/stablenormal.py:16,39
(``torch.hub.load("Stable-X/StableNormal")`` -> ``predictor(image)``); the hub repo is neither
vendored nor pinned, hence PARITY UNPINNED like the rest of oracle/.
Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from .unet_st import (SD, basic_transformer_block, conv2d, group_norm, linear, resnet_block_2d, sinusoid)


def transformer_2d(sd: SD, key: str, x: torch.Tensor, ctx: torch.Tensor, heads: int, groups: int,
                   gn_eps: float, ln_eps: float) -> torch.Tensor:
    """Transformer2DModel, linear projections, one BasicTransformerBlock; ctx [B,L,D]."""
    b, c, h, w = x.shape
    y = group_norm(sd, key + ".norm", x, groups, gn_eps)
    y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
    y = linear(sd, key + ".proj_in", y)
    y = basic_transformer_block(sd, key + ".transformer_blocks.0", y, ctx, heads, ln_eps)
    y = linear(sd, key + ".proj_out", y)
    return y.view(b, h, w, c).permute(0, 3, 1, 2) + x


def _time_embedding(sd: SD, cfg, timestep: float, batch: int, device, dtype) -> torch.Tensor:
    ts = torch.full((batch,), float(timestep), device=device)
    t_emb = sinusoid(ts, cfg.block_out_channels[0]).to(dtype)
    return linear(sd, "time_embedding.linear_2", F.silu(linear(sd, "time_embedding.linear_1", t_emb)))


def _encoder_half(sd: SD, cfg, x: torch.Tensor, emb: torch.Tensor, ctx: torch.Tensor
                  ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """conv_in + down blocks + mid block; returns (mid output, skip list)."""
    boc = cfg.block_out_channels
    nb, g = len(boc), cfg.norm_groups
    x = conv2d(sd, "conv_in", x)
    skips = [x]
    for i in range(nb):
        has_attn = i < nb - 1
        for j in range(cfg.layers_per_block):
            x = resnet_block_2d(sd, f"down_blocks.{i}.resnets.{j}", x, emb, g, cfg.eps_resnet)
            if has_attn:
                x = transformer_2d(sd, f"down_blocks.{i}.attentions.{j}", x, ctx, cfg.num_attention_heads[i], g,
                                   cfg.eps_transformer_norm, cfg.ln_eps)
            skips.append(x)
        if i < nb - 1:
            x = conv2d(sd, f"down_blocks.{i}.downsamplers.0.conv", x, stride=2, padding=1)
            skips.append(x)
    x = resnet_block_2d(sd, "mid_block.resnets.0", x, emb, g, cfg.eps_resnet)
    x = transformer_2d(sd, "mid_block.attentions.0", x, ctx, cfg.num_attention_heads[-1], g,
                       cfg.eps_transformer_norm, cfg.ln_eps)
    x = resnet_block_2d(sd, "mid_block.resnets.1", x, emb, g, cfg.eps_resnet)
    return x, skips


def controlnet_forward(sd: SD, cfg, sample: torch.Tensor, timestep: float, ctx: torch.Tensor,
                       conditioning_scale: float = 1.0) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """ControlNet without conditioning embedding: encoder half on ``sample`` + 1x1 'zero' convs.
    sample [B,C,h,w]; ctx [B,L,D] -> (12 down residuals, mid residual)."""
    b = sample.shape[0]
    if ctx.shape[0] == 1 and b > 1:
        ctx = ctx.expand(b, -1, -1)
    emb = _time_embedding(sd, cfg, timestep, b, sample.device, sample.dtype)
    mid, skips = _encoder_half(sd, cfg, sample, emb, ctx)
    down = [conv2d(sd, f"controlnet_down_blocks.{i}", s, padding=0) * conditioning_scale
            for i, s in enumerate(skips)]
    return down, conv2d(sd, "controlnet_mid_block", mid, padding=0) * conditioning_scale


def unet2d_forward(sd: SD, cfg, sample: torch.Tensor, timestep: float, ctx: torch.Tensor,
                   down_residuals: Optional[List[torch.Tensor]] = None,
                   mid_residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sample [B,Cin,h,w], scalar timestep, ctx [B or 1,L,D] -> [B,Cout,h,w]."""
    b = sample.shape[0]
    if ctx.shape[0] == 1 and b > 1:
        ctx = ctx.expand(b, -1, -1)
    boc = cfg.block_out_channels
    nb, g = len(boc), cfg.norm_groups
    emb = _time_embedding(sd, cfg, timestep, b, sample.device, sample.dtype)
    x, skips = _encoder_half(sd, cfg, sample, emb, ctx)
    if down_residuals is not None:
        skips = [s + r for s, r in zip(skips, down_residuals)]
    if mid_residual is not None:
        x = x + mid_residual
    rev_heads = tuple(reversed(cfg.num_attention_heads))
    for i in range(nb):
        has_attn = i > 0
        for j in range(cfg.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block_2d(sd, f"up_blocks.{i}.resnets.{j}", x, emb, g, cfg.eps_resnet)
            if has_attn:
                x = transformer_2d(sd, f"up_blocks.{i}.attentions.{j}", x, ctx, rev_heads[i], g,
                                   cfg.eps_transformer_norm, cfg.ln_eps)
        if i < nb - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv2d(sd, f"up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(group_norm(sd, "conv_norm_out", x, g, cfg.eps_resnet))
    return conv2d(sd, "conv_out", x)
