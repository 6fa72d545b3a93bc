"""AutoencoderKLTemporalDecoder encode / temporal decode -- oracle.

Functional restatement of [UPSTREAM] diffusers
``autoencoder_kl_temporal_decoder.py`` + ``vae.py`` (Encoder) as described in
SURVEY.md App. A.4.  Reference call site: model/depthcrafter.py:80-90 (the
pipeline encodes every frame and decodes in chunks of 8).
Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .unet_st import (SD, attention, conv2d, conv3d_t, group_norm, resnet_block_2d, st_res_block)


def _mid_attention(sd: SD, key: str, x: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    """diffusers Attention(norm_num_groups=32, residual_connection=True, bias=True), 1 head."""
    b, c, h, w = x.shape
    y = group_norm(sd, key + ".group_norm", x.view(b, c, h * w), groups, eps)
    y = y.transpose(1, 2)                                     # [B,HW,C]
    y = attention(sd, key, y, y, heads=1)
    return y.transpose(1, 2).reshape(b, c, h, w) + x


def vae_encode(sd: SD, cfg, img: torch.Tensor) -> torch.Tensor:
    """img [N,3,H,W] in [-1,1] -> latent_dist.mode() [N,4,H/8,W/8] (no scaling factor)."""
    g, eps = cfg.norm_groups, cfg.eps
    nb = len(cfg.block_out_channels)
    x = conv2d(sd, "encoder.conv_in", img)
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            x = resnet_block_2d(sd, f"encoder.down_blocks.{i}.resnets.{j}", x, None, g, eps)
        if i < nb - 1:
            x = F.pad(x, (0, 1, 0, 1))
            x = conv2d(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", x, stride=2, padding=0)
    x = resnet_block_2d(sd, "encoder.mid_block.resnets.0", x, None, g, eps)
    x = _mid_attention(sd, "encoder.mid_block.attentions.0", x, g, eps)
    x = resnet_block_2d(sd, "encoder.mid_block.resnets.1", x, None, g, eps)
    x = F.silu(group_norm(sd, "encoder.conv_norm_out", x, g, eps))
    x = conv2d(sd, "encoder.conv_out", x)
    moments = conv2d(sd, "quant_conv", x, padding=0)
    return moments[:, : cfg.latent_channels]


def vae_decode_chunk(sd: SD, cfg, z: torch.Tensor, num_frames: int) -> torch.Tensor:
    """TemporalDecoder on one chunk: z [F,4,h,w] (already / scaling_factor) -> [F,3,H,W]."""
    g, eps, teps = cfg.norm_groups, cfg.eps, cfg.temporal_eps
    nb = len(cfg.block_out_channels)

    def st(key, x):
        return st_res_block(sd, key, x, None, num_frames, g, eps, temporal_eps=teps, switch=True)

    x = conv2d(sd, "decoder.conv_in", z)
    x = st("decoder.mid_block.resnets.0", x)
    x = _mid_attention(sd, "decoder.mid_block.attentions.0", x, g, eps)
    x = st("decoder.mid_block.resnets.1", x)
    for i in range(nb):
        for j in range(cfg.layers_per_block + 1):
            x = st(f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i < nb - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv2d(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(group_norm(sd, "decoder.conv_norm_out", x, g, eps))
    x = conv2d(sd, "decoder.conv_out", x)
    f, c, h, w = x.shape
    b = f // num_frames
    x5 = x.view(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
    x5 = conv3d_t(sd, "decoder.time_conv_out", x5)
    return x5.permute(0, 2, 1, 3, 4).reshape(f, c, h, w)


def vae_decode(sd: SD, cfg, latents: torch.Tensor, chunk: int = 8) -> torch.Tensor:
    """latents [T,4,h,w] -> frames [T,3,H,W]; z / 0.18215, chunks of <= ``chunk`` frames,
    each chunk's temporal convs zero-padded at the chunk edges (App. A.1 step 9)."""
    z = latents / cfg.scaling_factor
    outs = []
    for i in range(0, z.shape[0], chunk):
        zc = z[i:i + chunk]
        outs.append(vae_decode_chunk(sd, cfg, zc, zc.shape[0]))
    return torch.cat(outs, dim=0)


def vae_decode_2d(sd: SD, cfg, latents: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.decode of the 2-D (StableNormal, SD-2.1 class) VAE: latents [N,4,h,w] (scaled) ->
    [N,3,8h,8w]; z / scaling_factor -> post_quant_conv -> Decoder ([UPSTREAM] diffusers
    ``autoencoder_kl.py`` / ``vae.py`` Decoder: mid Res-Attn-Res, 4 UpDecoderBlock2D of 3 resnets).
    Reference call site: model/stablenormal.py:39 (inside the hub predictor)."""
    g, eps = cfg.norm_groups, cfg.eps
    nb = len(cfg.block_out_channels)
    z = latents / cfg.scaling_factor
    x = conv2d(sd, "post_quant_conv", z, padding=0)
    x = conv2d(sd, "decoder.conv_in", x)
    x = resnet_block_2d(sd, "decoder.mid_block.resnets.0", x, None, g, eps)
    x = _mid_attention(sd, "decoder.mid_block.attentions.0", x, g, eps)
    x = resnet_block_2d(sd, "decoder.mid_block.resnets.1", x, None, g, eps)
    for i in range(nb):
        for j in range(cfg.layers_per_block + 1):
            x = resnet_block_2d(sd, f"decoder.up_blocks.{i}.resnets.{j}", x, None, g, eps)
        if i < nb - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv2d(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(group_norm(sd, "decoder.conv_norm_out", x, g, eps))
    return conv2d(sd, "decoder.conv_out", x)
