"""Spatio-temporal UNet forward (SVD-XT topology, DepthCrafter overrides) -- oracle.

Functional restatement over a diffusers-keyed state dict of [UPSTREAM]
``UNetSpatioTemporalConditionModel`` / ``unet_3d_blocks`` /
``transformer_temporal`` / ``attention`` / ``resnet`` / ``embeddings`` as described
in SURVEY.md App. A.3, with DepthCrafter's per-frame ``encoder_hidden_states``
(``[B,T,1024] -> [B*T,1,1024]``).  Reference call site: model/depthcrafter.py:80-90.
Layout is the upstream one (NCHW / [B,C,T,H,W]) on purpose: the CUDA path uses a
different layout (NHWC tokens) and must agree anyway.
Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# bench.py's library-baseline arm only (torch eager fp16 on the GPU): route self-attention through
# F.scaled_dot_product_attention (flash kernels) instead of materialising the scores, as the reference's
# xformers / SDPA processors do (model/depthcrafter.py:33).  Parity tests keep the explicit softmax.
USE_SDPA = False


# ------------------------------------------------------------------ small pieces
def sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t.float()[:, None] * freqs[None, :]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def linear(sd: SD, key: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def group_norm(sd: SD, key: str, x: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    return F.group_norm(x, groups, sd[key + ".weight"], sd[key + ".bias"], eps)


def layer_norm(sd: SD, key: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)


def conv2d(sd: SD, key: str, x: torch.Tensor, stride: int = 1, padding: int = 1) -> torch.Tensor:
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=padding)


def conv3d_t(sd: SD, key: str, x: torch.Tensor) -> torch.Tensor:
    """Conv3d kernel (3,1,1), padding (1,0,0) on [B,C,T,H,W]."""
    return F.conv3d(x, sd[key + ".weight"], sd[key + ".bias"], padding=(1, 0, 0))


def alpha_blend(sd: SD, key: str, x_spatial, x_temporal, switch: bool = False):
    """AlphaBlender with image_only_indicator == 0: alpha = sigmoid(mix_factor)."""
    alpha = torch.sigmoid(sd[key + ".mix_factor"].float()).to(x_spatial.dtype)
    if switch:
        alpha = 1.0 - alpha
    return alpha * x_spatial + (1.0 - alpha) * x_temporal


def attention(sd: SD, key: str, x: torch.Tensor, ctx: torch.Tensor, heads: int) -> torch.Tensor:
    """diffusers Attention (AttnProcessor): softmax(q k^T / sqrt(d)) v, then to_out.0."""
    b, n, c = x.shape
    q = linear(sd, key + ".to_q", x)
    k = linear(sd, key + ".to_k", ctx)
    v = linear(sd, key + ".to_v", ctx)
    d = c // heads
    q = q.view(b, n, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    if USE_SDPA:
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
        return linear(sd, key + ".to_out.0", o)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    p = torch.softmax(s.float(), dim=-1).to(v.dtype)
    o = torch.matmul(p, v).transpose(1, 2).reshape(b, n, c)
    return linear(sd, key + ".to_out.0", o)


def feed_forward(sd: SD, key: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward(activation_fn='geglu'): Linear(C,8C) -> a * gelu(gate) -> Linear(4C,C)."""
    h = linear(sd, key + ".net.0.proj", x)
    a, gate = h.chunk(2, dim=-1)
    return linear(sd, key + ".net.2", a * F.gelu(gate))


# ------------------------------------------------------------------ blocks
def resnet_block_2d(sd, key, x, temb, groups, eps):
    h = F.silu(group_norm(sd, key + ".norm1", x, groups, eps))
    h = conv2d(sd, key + ".conv1", h)
    if temb is not None and (key + ".time_emb_proj.weight") in sd:
        h = h + linear(sd, key + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = F.silu(group_norm(sd, key + ".norm2", h, groups, eps))
    h = conv2d(sd, key + ".conv2", h)
    if (key + ".conv_shortcut.weight") in sd:
        x = conv2d(sd, key + ".conv_shortcut", x, padding=0)
    return x + h


def temporal_resnet_block(sd, key, x, temb, groups, eps):
    """x: [B,C,T,H,W]; temb: [B,T,E] or None."""
    h = F.silu(group_norm(sd, key + ".norm1", x, groups, eps))
    h = conv3d_t(sd, key + ".conv1", h)
    if temb is not None and (key + ".time_emb_proj.weight") in sd:
        t = linear(sd, key + ".time_emb_proj", F.silu(temb))          # [B,T,C]
        h = h + t.permute(0, 2, 1)[:, :, :, None, None]
    h = F.silu(group_norm(sd, key + ".norm2", h, groups, eps))
    h = conv3d_t(sd, key + ".conv2", h)
    return x + h


def st_res_block(sd, key, x, temb, num_frames, groups, eps, temporal_eps=None, switch=False):
    """SpatioTemporalResBlock on [B*T,C,H,W]; temb [B*T,E] or None."""
    x = resnet_block_2d(sd, key + ".spatial_res_block", x, temb, groups, eps)
    bt, c, h, w = x.shape
    b = bt // num_frames
    x5 = x.view(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
    temb5 = temb.view(b, num_frames, -1) if temb is not None else None
    xt = temporal_resnet_block(sd, key + ".temporal_res_block", x5, temb5, groups,
                               eps if temporal_eps is None else temporal_eps)
    out = alpha_blend(sd, key + ".time_mixer", x5, xt, switch)
    return out.permute(0, 2, 1, 3, 4).reshape(bt, c, h, w)


def basic_transformer_block(sd, key, x, ctx, heads, ln_eps):
    x = x + attention(sd, key + ".attn1", layer_norm(sd, key + ".norm1", x, ln_eps),
                      layer_norm(sd, key + ".norm1", x, ln_eps), heads)
    x = x + attention(sd, key + ".attn2", layer_norm(sd, key + ".norm2", x, ln_eps), ctx, heads)
    x = x + feed_forward(sd, key + ".ff", layer_norm(sd, key + ".norm3", x, ln_eps))
    return x


def temporal_transformer_block(sd, key, x, ctx, heads, num_frames, ln_eps):
    """x: [B*T, HW, C] -> attention over T per pixel; ctx: [B*HW, 1, D]."""
    bt, s, c = x.shape
    b = bt // num_frames
    h = x.view(b, num_frames, s, c).permute(0, 2, 1, 3).reshape(b * s, num_frames, c)
    res = h
    h = feed_forward(sd, key + ".ff_in", layer_norm(sd, key + ".norm_in", h, ln_eps)) + res
    n1 = layer_norm(sd, key + ".norm1", h, ln_eps)
    h = h + attention(sd, key + ".attn1", n1, n1, heads)
    h = h + attention(sd, key + ".attn2", layer_norm(sd, key + ".norm2", h, ln_eps), ctx, heads)
    h = h + feed_forward(sd, key + ".ff", layer_norm(sd, key + ".norm3", h, ln_eps))
    return h.view(b, s, num_frames, c).permute(0, 2, 1, 3).reshape(bt, s, c)


def st_transformer(sd, key, x, enc, heads, num_frames, groups, gn_eps, ln_eps):
    """TransformerSpatioTemporalModel on [B*T,C,H,W]; enc [B*T,1,D]."""
    bt, c, hh, ww = x.shape
    b = bt // num_frames
    first = enc.view(b, num_frames, -1, enc.shape[-1])[:, 0]                  # [B,1,D]
    time_ctx = first[:, None].expand(b, hh * ww, first.shape[-2], first.shape[-1])
    time_ctx = time_ctx.reshape(b * hh * ww, first.shape[-2], first.shape[-1])
    res = x
    h = group_norm(sd, key + ".norm", x, groups, gn_eps)
    h = h.permute(0, 2, 3, 1).reshape(bt, hh * ww, c)
    h = linear(sd, key + ".proj_in", h)
    frame_idx = torch.arange(num_frames, device=x.device).repeat(b)
    t_emb = sinusoid(frame_idx, c).to(x.dtype)
    emb = linear(sd, key + ".time_pos_embed.linear_2",
                 F.silu(linear(sd, key + ".time_pos_embed.linear_1", t_emb)))[:, None, :]
    h = basic_transformer_block(sd, key + ".transformer_blocks.0", h, enc, heads, ln_eps)
    h_mix = temporal_transformer_block(sd, key + ".temporal_transformer_blocks.0", h + emb,
                                       time_ctx, heads, num_frames, ln_eps)
    h = alpha_blend(sd, key + ".time_mixer", h, h_mix)
    h = linear(sd, key + ".proj_out", h)
    h = h.view(bt, hh, ww, c).permute(0, 3, 1, 2)
    return h + res


# ------------------------------------------------------------------ whole UNet
def unet_forward(sd: SD, cfg, sample: torch.Tensor, timestep: float, enc: torch.Tensor,
                 added_time_ids: torch.Tensor) -> torch.Tensor:
    """sample [B,T,8,h,w], timestep scalar, enc [B,T,D], added_time_ids [B,3] -> [B,T,4,h,w]."""
    b, t, _, hh, ww = sample.shape
    dtype = sample.dtype
    boc = cfg.block_out_channels
    nb = len(boc)
    g = cfg.norm_groups

    ts = torch.full((b,), float(timestep), device=sample.device)
    t_emb = sinusoid(ts, boc[0]).to(dtype)
    emb = linear(sd, "time_embedding.linear_2", F.silu(linear(sd, "time_embedding.linear_1", t_emb)))
    ids = sinusoid(added_time_ids.flatten(), cfg.addition_time_embed_dim).reshape(b, -1).to(dtype)
    emb = emb + linear(sd, "add_embedding.linear_2", F.silu(linear(sd, "add_embedding.linear_1", ids)))
    emb = emb.repeat_interleave(t, dim=0)                                    # [B*T,E]

    x = sample.flatten(0, 1)
    enc = enc.flatten(0, 1).unsqueeze(1)                                     # DepthCrafter: per frame
    x = conv2d(sd, "conv_in", x)
    skips = [x]

    for i in range(nb):
        has_attn = i < nb - 1
        eps = cfg.eps_cross_attn_block if has_attn else cfg.eps_plain_block
        for j in range(cfg.layers_per_block):
            x = st_res_block(sd, f"down_blocks.{i}.resnets.{j}", x, emb, t, g, eps)
            if has_attn:
                x = st_transformer(sd, f"down_blocks.{i}.attentions.{j}", x, enc,
                                   cfg.num_attention_heads[i], t, g, cfg.eps_transformer_norm, cfg.ln_eps)
            skips.append(x)
        if i < nb - 1:
            x = conv2d(sd, f"down_blocks.{i}.downsamplers.0.conv", x, stride=2, padding=1)
            skips.append(x)

    x = st_res_block(sd, "mid_block.resnets.0", x, emb, t, g, cfg.eps_plain_block)
    x = st_transformer(sd, "mid_block.attentions.0", x, enc, cfg.num_attention_heads[-1], t, g,
                       cfg.eps_transformer_norm, cfg.ln_eps)
    x = st_res_block(sd, "mid_block.resnets.1", x, emb, t, g, cfg.eps_plain_block)

    rev_heads = tuple(reversed(cfg.num_attention_heads))
    for i in range(nb):
        has_attn = i > 0
        eps = cfg.eps_cross_attn_block if has_attn else cfg.eps_plain_up_block
        for j in range(cfg.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = st_res_block(sd, f"up_blocks.{i}.resnets.{j}", x, emb, t, g, eps)
            if has_attn:
                x = st_transformer(sd, f"up_blocks.{i}.attentions.{j}", x, enc, rev_heads[i], t, g,
                                   cfg.eps_transformer_norm, cfg.ln_eps)
        if i < nb - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv2d(sd, f"up_blocks.{i}.upsamplers.0.conv", x)

    x = F.silu(group_norm(sd, "conv_norm_out", x, g, cfg.eps_out_norm))
    x = conv2d(sd, "conv_out", x)
    return x.view(b, t, -1, hh, ww)
