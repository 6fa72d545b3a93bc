"""Parameter inventory (diffusers ``state_dict`` key names) and synthetic weights.

The key names follow the [UPSTREAM] diffusers modules the reference loads through
``from_pretrained`` (reference call sites: model/depthcrafter.py:18-29), so a real
``diffusion_pytorch_model.safetensors`` for the DepthCrafter UNet / SVD VAE can be
fed to ``ug_ctx_load_weight`` key by key the day one is on disk.  No weights exist
in this environment, so ``synthetic_state_dict`` draws seeded ones (SURVEY.md §8(d)
"Synthetic inputs"): N(0, 1/fan_in) matrices, N(0, 0.02) biases, N(1, 0.02) norm
gains, nothing zero-initialised, so every kernel contributes to the output.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import ClipConfig, UNet2DConfig, UNetSTConfig, VAEConfig

Shapes = "OrderedDict[str, Tuple[int, ...]]"


# --------------------------------------------------------------------------- helpers
def _lin(d, key, cin, cout, bias=True):
    d[key + ".weight"] = (cout, cin)
    if bias:
        d[key + ".bias"] = (cout,)


def _conv(d, key, cin, cout, k):
    d[key + ".weight"] = (cout, cin, k, k)
    d[key + ".bias"] = (cout,)


def _conv3d(d, key, cin, cout):
    d[key + ".weight"] = (cout, cin, 3, 1, 1)
    d[key + ".bias"] = (cout,)


def _norm(d, key, c):
    d[key + ".weight"] = (c,)
    d[key + ".bias"] = (c,)


def _resnet2d(d, key, cin, cout, temb):
    _norm(d, key + ".norm1", cin)
    _conv(d, key + ".conv1", cin, cout, 3)
    if temb:
        _lin(d, key + ".time_emb_proj", temb, cout)
    _norm(d, key + ".norm2", cout)
    _conv(d, key + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(d, key + ".conv_shortcut", cin, cout, 1)


def _temporal_resnet(d, key, c, temb):
    _norm(d, key + ".norm1", c)
    _conv3d(d, key + ".conv1", c, c)
    if temb:
        _lin(d, key + ".time_emb_proj", temb, c)
    _norm(d, key + ".norm2", c)
    _conv3d(d, key + ".conv2", c, c)


def _st_resblock(d, key, cin, cout, temb):
    _resnet2d(d, key + ".spatial_res_block", cin, cout, temb)
    _temporal_resnet(d, key + ".temporal_res_block", cout, temb)
    d[key + ".time_mixer.mix_factor"] = (1,)


def _attention(d, key, q_dim, ctx_dim, qkv_bias=False):
    _lin(d, key + ".to_q", q_dim, q_dim, bias=qkv_bias)
    _lin(d, key + ".to_k", ctx_dim, q_dim, bias=qkv_bias)
    _lin(d, key + ".to_v", ctx_dim, q_dim, bias=qkv_bias)
    _lin(d, key + ".to_out.0", q_dim, q_dim)


def _feed_forward(d, key, c):
    _lin(d, key + ".net.0.proj", c, 8 * c)       # GEGLU: value | gate
    _lin(d, key + ".net.2", 4 * c, c)


def _st_transformer(d, key, c, ctx):
    _norm(d, key + ".norm", c)
    _lin(d, key + ".proj_in", c, c)
    b = key + ".transformer_blocks.0"
    _norm(d, b + ".norm1", c)
    _attention(d, b + ".attn1", c, c)
    _norm(d, b + ".norm2", c)
    _attention(d, b + ".attn2", c, ctx)
    _norm(d, b + ".norm3", c)
    _feed_forward(d, b + ".ff", c)
    t = key + ".temporal_transformer_blocks.0"
    _norm(d, t + ".norm_in", c)
    _feed_forward(d, t + ".ff_in", c)
    _norm(d, t + ".norm1", c)
    _attention(d, t + ".attn1", c, c)
    _norm(d, t + ".norm2", c)
    _attention(d, t + ".attn2", c, ctx)
    _norm(d, t + ".norm3", c)
    _feed_forward(d, t + ".ff", c)
    _lin(d, key + ".time_pos_embed.linear_1", c, 4 * c)
    _lin(d, key + ".time_pos_embed.linear_2", 4 * c, c)
    d[key + ".time_mixer.mix_factor"] = (1,)
    _lin(d, key + ".proj_out", c, c)


# --------------------------------------------------------------------------- UNet
def unet_param_shapes(cfg: UNetSTConfig) -> Shapes:
    """Ordered {diffusers key: shape} of the spatio-temporal UNet (App. A.3)."""
    d: Shapes = OrderedDict()
    boc = cfg.block_out_channels
    temb = cfg.time_embed_dim
    ctx = cfg.cross_attention_dim
    nb = len(boc)
    _conv(d, "conv_in", cfg.in_channels, boc[0], 3)
    _lin(d, "time_embedding.linear_1", boc[0], temb)
    _lin(d, "time_embedding.linear_2", temb, temb)
    _lin(d, "add_embedding.linear_1", cfg.addition_time_embed_dim * cfg.num_added_ids, temb)
    _lin(d, "add_embedding.linear_2", temb, temb)
    # down path
    cout = boc[0]
    for i in range(nb):
        cin, cout = cout, boc[i]
        has_attn = i < nb - 1
        for j in range(cfg.layers_per_block):
            _st_resblock(d, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, temb)
            if has_attn:
                _st_transformer(d, f"down_blocks.{i}.attentions.{j}", cout, ctx)
        if i < nb - 1:
            _conv(d, f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    # mid
    _st_resblock(d, "mid_block.resnets.0", boc[-1], boc[-1], temb)
    _st_transformer(d, "mid_block.attentions.0", boc[-1], ctx)
    _st_resblock(d, "mid_block.resnets.1", boc[-1], boc[-1], temb)
    # up path
    rev = tuple(reversed(boc))
    cout = rev[0]
    for i in range(nb):
        prev = cout
        cout = rev[i]
        cin = rev[min(i + 1, nb - 1)]
        has_attn = i > 0
        for j in range(cfg.layers_per_block + 1):
            skip = cin if j == cfg.layers_per_block else cout
            rin = prev if j == 0 else cout
            _st_resblock(d, f"up_blocks.{i}.resnets.{j}", rin + skip, cout, temb)
            if has_attn:
                _st_transformer(d, f"up_blocks.{i}.attentions.{j}", cout, ctx)
        if i < nb - 1:
            _conv(d, f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    _norm(d, "conv_norm_out", boc[0])
    _conv(d, "conv_out", boc[0], cfg.out_channels, 3)
    return d


# --------------------------------------------------------------------------- VAE
def vae_param_shapes(cfg: VAEConfig) -> Shapes:
    """Ordered {diffusers key: shape} of AutoencoderKLTemporalDecoder (App. A.4)."""
    d: Shapes = OrderedDict()
    boc = cfg.block_out_channels
    nb = len(boc)
    # 2-D encoder
    _conv(d, "encoder.conv_in", cfg.in_channels, boc[0], 3)
    cout = boc[0]
    for i in range(nb):
        cin, cout = cout, boc[i]
        for j in range(cfg.layers_per_block):
            _resnet2d(d, f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, 0)
        if i < nb - 1:
            _conv(d, f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    c = boc[-1]
    _resnet2d(d, "encoder.mid_block.resnets.0", c, c, 0)
    _norm(d, "encoder.mid_block.attentions.0.group_norm", c)
    _attention(d, "encoder.mid_block.attentions.0", c, c, qkv_bias=True)
    _resnet2d(d, "encoder.mid_block.resnets.1", c, c, 0)
    _norm(d, "encoder.conv_norm_out", c)
    _conv(d, "encoder.conv_out", c, 2 * cfg.latent_channels, 3)
    _conv(d, "quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    # temporal decoder
    _conv(d, "decoder.conv_in", cfg.latent_channels, c, 3)
    _st_resblock(d, "decoder.mid_block.resnets.0", c, c, 0)
    _norm(d, "decoder.mid_block.attentions.0.group_norm", c)
    _attention(d, "decoder.mid_block.attentions.0", c, c, qkv_bias=True)
    _st_resblock(d, "decoder.mid_block.resnets.1", c, c, 0)
    rev = tuple(reversed(boc))
    cout = rev[0]
    for i in range(nb):
        cin, cout = cout, rev[i]
        for j in range(cfg.layers_per_block + 1):
            _st_resblock(d, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, 0)
        if i < nb - 1:
            _conv(d, f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    _norm(d, "decoder.conv_norm_out", boc[0])
    _conv(d, "decoder.conv_out", boc[0], cfg.in_channels, 3)
    _conv3d(d, "decoder.time_conv_out", cfg.in_channels, cfg.in_channels)
    return d


# --------------------------------------------------------------------------- 2-D UNet / ControlNet / VAE
def _transformer2d(d, key, c, ctx):
    _norm(d, key + ".norm", c)
    _lin(d, key + ".proj_in", c, c)
    b = key + ".transformer_blocks.0"
    _norm(d, b + ".norm1", c)
    _attention(d, b + ".attn1", c, c)
    _norm(d, b + ".norm2", c)
    _attention(d, b + ".attn2", c, ctx)
    _norm(d, b + ".norm3", c)
    _feed_forward(d, b + ".ff", c)
    _lin(d, key + ".proj_out", c, c)


def _encoder_half_2d(d, cfg: UNet2DConfig):
    boc = cfg.block_out_channels
    temb, ctx, nb = cfg.time_embed_dim, cfg.cross_attention_dim, len(boc)
    _conv(d, "conv_in", cfg.in_channels, boc[0], 3)
    _lin(d, "time_embedding.linear_1", boc[0], temb)
    _lin(d, "time_embedding.linear_2", temb, temb)
    skip_channels = [boc[0]]
    cout = boc[0]
    for i in range(nb):
        cin, cout = cout, boc[i]
        for j in range(cfg.layers_per_block):
            _resnet2d(d, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, temb)
            if i < nb - 1:
                _transformer2d(d, f"down_blocks.{i}.attentions.{j}", cout, ctx)
            skip_channels.append(cout)
        if i < nb - 1:
            _conv(d, f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
            skip_channels.append(cout)
    _resnet2d(d, "mid_block.resnets.0", boc[-1], boc[-1], temb)
    _transformer2d(d, "mid_block.attentions.0", boc[-1], ctx)
    _resnet2d(d, "mid_block.resnets.1", boc[-1], boc[-1], temb)
    return skip_channels


def unet2d_param_shapes(cfg: UNet2DConfig) -> Shapes:
    """Ordered {diffusers key: shape} of ``UNet2DConditionModel`` (SD-2.1 class, App. A.5)."""
    d: Shapes = OrderedDict()
    boc = cfg.block_out_channels
    temb, ctx, nb = cfg.time_embed_dim, cfg.cross_attention_dim, len(boc)
    _encoder_half_2d(d, cfg)
    rev = tuple(reversed(boc))
    cout = rev[0]
    for i in range(nb):
        prev = cout
        cout = rev[i]
        cin = rev[min(i + 1, nb - 1)]
        for j in range(cfg.layers_per_block + 1):
            skip = cin if j == cfg.layers_per_block else cout
            rin = prev if j == 0 else cout
            _resnet2d(d, f"up_blocks.{i}.resnets.{j}", rin + skip, cout, temb)
            if i > 0:
                _transformer2d(d, f"up_blocks.{i}.attentions.{j}", cout, ctx)
        if i < nb - 1:
            _conv(d, f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    _norm(d, "conv_norm_out", boc[0])
    _conv(d, "conv_out", boc[0], cfg.out_channels, 3)
    return d


def controlnet_param_shapes(cfg: UNet2DConfig) -> Shapes:
    """``ControlNetModel`` without conditioning embedding: encoder half + 1x1 'zero' convs."""
    d: Shapes = OrderedDict()
    skips = _encoder_half_2d(d, cfg)
    for i, c in enumerate(skips):
        _conv(d, f"controlnet_down_blocks.{i}", c, c, 1)
    _conv(d, "controlnet_mid_block", cfg.block_out_channels[-1], cfg.block_out_channels[-1], 1)
    return d


def vae2d_param_shapes(cfg: VAEConfig) -> Shapes:
    """``AutoencoderKL`` (SD class): the 2-D encoder of ``vae_param_shapes`` + a 2-D decoder."""
    d: Shapes = OrderedDict()
    for k, v in vae_param_shapes(cfg).items():
        if not k.startswith("decoder."):
            d[k] = v
    boc = cfg.block_out_channels
    nb, c = len(boc), boc[-1]
    _conv(d, "post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    _conv(d, "decoder.conv_in", cfg.latent_channels, c, 3)
    _resnet2d(d, "decoder.mid_block.resnets.0", c, c, 0)
    _norm(d, "decoder.mid_block.attentions.0.group_norm", c)
    _attention(d, "decoder.mid_block.attentions.0", c, c, qkv_bias=True)
    _resnet2d(d, "decoder.mid_block.resnets.1", c, c, 0)
    rev = tuple(reversed(boc))
    cout = rev[0]
    for i in range(nb):
        cin, cout = cout, rev[i]
        for j in range(cfg.layers_per_block + 1):
            _resnet2d(d, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, 0)
        if i < nb - 1:
            _conv(d, f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    _norm(d, "decoder.conv_norm_out", boc[0])
    _conv(d, "decoder.conv_out", boc[0], cfg.in_channels, 3)
    return d


# --------------------------------------------------------------------------- CLIP image encoder
def clip_param_shapes(cfg: ClipConfig) -> Shapes:
    """Ordered {transformers key: shape} of ``CLIPVisionModelWithProjection`` (state_dict names, including
    upstream's ``pre_layrnorm`` spelling)."""
    d: Shapes = OrderedDict()
    c, p = cfg.hidden_size, cfg.patch_size
    n_pos = (cfg.image_size // p) ** 2 + 1
    d["vision_model.embeddings.class_embedding"] = (c,)
    d["vision_model.embeddings.patch_embedding.weight"] = (c, 3, p, p)
    d["vision_model.embeddings.position_embedding.weight"] = (n_pos, c)
    _norm(d, "vision_model.pre_layrnorm", c)
    for i in range(cfg.num_hidden_layers):
        b = f"vision_model.encoder.layers.{i}"
        for k in ("k_proj", "v_proj", "q_proj", "out_proj"):
            _lin(d, f"{b}.self_attn.{k}", c, c)
        _norm(d, f"{b}.layer_norm1", c)
        _lin(d, f"{b}.mlp.fc1", c, cfg.intermediate_size)
        _lin(d, f"{b}.mlp.fc2", cfg.intermediate_size, c)
        _norm(d, f"{b}.layer_norm2", c)
    _norm(d, "vision_model.post_layernorm", c)
    _lin(d, "visual_projection", c, cfg.projection_dim, bias=False)
    return d


# --------------------------------------------------------------------------- synthetic
def _is_norm_key(key: str) -> bool:
    leaf = key.rsplit(".", 2)[-2]
    return (leaf.startswith("norm") or leaf.startswith("layer_norm")
            or leaf in ("group_norm", "conv_norm_out", "pre_layrnorm", "post_layernorm"))


def synthetic_state_dict(shapes: Shapes, seed: int, dtype=torch.float32, device="cpu") -> Dict[str, torch.Tensor]:
    """Seeded weights for a shape inventory (``dtype`` tensors on ``device``).

    One generator is advanced key by key in inventory order, so the same
    (shapes, seed, device type) always yields the same tensors.  Parity tests draw on the CPU (the
    oracle needs the identical values); benchmarks may draw on the GPU, where 1.5 B values take
    milliseconds instead of tens of seconds.
    """
    g = torch.Generator(device=device).manual_seed(seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in shapes.items():
        if key.endswith("mix_factor"):
            t = torch.randn(shape, generator=g, device=device) * 0.5
        elif _is_norm_key(key):
            t = torch.randn(shape, generator=g, device=device) * 0.02
            if key.endswith(".weight"):
                t = t + 1.0
        elif key.endswith(".bias"):
            t = torch.randn(shape, generator=g, device=device) * 0.02
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g, device=device) / float(fan_in) ** 0.5
        sd[key] = t.to(dtype)
    return sd


def count_params(shapes: Shapes) -> int:
    n = 0
    for s in shapes.values():
        p = 1
        for x in s:
            p *= x
        n += p
    return n


def load_diffusers_dir(path: str) -> Dict[str, torch.Tensor]:
    """Read every ``*.safetensors`` in a diffusers model directory into one {key: tensor} dict
    (the files ``from_pretrained`` reads at /root/reference/model/depthcrafter.py:18-29)."""
    import glob
    import os

    from safetensors.torch import load_file
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no .safetensors files under {path}")
    # prefer the fp16 variant when both exist, like variant="fp16" upstream
    fp16 = [f for f in files if ".fp16." in os.path.basename(f)]
    sd: Dict[str, torch.Tensor] = {}
    for f in (fp16 or files):
        sd.update(load_file(f))
    return sd
