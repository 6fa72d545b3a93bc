"""Architecture configs for the hot path.

Two sizes of the same topology:
  * ``full``  - the SVD-XT UNet / temporal-decoder VAE DepthCrafter ships
                ([UPSTREAM] config.json values, restated in SURVEY.md App. A).
  * ``tiny``  - same graph, narrow channels, for CPU-speed parity tests.  head_dim
                stays 64 and GroupNorm(32) divides every width, so every kernel
                variant the full model uses is exercised.

Every [UPSTREAM] detail that could not be verified in this container (GroupNorm
eps per block type, fps id, decode chunk) is a field here, not a constant in code.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Tuple


@dataclass(frozen=True)
class UNetSTConfig:
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    num_attention_heads: Tuple[int, ...] = (5, 10, 20, 20)
    layers_per_block: int = 2
    cross_attention_dim: int = 1024
    addition_time_embed_dim: int = 256
    num_added_ids: int = 3                     # fps, motion bucket, noise aug
    norm_groups: int = 32
    # [UPSTREAM] eps values (unet_3d_blocks.py): cross-attn blocks build their
    # SpatioTemporalResBlocks with 1e-6, plain down/up/mid blocks with 1e-5.
    eps_cross_attn_block: float = 1e-6
    eps_plain_block: float = 1e-5            # DownBlockSpatioTemporal / mid block: hard-coded 1e-5 upstream
    eps_plain_up_block: float = 1e-6         # UpBlockSpatioTemporal: resnet_eps default 1e-6, not overridden by get_up_block
    eps_transformer_norm: float = 1e-6
    eps_out_norm: float = 1e-5
    ln_eps: float = 1e-5

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @property
    def head_dim(self) -> int:
        return self.block_out_channels[0] // self.num_attention_heads[0]


@dataclass(frozen=True)
class VAEConfig:
    in_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_groups: int = 32
    eps: float = 1e-6
    temporal_eps: float = 1e-5
    scaling_factor: float = 0.18215


@dataclass(frozen=True)
class ClipConfig:
    """CLIP image encoder (transformers ``CLIPVisionConfig`` fields); default = ViT-H/14 as shipped in
    stable-video-diffusion-img2vid-xt/image_encoder."""
    hidden_size: int = 1280
    num_hidden_layers: int = 32
    num_attention_heads: int = 16
    intermediate_size: int = 5120
    patch_size: int = 14
    image_size: int = 224
    projection_dim: int = 1024
    layer_norm_eps: float = 1e-5


@dataclass(frozen=True)
class PipelineConfig:
    unet: UNetSTConfig = field(default_factory=UNetSTConfig)
    vae: VAEConfig = field(default_factory=VAEConfig)
    clip_embed_dim: int = 1024
    clip: ClipConfig = field(default_factory=ClipConfig)
    # scheduler
    sigma_min: float = 0.002
    sigma_max: float = 700.0
    rho: float = 7.0
    # pipeline (App. A.1)
    fps_id: float = 7.0
    motion_bucket_id: float = 127.0
    noise_aug_strength: float = 0.02
    decode_chunk_size: int = 8


@dataclass(frozen=True)
class UNet2DConfig:
    """SD-2.1 class ``UNet2DConditionModel`` / ``ControlNetModel`` dims."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    num_attention_heads: Tuple[int, ...] = (5, 10, 20, 20)    # head_dim 64 everywhere
    layers_per_block: int = 2
    cross_attention_dim: int = 1024
    context_len: int = 77                                     # CLIP text tokens of the fixed prompt
    norm_groups: int = 32
    eps_resnet: float = 1e-5
    eps_transformer_norm: float = 1e-6
    ln_eps: float = 1e-5

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4


@dataclass(frozen=True)
class StableNormalConfig:
    unet2d: UNet2DConfig = field(default_factory=UNet2DConfig)
    vae2d: VAEConfig = field(default_factory=VAEConfig)       # AutoencoderKL: same encoder, 2-D decoder
    num_train_timesteps: int = 1000
    beta_start: float = 0.00085
    beta_end: float = 0.012
    num_inference_steps: int = 10
    processing_multiple: int = 64                             # predictor resizes to a multiple of this


def stablenormal_config(name: str = "full") -> StableNormalConfig:
    if name == "full":
        return StableNormalConfig()
    if name == "tiny":
        return StableNormalConfig(
            unet2d=UNet2DConfig(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4),
                                cross_attention_dim=64, context_len=13),
            vae2d=VAEConfig(block_out_channels=(32, 64, 128, 128)))
    raise ValueError(f"unknown config {name!r} (expected 'full' or 'tiny')")


def full_config() -> PipelineConfig:
    return PipelineConfig()


def tiny_config() -> PipelineConfig:
    """Narrow copy of the full graph: (64,128,256,256) UNet, (32,64,128,128) VAE."""
    return PipelineConfig(
        unet=UNetSTConfig(
            block_out_channels=(64, 128, 256, 256),
            num_attention_heads=(1, 2, 4, 4),
            cross_attention_dim=64,
            addition_time_embed_dim=32,
        ),
        vae=VAEConfig(block_out_channels=(32, 64, 128, 128)),
        clip_embed_dim=64,
        clip=ClipConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128,
                        patch_size=32, image_size=224, projection_dim=64),
    )


def get_config(name: str) -> PipelineConfig:
    if name == "full":
        return full_config()
    if name == "tiny":
        return tiny_config()
    raise ValueError(f"unknown config {name!r} (expected 'full' or 'tiny')")


def config_to_dict(cfg) -> dict:
    return asdict(cfg)
