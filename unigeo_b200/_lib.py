"""ctypes binding of libunigeo_b200.so (C ABI: include/unigeo_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing and cannot be
built, importing raises; if a call fails, ``UgError`` carries ``ug_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# UG_LIB selects a differently-compiled copy of the library (development tooling, e.g. the trace build of
# tools/trace_tapgemm.py); it must exist -- nothing is built or substituted on its behalf
LIB_PATH = os.environ.get("UG_LIB") or os.path.join(HERE, "libunigeo_b200.so")

UG_F16, UG_BF16, UG_F32 = 0, 1, 2


class UgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"unigeo_b200 error {code}: {msg}")
        self.code = code


class ModelCfg(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("unet_in_channels", C.c_int32), ("unet_out_channels", C.c_int32),
        ("unet_num_blocks", C.c_int32),
        ("unet_block_out", C.c_int32 * 4),
        ("unet_heads", C.c_int32 * 4),
        ("unet_layers_per_block", C.c_int32),
        ("cross_attention_dim", C.c_int32),
        ("addition_time_embed_dim", C.c_int32),
        ("num_added_ids", C.c_int32),
        ("norm_groups", C.c_int32),
        ("eps_cross_attn_block", C.c_float), ("eps_plain_block", C.c_float), ("eps_plain_up_block", C.c_float),
        ("eps_transformer_norm", C.c_float), ("eps_out_norm", C.c_float), ("ln_eps", C.c_float),
        ("vae_in_channels", C.c_int32), ("vae_latent_channels", C.c_int32),
        ("vae_num_blocks", C.c_int32),
        ("vae_block_out", C.c_int32 * 4),
        ("vae_layers_per_block", C.c_int32),
        ("vae_norm_groups", C.c_int32),
        ("vae_eps", C.c_float), ("vae_temporal_eps", C.c_float), ("vae_scaling_factor", C.c_float),
        ("sigma_min", C.c_float), ("sigma_max", C.c_float), ("rho", C.c_float),
    ]


class UNet2DCfg(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32), ("out_channels", C.c_int32),
        ("num_blocks", C.c_int32),
        ("block_out", C.c_int32 * 4),
        ("heads", C.c_int32 * 4),
        ("layers_per_block", C.c_int32),
        ("cross_attention_dim", C.c_int32),
        ("eps_resnet", C.c_float), ("eps_transformer_norm", C.c_float), ("ln_eps", C.c_float),
        ("num_train_timesteps", C.c_int32),
        ("beta_start", C.c_float), ("beta_end", C.c_float),
    ]


class ClipCfg(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32), ("mlp", C.c_int32),
                ("patch", C.c_int32), ("image_size", C.c_int32), ("proj_dim", C.c_int32), ("ln_eps", C.c_float)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_longlong, C.c_float

_SIGNATURES = {
    "ug_version": ([], C.c_int),
    "ug_last_error": ([], C.c_char_p),
    "ug_ctx_create": ([C.POINTER(_P), _I, C.POINTER(ModelCfg)], C.c_int),
    "ug_ctx_destroy": ([_P], C.c_int),
    "ug_ctx_load_weight": ([_P, C.c_char_p, _P, _I, C.POINTER(C.c_int64), _I, _P], C.c_int),
    "ug_ctx_set_vae_encode_dtype": ([_P, _I], C.c_int),
    "ug_ctx_load_weights": ([_P, _I, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                             C.POINTER(C.c_int64), C.POINTER(C.c_int), _P], C.c_int),
    "ug_ctx_finalize": ([_P, _P], C.c_int),
    "ug_ctx_prepare": ([_P, _I, _I, _I, _P], C.c_int),
    "ug_set_clip_context": ([_P, _P, _P], C.c_int),
    "ug_unet_st_forward": ([_P, _P, _F, C.POINTER(C.c_float), _P, _P], C.c_int),
    "ug_denoise_clip": ([_P, _P, _P, C.POINTER(C.c_float), _I, _P, _P], C.c_int),
    "ug_vae_encode": ([_P, _P, _P, _F, _I, _I, _I, _P, _P], C.c_int),
    "ug_vae_decode_temporal": ([_P, _P, _I, _I, _I, _I, _P, _P], C.c_int),
    "ug_ctx_set_unet2d_cfg": ([_P, C.POINTER(UNet2DCfg)], C.c_int),
    "ug_set_text_context": ([_P, C.c_char_p, _P, _I, _I, _P], C.c_int),
    "ug_unet2d_forward": ([_P, C.c_char_p, _P, _I, _I, _I, _F, C.c_char_p, _P, _P, _P], C.c_int),
    "ug_refine_frames_2d": ([_P, C.c_char_p, C.c_char_p, _P, _P, _I, _I, _I, _I, _I, _P, _P], C.c_int),
    "ug_vae2d_encode": ([_P, _P, _I, _I, _I, _F, _P, _P], C.c_int),
    "ug_vae2d_decode": ([_P, _P, _I, _I, _I, _P, _P, _P], C.c_int),
    "ug_ctx_set_clip_cfg": ([_P, C.POINTER(ClipCfg)], C.c_int),
    "ug_clip_embed": ([_P, _P, _I, _I, _I, _P, _P], C.c_int),
    "ug_prepare_frames": ([_P, _P, _I, _I, _I, _P, _P], C.c_int),
    "ug_vae_encode_frames": ([_P, _P, _P, _F, _I, _I, _I, _P, _P, _P], C.c_int),
    "ug_vae_decode_frames": ([_P, _P, _I, _I, _I, _I, _P, _P], C.c_int),
    "ug_depth_postprocess": ([_P, _P, _P, _I, _I, _I, _P, _P, _P], C.c_int),
    "ug_depth_metrics": ([_P, _P, _P, _P, _L, _F, C.POINTER(C.c_double), _P, _P, _P, _P], C.c_int),
    "ug_normal_metrics": ([_P, _P, _P, _P, _L, C.POINTER(C.c_double), _P, _P], C.c_int),
    "ug_stitch_fit": ([_P, _P, _I, _I, _I, _L, _I, _F, _P, _P], C.c_int),
    "ug_stitch_apply": ([_P, _P, _L, _P, _L, _L, _I, _P, _I, _I, _F, _P, _P], C.c_int),
    "ug_karras_schedule": ([C.POINTER(ModelCfg), _I, C.POINTER(C.c_double), C.POINTER(C.c_double),
                            C.POINTER(C.c_double)], C.c_int),
    "ug_ddim_schedule": ([C.POINTER(UNet2DCfg), _I, _I, C.POINTER(C.c_int), C.POINTER(C.c_double),
                          C.POINTER(C.c_double)], C.c_int),
    "ug_fastdiv": ([_I, _I], C.c_longlong),
    "ug_tile_schedule": ([_I, _I, _I, _I, _I, _I, _I, _I, C.POINTER(C.c_int), _I, C.POINTER(C.c_longlong),
                          C.POINTER(C.c_longlong)], C.c_int),
    "ug_ctx_launch_count": ([_P, _I], C.c_longlong),
    "ug_ctx_workspace_bytes": ([_P], C.c_longlong),
    "ug_ctx_graph_count": ([_P], C.c_longlong),
    "ug_ctx_profile": ([_P, _I], C.c_int),
    "ug_ctx_profile_read": ([_P, _I, C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                             C.POINTER(C.c_double), C.POINTER(C.c_double)], C.c_int),
    "ug_op_linear": ([_I, _P, _L, _I, _P, _I, _P, _P, _I, _I, _P, _P], C.c_int),
    "ug_op_linear_blend": ([_I, _P, _L, _I, _P, _I, _P, _P, _P, _F, _P, _P], C.c_int),
    "ug_op_conv3x3": ([_I, _P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P], C.c_int),
    "ug_op_tconv3": ([_I, _P, _I, _L, _I, _P, _I, _I, _P, _P, _P, _F, _P, _P], C.c_int),
    "ug_op_groupnorm": ([_I, _P, _I, _P, _I, _L, _L, _I, _P, _P, _F, _I, _P, _P], C.c_int),
    "ug_op_layernorm": ([_I, _P, _L, _I, _P, _P, _F, _P, _I, _P, _P], C.c_int),
    "ug_op_ln_linear": ([_I, _P, _I, _P, _P, _P, _P, _L, _I, _P, _P, _F, _P, _I, _P, _I, _P, _P], C.c_int),
    "ug_op_spatial_attention": ([_I, _P, _I, _I, _I, _I, _P, _P], C.c_int),
    "ug_op_temporal_attention": ([_I, _P, _I, _L, _I, _P, _P], C.c_int),
    "ug_op_cross_attention": ([_I, _P, _P, _I, _I, _I, _I, _I, _P, _P], C.c_int),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Load (building first if needed) the shared library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if os.environ.get("UG_LIB"):
            raise FileNotFoundError(f"UG_LIB={LIB_PATH} does not exist")
        from .build import build  # needs nvcc; raises with the compiler output otherwise
        build()
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError = the library does not match the header
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        raise UgError(code, load().ug_last_error().decode("utf-8", "replace"))


def cfg_struct(cfg, dtype: int) -> ModelCfg:
    """unigeo_b200.config.PipelineConfig -> ug_model_cfg."""
    u, v = cfg.unet, cfg.vae
    m = ModelCfg()
    m.dtype = dtype
    m.unet_in_channels, m.unet_out_channels = u.in_channels, u.out_channels
    m.unet_num_blocks = len(u.block_out_channels)
    for i, (c, h) in enumerate(zip(u.block_out_channels, u.num_attention_heads)):
        m.unet_block_out[i] = c
        m.unet_heads[i] = h
    m.unet_layers_per_block = u.layers_per_block
    m.cross_attention_dim = u.cross_attention_dim
    m.addition_time_embed_dim = u.addition_time_embed_dim
    m.num_added_ids = u.num_added_ids
    m.norm_groups = u.norm_groups
    m.eps_cross_attn_block, m.eps_plain_block = u.eps_cross_attn_block, u.eps_plain_block
    m.eps_plain_up_block = u.eps_plain_up_block
    m.eps_transformer_norm, m.eps_out_norm, m.ln_eps = u.eps_transformer_norm, u.eps_out_norm, u.ln_eps
    m.vae_in_channels, m.vae_latent_channels = v.in_channels, v.latent_channels
    m.vae_num_blocks = len(v.block_out_channels)
    for i, c in enumerate(v.block_out_channels):
        m.vae_block_out[i] = c
    m.vae_layers_per_block = v.layers_per_block
    m.vae_norm_groups = v.norm_groups
    m.vae_eps, m.vae_temporal_eps, m.vae_scaling_factor = v.eps, v.temporal_eps, v.scaling_factor
    m.sigma_min, m.sigma_max, m.rho = cfg.sigma_min, cfg.sigma_max, cfg.rho
    return m


def unet2d_cfg_struct(sn_cfg) -> UNet2DCfg:
    """unigeo_b200.config.StableNormalConfig -> ug_unet2d_cfg."""
    u = sn_cfg.unet2d
    m = UNet2DCfg()
    m.in_channels, m.out_channels = u.in_channels, u.out_channels
    m.num_blocks = len(u.block_out_channels)
    for i, (c, h) in enumerate(zip(u.block_out_channels, u.num_attention_heads)):
        m.block_out[i] = c
        m.heads[i] = h
    m.layers_per_block = u.layers_per_block
    m.cross_attention_dim = u.cross_attention_dim
    m.eps_resnet, m.eps_transformer_norm, m.ln_eps = u.eps_resnet, u.eps_transformer_norm, u.ln_eps
    m.num_train_timesteps = sn_cfg.num_train_timesteps
    m.beta_start, m.beta_end = sn_cfg.beta_start, sn_cfg.beta_end
    return m
