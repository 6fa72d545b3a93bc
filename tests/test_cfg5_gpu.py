"""BASELINE cfg5 (49 x 576 x 1024 long clip, bf16): the shapes that only exist there -- spatial attention over
N = 72 x 128 = 9216 tokens, temporal attention over T = 49 frames, the VAE at 576 x 1024 -- checked against plain
torch restatements of the same op at that size, and the full-size UNet step through size-independent properties
(finite, rerun bit-identical).  Tolerance: bf16 per-op rel-L2 <= 2e-2."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

T, H, W = 49, 576, 1024
h, w = H // 8, W // 8


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_spatial_attention_9216_tokens(cuda):
    from unigeo_b200 import ops
    F_, N, C = 2, h * w, 320                                                   # L0 of the UNet: 5 heads of 64
    g = torch.Generator(device=cuda).manual_seed(1)
    qkv = torch.randn(F_ * N, 3 * C, generator=g, device=cuda).to(torch.bfloat16)
    out = ops.spatial_attention(qkv, F_, N, C, head_dim=64)
    q, k, v = (t.float().reshape(F_, N, C // 64, 64).transpose(1, 2) for t in qkv.split(C, dim=1))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(F_ * N, C)
    assert rel_l2(out, ref) <= 2e-2, rel_l2(out, ref)


def test_temporal_attention_49_frames(cuda):
    from unigeo_b200 import ops
    P, C = 2304, 320
    g = torch.Generator(device=cuda).manual_seed(2)
    qkv = torch.randn(T, P, 3 * C, generator=g, device=cuda).to(torch.bfloat16)
    out = ops.temporal_attention(qkv, T, P, C)
    q, k, v = (t.float().reshape(T, P, C // 64, 64).permute(1, 2, 0, 3) for t in qkv.split(C, dim=2))   # [P, heads, T, 64]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(2, 0, 1, 3).reshape(T, P, C)
    assert rel_l2(out, ref) <= 2e-2, rel_l2(out, ref)


@pytest.fixture(scope="module")
def engine(cuda):
    from unigeo_b200.config import full_config
    from unigeo_b200.engine import Engine
    from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes, vae_param_shapes
    cfg = full_config()
    e = Engine(cfg, dtype="bf16", device=0)
    e.load_state_dict("unet", synthetic_state_dict(unet_param_shapes(cfg.unet), 1000, torch.float16, e.device))
    e.load_state_dict("vae", synthetic_state_dict(vae_param_shapes(cfg.vae), 2000, torch.float16, e.device))
    e.finalize()
    yield cfg, e
    e.close()
    torch.cuda.empty_cache()


def test_cfg5_denoising_step_is_finite_and_deterministic(engine, cuda):
    cfg, e = engine
    g = torch.Generator(device=cuda).manual_seed(3)
    cond = torch.randn(T, 4, h, w, generator=g, device=cuda)
    noise = torch.randn(T, 4, h, w, generator=g, device=cuda)
    e.prepare(T, h, w)
    e.set_clip_context(torch.randn(T, cfg.clip_embed_dim, generator=g, device=cuda))
    ids = [cfg.fps_id, cfg.motion_bucket_id, cfg.noise_aug_strength]
    a = e.denoise(cond, noise, ids, 1).clone()
    # eager call, graph capture, then replays: this size is where an intermittent (1 in 10) deviation under programmatic
    # dependent launch showed up (profiles/r02_pdl_race.txt); 16 calls make a regression visible with p = 0.8
    outs = [e.denoise(cond, noise, ids, 1).clone() for _ in range(15)]
    torch.cuda.synchronize()
    assert torch.isfinite(a).all()
    assert all(torch.equal(a, b) for b in outs), [float((a - b).abs().max()) for b in outs]
    assert e.workspace_bytes() < 60 * 2 ** 30                                   # one clip fits a fraction of 180 GB


def test_cfg5_vae_frames_and_chunks(engine, cuda):
    """576 x 1024 frames: decode chunks are independent (bit-exact), the encoder is per frame, values in range."""
    cfg, e = engine
    g = torch.Generator(device=cuda).manual_seed(4)
    lat = torch.randn(9, 4, h, w, generator=g, device=cuda) * 0.5
    dec = e.vae_decode_frames(lat, 8)
    assert dec.shape == (9, H, W, 3) and float(dec.min()) >= 0.0 and float(dec.max()) <= 1.0
    parts = torch.cat([e.vae_decode_frames(lat[:8].contiguous(), 8), e.vae_decode_frames(lat[8:].contiguous(), 8)])
    assert torch.equal(dec, parts)
    frames = torch.rand(3, H, W, 3, generator=g, device=cuda)
    whole = e.vae_encode_frames(frames)
    one = e.vae_encode_frames(frames[1:2].contiguous())
    assert torch.isfinite(whole).all() and rel_l2(one[0], whole[1]) <= 2e-2
