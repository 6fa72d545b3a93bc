"""Per-frame CLIP image embeddings (pipeline step 3) -- oracle / checker.

[UPSTREAM] ``encode_video``: ``_resize_with_antialiasing`` (restated below), CLIP normalisation, then the
LIBRARY ``transformers.CLIPVisionModelWithProjection`` run in fp32 on the same state dict the CUDA path
(``ug_clip_embed``, csrc/clip.cu) is given.  Reference call site: model/depthcrafter.py:80-90.
Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _gaussian_kernel1d(ks: int, sigma: float, device, dtype):
    x = torch.arange(ks, device=device, dtype=dtype) - (ks - 1) / 2.0
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def resize_with_antialiasing(x: torch.Tensor, size=(224, 224)) -> torch.Tensor:
    """[UPSTREAM] _resize_with_antialiasing: Gaussian blur (sigma = max((factor-1)/2, 0.001),
    odd kernel >= 3), then bicubic resize with align_corners=True."""
    h, w = x.shape[-2:]
    fh, fw = h / size[0], w / size[1]
    sh, sw = max((fh - 1.0) / 2.0, 0.001), max((fw - 1.0) / 2.0, 0.001)
    kh = max(int(2 * 2 * sh) | 1, 3)
    kw = max(int(2 * 2 * sw) | 1, 3)
    c = x.shape[1]
    k_h = _gaussian_kernel1d(kh, sh, x.device, x.dtype).view(1, 1, kh, 1).repeat(c, 1, 1, 1)
    k_w = _gaussian_kernel1d(kw, sw, x.device, x.dtype).view(1, 1, 1, kw).repeat(c, 1, 1, 1)
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2), mode="reflect")
    x = F.conv2d(F.conv2d(x, k_h, groups=c), k_w, groups=c)
    return F.interpolate(x, size=size, mode="bicubic", align_corners=True)


def clip_model(clip_cfg, state_dict):
    """transformers CLIPVisionModelWithProjection (fp32, eval) holding ``state_dict``."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    conf = CLIPVisionConfig(hidden_size=clip_cfg.hidden_size, intermediate_size=clip_cfg.intermediate_size,
                            num_hidden_layers=clip_cfg.num_hidden_layers,
                            num_attention_heads=clip_cfg.num_attention_heads, image_size=clip_cfg.image_size,
                            patch_size=clip_cfg.patch_size, projection_dim=clip_cfg.projection_dim,
                            hidden_act="gelu", layer_norm_eps=clip_cfg.layer_norm_eps)
    m = CLIPVisionModelWithProjection(conf)
    missing, unexpected = m.load_state_dict({k: v.float() for k, v in state_dict.items()}, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m.float().eval()


@torch.no_grad()
def clip_embed(clip_cfg, state_dict, video: torch.Tensor, chunk: int = 8) -> torch.Tensor:
    """video [T,3,H,W] in [-1,1] -> image_embeds [T, projection_dim] float32."""
    m = clip_model(clip_cfg, state_dict)
    s = clip_cfg.image_size
    x = resize_with_antialiasing(video.float(), (s, s))
    x = (x + 1.0) / 2.0
    mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    x = (x - mean) / std
    outs = [m(pixel_values=x[i:i + chunk]).image_embeds for i in range(0, x.shape[0], chunk)]
    return torch.cat(outs, 0).float()
