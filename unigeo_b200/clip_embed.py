"""Per-frame CLIP image embeddings (pipeline step 3) -- library code.

§8(f)-2 keeps this stage on torch/transformers for now: antialiased 224x224 resize, CLIP
normalisation, ``CLIPVisionModelWithProjection`` (ViT-H/14 for the full model; a narrow
random-init ViT for the tiny test config).  No checkpoint exists in this environment, so
weights are seeded random-init; a real ``image_encoder`` directory can be passed instead.
Both arms of every parity test are fed the embeddings produced here.
"""
from __future__ import annotations

import math

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


class ClipEmbedder:
    def __init__(self, embed_dim: int, device, dtype=torch.float16, seed: int = 7, pretrained: str | None = None):
        from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
        self.device, self.dtype = torch.device(device), dtype
        if pretrained:
            model = CLIPVisionModelWithProjection.from_pretrained(pretrained)
        else:
            if embed_dim == 1024:       # ViT-H/14 as in stable-video-diffusion-img2vid-xt/image_encoder
                conf = CLIPVisionConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32,
                                        num_attention_heads=16, image_size=224, patch_size=14,
                                        projection_dim=1024, hidden_act="gelu")
            else:                       # tiny test encoder
                conf = CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                        num_attention_heads=2, image_size=224, patch_size=32,
                                        projection_dim=embed_dim)
            rng = torch.random.get_rng_state()
            torch.manual_seed(seed)
            model = CLIPVisionModelWithProjection(conf)
            torch.random.set_rng_state(rng)
        self.model = model.to(self.device, dtype).eval()
        self.embed_dim = embed_dim

    @torch.no_grad()
    def __call__(self, video: torch.Tensor, chunk: int = 8) -> torch.Tensor:
        """video [T,3,H,W] in [-1,1] -> [T, embed_dim] float32."""
        x = resize_with_antialiasing(video.to(self.device).float(), (224, 224))
        x = (x + 1.0) / 2.0
        mean = torch.tensor(CLIP_MEAN, device=self.device).view(1, 3, 1, 1)
        std = torch.tensor(CLIP_STD, device=self.device).view(1, 3, 1, 1)
        x = ((x - mean) / std).to(self.dtype)
        outs = [self.model(pixel_values=x[i:i + chunk]).image_embeds for i in range(0, x.shape[0], chunk)]
        return torch.cat(outs, 0).float()
