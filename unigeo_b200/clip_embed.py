"""Per-frame CLIP image embeddings (pipeline step 3) on the B200 engine.

``ug_clip_embed`` (csrc/clip.cu) runs the antialiased 224x224 resize, the CLIP normalisation and the ViT +
projection as CUDA kernels; this class only owns the weights' way in.  No checkpoint exists in this
environment, so by default the weights are seeded random-init of the configured architecture (ViT-H/14 for
the full model); a real ``image_encoder`` directory (safetensors, transformers key names) can be passed instead.
There is no torch / CPU path: the embedder needs the engine.
"""
from __future__ import annotations

import torch

from .weights import clip_param_shapes, load_diffusers_dir, synthetic_state_dict


class ClipEmbedder:
    def __init__(self, engine, seed: int = 7, pretrained: str | None = None, device_weights: bool = False):
        self.engine = engine
        cfg = engine.cfg.clip
        if pretrained:
            sd = load_diffusers_dir(pretrained)
        else:
            wdev = engine.device if device_weights else "cpu"
            wdt = torch.float16 if device_weights else torch.float32
            sd = synthetic_state_dict(clip_param_shapes(cfg), seed, wdt, wdev)
        sd = {k: v for k, v in sd.items() if "position_ids" not in k}
        engine.load_state_dict("clip", sd)
        self.embed_dim = cfg.projection_dim

    @torch.no_grad()
    def __call__(self, video: torch.Tensor) -> torch.Tensor:
        """video [T,3,H,W] in [-1,1] -> [T, embed_dim] float32 (CUDA)."""
        return self.engine.clip_embed(video)
