"""``StableNormal`` plugin adapter (reference: /root/reference/model/stablenormal.py).

The reference pulls ``Stable-X/StableNormal`` through ``torch.hub.load`` (:16, network) and its
arithmetic (YOSO init + SD-2.1 UNet/ControlNet + DINOv2) is not available in
this environment in any form.  What IS in-tree and reproduced here exactly is the adapter-side
contract: per-frame 8-bit predictor output, the uint8 x-flip wraparound (:43, App. B.10),
``/255*2-1`` (:45) and zero depths (:49).  The predictor is injectable: any callable
``PIL.Image -> PIL.Image`` (the hub predictor's signature, :39).  The 2-D UNet kernels it needs
are the spatial half of the DepthCrafter path (conv3x3 / GroupNorm / attention / GEGLU in
libunigeo_b200.so); wiring a 2-D SD UNet graph over them is listed as "next" in DESIGN.md.
"""
from __future__ import annotations

import numpy as np
import torch


class StableNormal:
    def __init__(self, model_dir=None, predictor=None, **kwargs):
        if predictor is None:
            raise RuntimeError(
                "StableNormal needs a predictor (PIL.Image -> PIL.Image); the reference obtains it with "
                "torch.hub.load('Stable-X/StableNormal', ...) which is unavailable offline")
        self.predictor = predictor

    def prepare_input(self, data):
        frames = [np.asarray(x).transpose(1, 2, 0).astype(np.uint8) for x in data["images"]]
        return np.stack(frames, axis=0).astype(np.float32) / 255.0

    @staticmethod
    def postprocess(normals_u8):
        """reference :41-50 on a list of uint8 [H,W,3] arrays."""
        out = []
        for n in normals_u8:
            n = np.array(n, dtype=np.uint8)
            n[:, :, 0] = -n[:, :, 0]                      # uint8 negation wraps: v -> (256 - v) % 256
            out.append(torch.from_numpy(n / 255.0 * 2 - 1).float())
        pn = torch.stack(out, dim=0)
        return {"pred_normals": pn, "pred_depths": torch.zeros_like(pn[..., 0])}

    def forward(self, data):
        from PIL import Image
        images = [Image.fromarray(np.asarray(x).transpose(1, 2, 0).astype(np.uint8)) for x in data["images"]]
        preds = [np.array(self.predictor(im)) for im in images]
        return self.postprocess(preds)
