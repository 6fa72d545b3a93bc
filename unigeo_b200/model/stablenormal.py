"""``StableNormal`` plugin adapter (reference: /root/reference/model/stablenormal.py).

The reference pulls ``Stable-X/StableNormal`` through ``torch.hub.load`` (:16, network) and calls
``predictor(PIL image) -> PIL image`` per frame (:39).  Here the predictor is the B200 engine:
2-D VAE encode -> [YOSO one-step start] -> DDIM refinement with the SD-2.1 class UNet + ControlNet ->
2-D VAE decode -> 8-bit unit normals (``unigeo_b200/pipeline_stablenormal.py``, all frames of the clip
in one batch).  The adapter-side contract is reproduced exactly: per-frame 8-bit predictor output, the
uint8 x-flip wraparound (:43, App. B.10), ``/255*2-1`` (:45) and zero depths (:49).
``model_params`` (ride in **kwargs like the reference tolerates): ``config="full"|"tiny"``,
``dtype``, ``num_inference_steps=10``, ``seed``, ``weights=<dir>|"synthetic"``, ``controlnet=True``,
``yoso=False``, ``device=0``; or ``predictor=<callable PIL -> PIL>`` to inject one (the hub signature).
``weights`` defaults to ``model_dir`` (the reference's only key, configs/stablenormal_scannetpp.yaml:10-11); a missing
checkpoint directory raises FileNotFoundError like the reference's hub load does -- seeded random weights load ONLY on
an explicit ``weights="synthetic"``.
There is no CPU path: constructing this class without a B200 (and without ``predictor``) raises.
"""
from __future__ import annotations

import os

import numpy as np
import torch


class StableNormal:
    def __init__(self, model_dir=None, predictor=None, **kwargs):
        self.predictor = predictor
        self.pipeline = None
        if predictor is not None:
            return
        weights = kwargs.get("weights", model_dir)
        if weights != "synthetic" and not (weights and os.path.isdir(str(weights))):
            raise FileNotFoundError(f"StableNormal: weights={weights!r} is not a checkpoint directory "
                                    "(pass weights='synthetic' for seeded random weights)")
        from ..config import get_config, stablenormal_config
        from ..engine import Engine
        from ..pipeline_stablenormal import StableNormalPipelineB200 as P
        from ..weights import (controlnet_param_shapes, load_diffusers_dir, synthetic_state_dict,
                               unet2d_param_shapes, vae2d_param_shapes)
        name = kwargs.get("config", "full")
        self.sn_cfg = stablenormal_config(name)
        self.device = torch.device("cuda", int(kwargs.get("device", 0)))
        self.num_inference_steps = int(kwargs.get("num_inference_steps", self.sn_cfg.num_inference_steps))
        self.seed = kwargs.get("seed")
        use_ctrl, use_yoso = bool(kwargs.get("controlnet", True)), bool(kwargs.get("yoso", False))
        self.engine = Engine(get_config(name), dtype=kwargs.get("dtype", "fp16"), device=self.device.index,
                             sn_cfg=self.sn_cfg)
        nets = [(P.UNET, unet2d_param_shapes)] + ([(P.CONTROLNET, controlnet_param_shapes)] if use_ctrl else [])
        if use_yoso:
            nets += [(P.YOSO_UNET, unet2d_param_shapes)] + ([(P.YOSO_CONTROLNET, controlnet_param_shapes)] if use_ctrl else [])
        u = self.sn_cfg.unet2d
        if weights == "synthetic":
            s = int(kwargs.get("weight_seed", 0))
            wdev = self.device if kwargs.get("device_weights") else "cpu"
            wdt = torch.float16 if kwargs.get("device_weights") else torch.float32
            for i, (net, shapes) in enumerate(nets):
                self.engine.load_state_dict(net, synthetic_state_dict(shapes(u), 3000 + 10 * i + s, wdt, wdev))
            self.engine.load_state_dict("vae2d", synthetic_state_dict(vae2d_param_shapes(self.sn_cfg.vae2d), 4000 + s,
                                                                      wdt, wdev))
            g = torch.Generator().manual_seed(5000 + s)
            prompt = torch.randn(u.context_len, u.cross_attention_dim, generator=g)
        else:                                  # a directory laid out like the hub checkpoint
            for net, _ in nets:
                self.engine.load_state_dict(net, load_diffusers_dir(os.path.join(weights, net)))
            self.engine.load_state_dict("vae2d", load_diffusers_dir(os.path.join(weights, "vae")))
            prompt = torch.load(os.path.join(weights, "prompt_embeds.pt")).float().reshape(-1, u.cross_attention_dim)
        self.engine.finalize()
        self.pipeline = P(self.sn_cfg, self.engine, prompt, controlnet=use_ctrl, yoso=use_yoso)
        self.weight_source = "synthetic (seeded random)" if weights == "synthetic" else f"checkpoint directory {weights}"
        print(f"Using device: {self.device}; weights: {self.weight_source}")

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

    def forward(self, data, **debug_inputs):
        """``debug_inputs``: ``init_noise`` [F,4,h,w] to pin the random draw (parity tests)."""
        if self.predictor is not None:
            from PIL import Image
            images = [Image.fromarray(np.asarray(x).transpose(1, 2, 0).astype(np.uint8)) for x in data["images"]]
            preds = [np.array(self.predictor(im)) for im in images]
            return self.postprocess(preds)
        frames = np.stack([np.asarray(x).transpose(1, 2, 0).astype(np.uint8) for x in data["images"]], axis=0)
        gen = None
        if self.seed is not None:
            gen = torch.Generator(device=self.device).manual_seed(int(self.seed))
        normals = self.pipeline(frames, self.num_inference_steps, generator=gen, **debug_inputs)
        return self.postprocess(list(normals.cpu().numpy()))
