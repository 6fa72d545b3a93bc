"""BASELINE cfg3: StableNormal on ONE 512 x 512 frame, 10 DDIM refinement steps, the full SD-2.1 class UNet2D +
ControlNet (865.9 M + 363.1 M parameters) and the 2-D VAE, fp16 kernels against the fp32 oracle restatement RUN ON
THE GPU (torch eager, TF32 off, cuDNN off -- see tests/test_parity_cfg2_gpu.py) on the same seeded weights, prompt
embeddings and start noise.  Normal-mean parity: the two 8-bit normal maps go through the reference adapter's
post-processing (model/stablenormal.py:41-50) and the reference metric restatement (metrics/eval_normal.py) against a
synthetic ground truth; |d normal mean| <= 0.1 deg, mean angle between the two predictions <= 1 deg
(the 8-bit quantisation of the predictor output alone is ~0.3 deg)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H = W = 512
STEPS = 10


def angular_deg(a, b):
    a = a / a.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    b = b / b.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    return torch.rad2deg(torch.acos((a * b).sum(-1).clamp(-1, 1)))


def test_cfg3_normal_mean_parity(cuda):
    from harness.synthetic import gt_label, make_clip
    from oracle import metrics as OM
    from oracle.stablenormal import normals_to_u8, refine
    from oracle.vae import vae_decode_2d, vae_encode
    from unigeo_b200.config import get_config, stablenormal_config
    from unigeo_b200.engine import Engine
    from unigeo_b200.model import StableNormal
    from unigeo_b200.pipeline_stablenormal import StableNormalPipelineB200 as P
    from unigeo_b200.weights import (controlnet_param_shapes, synthetic_state_dict, unet2d_param_shapes,
                                     vae2d_param_shapes)
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.enabled)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.enabled = False
    try:
        sn = stablenormal_config("full")
        u = sn.unet2d
        usd = synthetic_state_dict(unet2d_param_shapes(u), 3000, torch.float32, cuda)
        csd = synthetic_state_dict(controlnet_param_shapes(u), 3010, torch.float32, cuda)
        vsd = synthetic_state_dict(vae2d_param_shapes(sn.vae2d), 4000, torch.float32, cuda)
        g = torch.Generator(device=cuda).manual_seed(5000)
        prompt = torch.randn(u.context_len, u.cross_attention_dim, generator=g, device=cuda)
        noise = torch.randn(1, 4, H // 8, W // 8, generator=g, device=cuda)
        data = make_clip(1, H, W, seed=31)
        u8 = np.stack([np.asarray(x).transpose(1, 2, 0).astype(np.uint8) for x in data["images"]], 0)

        e = Engine(get_config("full"), dtype="fp16", device=0, sn_cfg=sn)
        e.load_state_dict(P.UNET, usd)
        e.load_state_dict(P.CONTROLNET, csd)
        e.load_state_dict("vae2d", vsd)
        e.finalize()
        pipe = P(sn, e, prompt, controlnet=True, yoso=False)
        got_u8 = pipe(u8, STEPS, init_noise=noise).cpu().numpy()
        launches = e.launch_count()
        e.close()

        with torch.no_grad():
            img = torch.from_numpy(u8.astype(np.float32) / 255.0).to(cuda).permute(0, 3, 1, 2) * 2.0 - 1.0
            il = vae_encode(vsd, sn.vae2d, img) * sn.vae2d.scaling_factor
            lat = refine(usd, csd, sn, il, prompt[None], noise, STEPS)
            ref_u8 = normals_to_u8(vae_decode_2d(vsd, sn.vae2d, lat))
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.enabled = saved
    assert launches > 0 and got_u8.shape == ref_u8.shape == (1, H, W, 3)
    got = StableNormal.postprocess(list(got_u8))["pred_normals"]
    ref = StableNormal.postprocess(list(ref_u8))["pred_normals"]
    ang = angular_deg(got, ref).mean().item()
    gt = gt_label(data)
    n_got = OM.normal_evaluation(got, gt["gt_normals"], gt["gt_masks"])
    n_ref = OM.normal_evaluation(ref, gt["gt_normals"], gt["gt_masks"])
    print("cfg3 mean angle between arms", ang, "normal mean b200 / oracle", n_got["normal mean"], n_ref["normal mean"])
    assert ang <= 1.0, ang
    assert abs(n_got["normal mean"] - n_ref["normal mean"]) <= 0.1
    assert abs(n_got["normal median"] - n_ref["normal median"]) <= 0.3
