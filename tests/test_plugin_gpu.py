"""End-to-end parity at the plugin boundary: unigeo_b200.model.DepthCrafter.forward(data) against the
oracle pipeline + the oracle's restatement of the reference adapter/metrics, on identical seeded
inputs (synthetic Unified-Data-Format clip, tiny config).  Tolerances from SURVEY.md §8(d):
|dAbs Rel| <= 1e-3, |d delta| <= 2e-3; decoded-depth driven normals within 0.5 deg of mean error."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

T, H, W, STEPS = 5, 128, 256, 3


def test_depthcrafter_plugin_matches_oracle(cuda):
    from harness.synthetic import gt_label, make_clip
    from oracle import metrics as OM
    from oracle import postprocess as OP
    from oracle.pipeline import depthcrafter_pipeline
    from unigeo_b200.model import DepthCrafter
    from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes, vae_param_shapes

    data = make_clip(T, H, W, seed=77)
    plug = DepthCrafter(config="tiny", dtype="fp16", weights="synthetic", num_inference_steps=STEPS, clip="none")
    cfg = plug.cfg
    g = torch.Generator().manual_seed(5)
    enc = torch.randn(T, cfg.clip_embed_dim, generator=g)
    aug = torch.randn(T, 3, H, W, generator=g)
    init = torch.randn(T, 4, H // 8, W // 8, generator=g)
    out = plug.forward(data, enc=enc, aug_noise=aug, init_noise=init)
    assert out["pred_depths"].shape == (T, H, W) and out["pred_depths"].dtype == torch.float32
    assert out["pred_normals"].shape == (T, H, W, 3) and not out["pred_depths"].is_cuda
    assert plug.engine.launch_count() > 0

    usd = synthetic_state_dict(unet_param_shapes(cfg.unet), 1000)
    vsd = synthetic_state_dict(vae_param_shapes(cfg.vae), 2000)
    frames = torch.from_numpy(OP.prepare_input(data["images"]))
    with torch.no_grad():
        ref_frames = depthcrafter_pipeline(usd, vsd, cfg, frames, enc[None], aug, init[None], STEPS).numpy()
    ref_depth = OP.disparity_to_depth(ref_frames)
    ref = OP.prepare_output(ref_depth, data["intrinsics"])

    gt = gt_label(data)
    m_ref = OM.depth_evaluation(ref["pred_depths"], gt["gt_depths"], gt["gt_masks"])
    m_got = OM.depth_evaluation(out["pred_depths"], gt["gt_depths"], gt["gt_masks"])
    n_ref = OM.normal_evaluation(ref["pred_normals"], gt["gt_normals"], gt["gt_masks"])
    n_got = OM.normal_evaluation(out["pred_normals"], gt["gt_normals"], gt["gt_masks"])
    print("depth ref", m_ref, "\ndepth got", m_got, "\nnormal ref", n_ref, "\nnormal got", n_got)
    rel = (out["pred_depths"] - ref["pred_depths"]).abs().max().item()
    print("max |d depth|", rel)
    assert abs(m_ref["Abs Rel"] - m_got["Abs Rel"]) <= 1e-3
    for k in ("delta < 1.25", "delta < 1.25^2", "delta < 1.25^3"):
        assert abs(m_ref[k] - m_got[k]) <= 2e-3, k
    assert m_ref["valid_pixels"] == m_got["valid_pixels"]
    assert abs(n_ref["normal mean"] - n_got["normal mean"]) <= 0.5


def test_plugin_is_deterministic_with_seed(cuda):
    from harness.synthetic import make_clip
    from unigeo_b200.model import DepthCrafter
    data = make_clip(3, 128, 256, seed=3)
    plug = DepthCrafter(config="tiny", dtype="bf16", weights="synthetic", num_inference_steps=2, seed=11)
    a = plug.forward(data)
    b = plug.forward(data)
    assert torch.equal(a["pred_depths"], b["pred_depths"]) and torch.equal(a["pred_normals"], b["pred_normals"])
    assert a["pred_depths"].min() >= 1 / 1.1 - 1e-5 and a["pred_depths"].max() <= 10 + 1e-4
