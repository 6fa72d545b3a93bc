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


def _angle_deg(a, b):
    import numpy as np
    a = a / np.linalg.norm(a, axis=-1, keepdims=True)
    b = b / np.linalg.norm(b, axis=-1, keepdims=True)
    return np.degrees(np.arccos(np.clip((a * b).sum(-1), -1.0, 1.0)))


def test_depth_postprocess_kernel_matches_reference_golden(cuda):
    """ug_depth_postprocess (csrc/post.cu) against the output of the UNMODIFIED reference adapter
    (model/depthcrafter.py:92-97 + :48-69, fixture minted by tests/golden/make_golden.py): depth bit-exact,
    normals within 0.1 deg (the reference's own fp32 lstsq is not reproducible below ~0.04 deg)."""
    import os
    import numpy as np
    from unigeo_b200.config import tiny_config
    from unigeo_b200.engine import Engine
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "depthcrafter_post.npz"))
    e = Engine(tiny_config(), dtype="fp16", device=0)
    d, n = e.depth_postprocess(torch.from_numpy(g["frames"]), torch.from_numpy(g["intrinsics"]))
    torch.cuda.synchronize()
    assert torch.equal(d.cpu(), torch.from_numpy(g["pred_depths"]))
    ang = _angle_deg(n.cpu().numpy().astype(np.float64), g["pred_normals"].astype(np.float64))
    assert ang.max() <= 0.1 and np.median(ang) <= 0.01, (ang.max(), np.median(ang))
    d2, n2 = e.depth_postprocess(torch.from_numpy(g["frames"]), torch.from_numpy(g["intrinsics"]))
    assert torch.equal(d, d2) and torch.equal(n, n2)


def test_frames_io_kernels(cuda):
    """ug_vae_encode_frames / ug_vae_decode_frames == ug_vae_encode / ug_vae_decode_temporal + the torch glue
    of the upstream pipeline (x*2-1 + 0.02*noise; (x/2+0.5).clamp(0,1) -> [T,H,W,3]) -- bit-identical."""
    from unigeo_b200.config import tiny_config
    from unigeo_b200.engine import Engine
    from unigeo_b200.weights import synthetic_state_dict, vae_param_shapes
    cfg = tiny_config()
    e = Engine(cfg, dtype="fp16", device=0)
    e.load_state_dict("vae", synthetic_state_dict(vae_param_shapes(cfg.vae), 12))
    e.finalize()
    g = torch.Generator().manual_seed(3)
    T, H, W = 3, 64, 128
    frames = torch.rand(T, H, W, 3, generator=g).cuda()
    noise = torch.randn(T, 3, H, W, generator=g).cuda()
    video = frames.permute(0, 3, 1, 2).contiguous() * 2.0 - 1.0
    a = e.vae_encode(video, noise, 0.02)
    b, vid = e.vae_encode_frames(frames, noise, 0.02, want_video=True)
    assert torch.equal(vid, video)
    assert (a - b).abs().max().item() <= 2e-3 * a.abs().max().item()      # x*2-1+ns*n rounded once vs twice
    lat = torch.randn(T, 4, H // 8, W // 8, generator=g).cuda() * 0.5
    img = e.vae_decode(lat, 8)
    ref = (img / 2.0 + 0.5).clamp(0.0, 1.0).permute(0, 2, 3, 1).contiguous()
    got = e.vae_decode_frames(lat, 8)
    assert torch.equal(got, ref)


def test_prepare_input_device_is_bit_identical(cuda):
    """ug_prepare_frames == DepthCrafter.prepare_input (reference :39-45, uint8 truncation then /255)."""
    import numpy as np
    from unigeo_b200.model import DepthCrafter
    plug = DepthCrafter(config="tiny", dtype="fp16", weights="synthetic", clip="none")
    rng = np.random.default_rng(0)
    imgs = [(rng.random((3, 64, 128)) * 255.999).astype(np.float32) for _ in range(3)]
    imgs[0][:, 0, :4] = [0.0, 0.999, 254.5, 255.0]
    data = {"images": imgs}
    host = plug.prepare_input(data)
    dev = plug.prepare_input_device(data)
    assert np.array_equal(dev.cpu().numpy(), host)
