"""Parity at the BENCHMARK configuration (BASELINE cfg2: full SVD-XT UNet / temporal VAE, 25 x 384 x 512, fp16) against
the fp32 oracle RUN ON THE GPU (torch eager, TF32 off) on the SAME seeded weights and inputs.

The fp32 oracle cannot finish cfg2 on host cores in test time, but it is pure torch: on the B200 one full-size UNet step
takes about a second.  This puts an oracle comparison behind the code paths that only trigger at size (CTA-pair
256 x BN tiles at M = 76 800, N-fastest order, balanced unit lists, cluster GroupNorm, the TMA residual ring).
Tolerances are SURVEY.md §8(d)'s, fp32 oracle <-> fp16 kernels: UNet output per step rel-L2 <= 1e-2; decoded frames
max-abs <= 2e-2 on [0,1] (x2 margin for the random-weight VAE, as in test_model_gpu.py); |dAbs Rel| <= 1e-3,
|d delta| <= 2e-3, |d normal mean| <= 0.1 deg -- the yardstick being /root/reference/metrics/eval_depth.py:141-164 and
eval.py:47-56 through their bit-exact restatement (oracle/metrics.py, pinned by tests/golden/)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

T, H, W = 25, 384, 512
h, w = H // 8, W // 8


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def arms(cuda):
    """(cfg, engine, unet_sd, vae_sd): fp32 weights drawn once on the device, loaded into the engine (which stores them
    in fp16) and handed unchanged to the oracle."""
    from unigeo_b200.config import full_config
    from unigeo_b200.engine import Engine
    from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes, vae_param_shapes
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    # cuDNN OFF for the oracle: this image's cuDNN returns WRONG fp32 results for one of the decoder's convolutions
    # (N=8, 512 -> 256 channels at 96x128: output channels 254-255 of the left half of every row are off by O(1);
    # an fp64 unfold + matmul of the same inputs agrees with the tapgemm kernel to 2e-3 and disagrees with cuDNN by
    # 3.9 -- tools/bisect_conv.py, profiles/r02_cudnn_fp32_conv_mismatch.txt).  torch's native im2col + cuBLAS path
    # is slower but correct.
    cudnn_was = torch.backends.cudnn.enabled
    torch.backends.cudnn.enabled = False
    cfg = full_config()
    usd = synthetic_state_dict(unet_param_shapes(cfg.unet), 1000, torch.float32, cuda)
    vsd = synthetic_state_dict(vae_param_shapes(cfg.vae), 2000, torch.float32, cuda)
    e = Engine(cfg, dtype="fp16", device=0)
    e.load_state_dict("unet", usd)
    e.load_state_dict("vae", vsd)
    e.finalize()
    yield cfg, e, usd, vsd
    e.close()
    torch.backends.cudnn.enabled = cudnn_was
    torch.cuda.empty_cache()


def test_cfg2_unet_step_matches_fp32_oracle(arms, cuda):
    from oracle.pipeline import added_time_ids
    from oracle.unet_st import unet_forward
    cfg, e, usd, _ = arms
    g = torch.Generator(device=cuda).manual_seed(21)
    x = torch.randn(1, T, 8, h, w, generator=g, device=cuda)
    enc = torch.randn(1, T, cfg.clip_embed_dim, generator=g, device=cuda)
    ids = added_time_ids(cfg, cuda)
    with torch.no_grad():
        ref = unet_forward(usd, cfg.unet, x, 0.9, enc, ids)
    e.prepare(T, h, w)
    e.set_clip_context(enc[0])
    got = e.unet_forward(x, 0.9, ids[0].tolist())
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    err = rel_l2(got, ref)
    print("cfg2 UNet step rel-L2", err)
    assert err <= 1e-2, err
    # per-frame: no frame may hide behind the average (tile-order / frame-bias bugs are per frame)
    per_frame = [rel_l2(got[0, t], ref[0, t]) for t in range(T)]
    assert max(per_frame) <= 2e-2, per_frame


def test_cfg2_vae_decode_chunk_and_encode_match_fp32_oracle(arms, cuda):
    from oracle.vae import vae_decode, vae_encode
    cfg, e, _, vsd = arms
    g = torch.Generator(device=cuda).manual_seed(22)
    lat = torch.randn(9, 4, h, w, generator=g, device=cuda) * 0.5           # one full chunk of 8 + a ragged chunk of 1
    with torch.no_grad():
        ref = vae_decode(vsd, cfg.vae, lat, 8)
    got = e.vae_decode(lat, chunk=8)
    torch.cuda.synchronize()
    err, mx = rel_l2(got, ref), (got - ref).abs().max().item() / 2.0        # /2: [-1,1] image -> [0,1] frames
    print("cfg2 VAE decode rel-L2", err, "max-abs on [0,1]", mx)
    assert err <= 1e-2 and mx <= 4e-2, (err, mx)
    img = torch.rand(4, 3, H, W, generator=g, device=cuda) * 2 - 1
    with torch.no_grad():
        ref_l = vae_encode(vsd, cfg.vae, img)
    got_l = e.vae_encode(img)
    torch.cuda.synchronize()
    err = rel_l2(got_l, ref_l)
    print("cfg2 VAE encode rel-L2", err)
    assert err <= 1e-2, err


def test_cfg2_plugin_abs_rel_parity(arms, cuda):
    """The reference's shipped call (5 steps, model/depthcrafter.py:86) at cfg2 through the plugin adapter, both arms on
    identical frames / noise draws / CLIP embeddings; scored by the reference metric restatement against the synthetic
    scene's ground truth."""
    from harness.synthetic import gt_label, make_clip
    from oracle import metrics as OM
    from oracle import postprocess as OP
    from oracle.pipeline import depthcrafter_pipeline
    from unigeo_b200.model.depthcrafter import DepthCrafter
    from unigeo_b200.pipeline import DepthCrafterPipelineB200
    cfg, e, usd, vsd = arms
    steps = 5
    data = make_clip(T, H, W, seed=77)
    plug = object.__new__(DepthCrafter)
    plug.device, plug.cfg, plug.dtype, plug.engine = e.device, cfg, "fp16", e
    plug.num_inference_steps, plug.seed, plug._stage = steps, None, None
    plug.pipeline = DepthCrafterPipelineB200(cfg, e, None)
    g = torch.Generator(device=cuda).manual_seed(5)
    enc = torch.randn(T, cfg.clip_embed_dim, generator=g, device=cuda)
    aug = torch.randn(T, 3, H, W, generator=g, device=cuda)
    init = torch.randn(T, 4, h, w, generator=g, device=cuda)
    out = plug.forward(data, enc=enc, aug_noise=aug, init_noise=init)

    frames = torch.from_numpy(OP.prepare_input(data["images"])).to(cuda)
    with torch.no_grad():
        ref_frames = depthcrafter_pipeline(usd, vsd, cfg, frames, enc[None], aug, init[None], steps)
    ref_frames = ref_frames.cpu().numpy()
    ref_depth = OP.disparity_to_depth(ref_frames)
    # the reference's plane-fit normals are 22 s per clip on host cores; depth parity is scored on every
    # frame, normal-mean parity on the first 3 frames through the reference restatement
    nf = 3
    ref = OP.prepare_output(ref_depth[:nf], data["intrinsics"][:nf])
    gt = gt_label(data)
    m_ref = OM.depth_evaluation(torch.from_numpy(np.asarray(ref_depth, dtype=np.float32)), gt["gt_depths"], gt["gt_masks"])
    m_got = OM.depth_evaluation(out["pred_depths"], gt["gt_depths"], gt["gt_masks"])
    n_ref = OM.normal_evaluation(ref["pred_normals"], gt["gt_normals"][:nf], gt["gt_masks"][:nf])
    n_got = OM.normal_evaluation(out["pred_normals"][:nf], gt["gt_normals"][:nf], gt["gt_masks"][:nf])
    dmax = (out["pred_depths"] - torch.from_numpy(np.asarray(ref_depth, dtype=np.float32))).abs().max().item()
    print("cfg2 depth ref", m_ref, "\ncfg2 depth got", m_got, "\nnormal mean ref/got", n_ref["normal mean"],
          n_got["normal mean"], "max |d depth|", dmax)
    assert abs(m_ref["Abs Rel"] - m_got["Abs Rel"]) <= 1e-3
    for k in ("delta < 1.25", "delta < 1.25^2", "delta < 1.25^3"):
        assert abs(m_ref[k] - m_got[k]) <= 2e-3, k
    assert m_ref["valid_pixels"] == m_got["valid_pixels"]
    assert abs(n_ref["normal mean"] - n_got["normal mean"]) <= 0.1
    # the same two arms against a label that depends on the prediction (harness.synthetic.correlated_gt): with randomly
    # initialised weights the scene label above aligns both arms to the same constant, this one does not
    from harness.synthetic import correlated_gt
    ref_t = torch.from_numpy(np.asarray(ref_depth, dtype=np.float32))
    cg = correlated_gt(ref_t)
    s_ref = OM.depth_evaluation(ref_t, cg["gt_depths"], cg["gt_masks"])
    s_got = OM.depth_evaluation(out["pred_depths"], cg["gt_depths"], cg["gt_masks"])
    print("cfg2 correlated-label ref", s_ref, "\ncfg2 correlated-label got", s_got)
    assert 0.01 < s_ref["Abs Rel"] < 0.2, s_ref                 # the label is neither trivial nor unrelated
    assert abs(s_ref["Abs Rel"] - s_got["Abs Rel"]) <= 1e-3, (s_ref["Abs Rel"], s_got["Abs Rel"])
    for k in ("delta < 1.25", "delta < 1.25^2", "delta < 1.25^3"):
        assert abs(s_ref[k] - s_got[k]) <= 2e-3, k


def test_vae_encoder_in_bf16_survives_activations_beyond_the_fp16_range(cuda):
    """upstream upcasts the VAE encoder to fp32 because real SVD activations leave the fp16 range.  With
    the first conv's weights scaled so that activations reach ~3e5 (> 65504), the fp16 encoder overflows while the
    bf16 encoder (ug_ctx_set_vae_encode_dtype; same tensor-core path, fp32 exponent range) stays finite and inside the
    bf16 tolerance of the fp32 oracle.  GroupNorm right after the conv brings the scale back, so the oracle's output is
    an ordinary latent."""
    from oracle.vae import vae_encode
    from unigeo_b200.config import tiny_config
    from unigeo_b200.engine import Engine
    from unigeo_b200.weights import synthetic_state_dict, vae_param_shapes
    cfg = tiny_config()
    vsd = synthetic_state_dict(vae_param_shapes(cfg.vae), 12)
    vsd["encoder.conv_in.weight"] = vsd["encoder.conv_in.weight"] * 4e5
    g = torch.Generator().manual_seed(3)
    img = torch.rand(2, 3, 64, 128, generator=g) * 2 - 1
    with torch.no_grad():
        ref = vae_encode(vsd, cfg.vae, img)
    assert torch.isfinite(ref).all()
    outs = {}
    for enc_dt in (None, "bf16"):
        e = Engine(cfg, dtype="fp16", device=0, vae_encode_dtype=enc_dt)
        e.load_state_dict("vae", vsd)
        outs[enc_dt] = e.vae_encode(img).cpu()
        e.close()
    assert not torch.isfinite(outs[None]).all()                  # the fp16 encoder does overflow on this input
    assert torch.isfinite(outs["bf16"]).all()
    assert rel_l2(outs["bf16"], ref) <= 3e-2, rel_l2(outs["bf16"], ref)
