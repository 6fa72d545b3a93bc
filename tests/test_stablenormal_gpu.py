"""StableNormal path (BASELINE cfg3) through the C ABI: the 2-D UNet / ControlNet,
the DDIM refinement loop, the 2-D VAE and the plugin adapter against the fp32 oracle restatement
(oracle/unet_2d.py, oracle/stablenormal.py) on identical seeded weights and inputs, tiny config.
Tolerances: per-op rel-L2 <= 5e-3 (fp16) / 2e-2 (bf16); UNet output <= 1e-2 / 3e-2;
decoded images max-abs <= 2e-2 / 5e-2 (x2 margin for the random-weight VAE, as in test_model_gpu.py);
normal-mean difference <= 0.1 deg is the BASELINE bar, asserted on the 8-bit maps at <= 1 deg because
8-bit quantisation alone moves a normal by up to 0.6 deg."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

Fr, H, W = 2, 128, 256           # latent 16x32: every UNet level keeps (tokens % 8 == 0)
OP_TOL = {torch.float16: 5e-3, torch.bfloat16: 2e-2}
UNET_TOL = {"fp16": 1e-2, "bf16": 3e-2}
IMG_TOL = {"fp16": 2e-2, "bf16": 5e-2}


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("F_,N,C,Lk,per_frame", [(2, 512, 128, 77, False), (3, 100, 64, 13, True),
                                                 (1, 64, 320, 128, False), (2, 8, 64, 1, True)])
def test_cross_attention_op(cuda, dtype, F_, N, C, Lk, per_frame):
    from unigeo_b200 import ops
    g = torch.Generator().manual_seed(F_ * 1000 + N + Lk)
    Fk = F_ if per_frame else 1
    q = torch.randn(F_ * N, C, generator=g).to(dtype)
    kv = torch.randn(Fk * Lk, 2 * C, generator=g).to(dtype)
    got = ops.cross_attention(q.cuda(), kv.cuda(), F_, N, C, Lk, per_frame)
    heads = C // 64
    qf = q.float().view(F_, N, heads, 64).transpose(1, 2)
    k = kv.float()[:, :C].reshape(Fk, Lk, heads, 64).transpose(1, 2).expand(F_, -1, -1, -1)
    v = kv.float()[:, C:].reshape(Fk, Lk, heads, 64).transpose(1, 2).expand(F_, -1, -1, -1)
    p = torch.softmax(qf @ k.transpose(-1, -2) / 8.0, dim=-1)
    ref = (p @ v).transpose(1, 2).reshape(F_ * N, C)
    assert rel_l2(got, ref) <= OP_TOL[dtype]


@pytest.fixture(scope="module")
def bundle(cuda):
    from unigeo_b200.config import stablenormal_config, tiny_config
    from unigeo_b200.weights import (controlnet_param_shapes, synthetic_state_dict, unet2d_param_shapes,
                                     vae2d_param_shapes)
    sn = stablenormal_config("tiny")
    usd = synthetic_state_dict(unet2d_param_shapes(sn.unet2d), 21)
    csd = synthetic_state_dict(controlnet_param_shapes(sn.unet2d), 22)
    vsd = synthetic_state_dict(vae2d_param_shapes(sn.vae2d), 23)
    g = torch.Generator().manual_seed(4321)
    h, w = H // 8, W // 8
    d = dict(
        x=torch.randn(Fr, 4, h, w, generator=g),
        il=torch.randn(Fr, 4, h, w, generator=g),
        ctx=torch.randn(1, sn.unet2d.context_len, sn.unet2d.cross_attention_dim, generator=g),
        ctx_pf=torch.randn(Fr, sn.unet2d.context_len, sn.unet2d.cross_attention_dim, generator=g),
        img=torch.rand(Fr, 3, H, W, generator=g) * 2 - 1,
        lat=torch.randn(Fr, 4, h, w, generator=g) * 0.18215 * 3,
        u8=(torch.rand(Fr, H, W, 3, generator=g) * 255).numpy().astype(np.uint8),
    )
    return tiny_config(), sn, usd, csd, vsd, d


def make_engine(cfg, sn, usd, csd, vsd, dtype):
    from unigeo_b200.engine import Engine
    e = Engine(cfg, dtype=dtype, device=0, sn_cfg=sn)
    e.load_state_dict("unet2d", usd)
    e.load_state_dict("controlnet", csd)
    e.load_state_dict("vae2d", vsd)
    e.finalize()
    return e


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
@pytest.mark.parametrize("per_frame_ctx", [False, True])
def test_unet2d_forward(bundle, dtype, per_frame_ctx):
    from oracle.unet_2d import unet2d_forward
    cfg, sn, usd, csd, vsd, d = bundle
    ctx = d["ctx_pf"] if per_frame_ctx else d["ctx"]
    with torch.no_grad():
        ref = unet2d_forward(usd, sn.unet2d, d["x"], 431.0, ctx)
    e = make_engine(cfg, sn, usd, csd, vsd, dtype)
    e.set_text_context("unet2d", ctx)
    got = e.unet2d_forward("unet2d", d["x"], 431.0)
    torch.cuda.synchronize()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert rel_l2(got, ref) <= UNET_TOL[dtype]
    again = e.unet2d_forward("unet2d", d["x"], 431.0)          # no atomics anywhere: reruns are bit-identical
    assert torch.equal(got, again)


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_unet2d_with_controlnet(bundle, dtype):
    from oracle.unet_2d import controlnet_forward, unet2d_forward
    cfg, sn, usd, csd, vsd, d = bundle
    with torch.no_grad():
        down, mid = controlnet_forward(csd, sn.unet2d, d["il"], 77.0, d["ctx"])
        ref = unet2d_forward(usd, sn.unet2d, d["x"], 77.0, d["ctx"], down, mid)
        plain = unet2d_forward(usd, sn.unet2d, d["x"], 77.0, d["ctx"])
    e = make_engine(cfg, sn, usd, csd, vsd, dtype)
    e.set_text_context("unet2d", d["ctx"])
    e.set_text_context("controlnet", d["ctx"])
    got = e.unet2d_forward("unet2d", d["x"], 77.0, "controlnet", d["il"])
    torch.cuda.synchronize()
    assert rel_l2(got, ref) <= UNET_TOL[dtype]
    assert rel_l2(plain, ref) > 10 * UNET_TOL[dtype]            # the ControlNet residuals matter in this test


def test_refine_loop(bundle):
    from oracle.stablenormal import refine
    cfg, sn, usd, csd, vsd, d = bundle
    trace = []
    with torch.no_grad():
        ref = refine(usd, csd, sn, d["il"], d["ctx"], d["x"], 3, trace=trace)
    e = make_engine(cfg, sn, usd, csd, vsd, "fp16")
    e.set_text_context("unet2d", d["ctx"])
    e.set_text_context("controlnet", d["ctx"])
    got = e.refine_2d("unet2d", "controlnet", d["il"], d["x"], 3)
    torch.cuda.synchronize()
    assert rel_l2(got, ref) <= 2e-2
    one = e.refine_2d("unet2d", "controlnet", d["il"], d["x"], 1)      # 1 step == its own x0 (a_prev = 1)
    with torch.no_grad():
        ref1 = refine(usd, csd, sn, d["il"], d["ctx"], d["x"], 1)
    assert rel_l2(one, ref1) <= UNET_TOL["fp16"]
    assert e.launch_count() > 0


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_vae2d(bundle, dtype):
    from oracle.stablenormal import normals_to_u8
    from oracle.vae import vae_decode_2d, vae_encode
    cfg, sn, usd, csd, vsd, d = bundle
    with torch.no_grad():
        ref_lat = vae_encode(vsd, sn.vae2d, d["img"]) * sn.vae2d.scaling_factor
        ref_img = vae_decode_2d(vsd, sn.vae2d, d["lat"])
    e = make_engine(cfg, sn, usd, csd, vsd, dtype)
    lat = e.vae2d_encode(d["img"], sn.vae2d.scaling_factor)
    img, nrm = e.vae2d_decode(d["lat"], want_image=True, want_normals_u8=True)
    torch.cuda.synchronize()
    assert rel_l2(lat, ref_lat) <= UNET_TOL[dtype]
    assert (img.cpu() - ref_img).abs().max().item() <= 2 * IMG_TOL[dtype] * max(1.0, ref_img.abs().max().item())
    # the fused 8-bit output is exactly the oracle's mapping applied to the kernel's own decoded image
    own = normals_to_u8(img.cpu())
    diff = np.abs(own.astype(np.int16) - nrm.cpu().numpy().astype(np.int16))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.02         # 16-bit rounding of the stored image only


def angular_deg(a, b):
    a = a / a.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    b = b / b.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    return torch.rad2deg(torch.acos((a * b).sum(-1).clamp(-1, 1)))


def test_plugin_end_to_end(bundle):
    """StableNormal.forward(data) (tiny config, synthetic weights) against the oracle predictor + the
    reference adapter's post-processing (x flip wraparound, /255*2-1, zero depths)."""
    from oracle.stablenormal import stablenormal_predict
    from unigeo_b200.model import StableNormal
    from unigeo_b200.weights import (controlnet_param_shapes, synthetic_state_dict, unet2d_param_shapes,
                                     vae2d_param_shapes)
    cfg, sn, _, _, _, d = bundle
    plug = StableNormal(config="tiny", weights="synthetic", num_inference_steps=2, weight_seed=0)
    data = {"images": [f.transpose(2, 0, 1).astype(np.float32) for f in d["u8"]]}
    noise = torch.randn(Fr, 4, H // 8, W // 8, generator=torch.Generator().manual_seed(9))
    out = plug.forward(data, init_noise=noise)
    assert out["pred_normals"].shape == (Fr, H, W, 3) and out["pred_normals"].dtype == torch.float32
    assert out["pred_normals"].device.type == "cpu"
    assert out["pred_depths"].shape == (Fr, H, W) and not out["pred_depths"].any()
    usd = synthetic_state_dict(unet2d_param_shapes(sn.unet2d), 3000)
    csd = synthetic_state_dict(controlnet_param_shapes(sn.unet2d), 3010)
    vsd = synthetic_state_dict(vae2d_param_shapes(sn.vae2d), 4000)
    prompt = torch.randn(sn.unet2d.context_len, sn.unet2d.cross_attention_dim,
                         generator=torch.Generator().manual_seed(5000))
    with torch.no_grad():
        ref_u8 = stablenormal_predict(usd, csd, vsd, sn, d["u8"], prompt[None], noise, 2)
    ref = StableNormal.postprocess(list(ref_u8))["pred_normals"]
    ang = angular_deg(out["pred_normals"], ref)
    assert ang.mean().item() <= 1.0, ang.mean().item()


def test_missing_text_context_is_an_error(bundle):
    from unigeo_b200._lib import UgError
    cfg, sn, usd, csd, vsd, d = bundle
    e = make_engine(cfg, sn, usd, csd, vsd, "fp16")
    with pytest.raises(UgError, match="ug_set_text_context"):
        e.unet2d_forward("unet2d", d["x"], 10.0)
    with pytest.raises(UgError, match="no 2-D network"):
        e.unet2d_forward("nope", d["x"], 10.0)


def test_refine_loop_graph_replay_is_bit_identical(bundle):
    """The refinement loop is captured into a CUDA graph on its second call with the same signature and replayed
    afterwards: eager (call 1), capture + launch (call 2) and replay (call 3) must agree bit for bit, also when
    the inputs change between replays (external pointers stay outside the graph)."""
    cfg, sn, usd, csd, vsd, d = bundle
    e = make_engine(cfg, sn, usd, csd, vsd, "fp16")
    e.set_text_context("unet2d", d["ctx"])
    e.set_text_context("controlnet", d["ctx"])
    n0 = e.launch_count()
    a = e.refine_2d("unet2d", "controlnet", d["il"], d["x"], 2)
    n1 = e.launch_count()
    b = e.refine_2d("unet2d", "controlnet", d["il"], d["x"], 2)
    c = e.refine_2d("unet2d", "controlnet", d["il"].clone(), d["x"].clone(), 2)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(a, c)
    import os
    assert e.graph_count() == (0 if os.environ.get("UG_NO_GRAPH") else 1)
    other = e.refine_2d("unet2d", "controlnet", d["il"] * 0.5, d["x"] + 1.0, 2)      # replay on new inputs
    e2 = make_engine(cfg, sn, usd, csd, vsd, "fp16")
    e2.set_text_context("unet2d", d["ctx"])
    e2.set_text_context("controlnet", d["ctx"])
    ref = e2.refine_2d("unet2d", "controlnet", d["il"] * 0.5, d["x"] + 1.0, 2)       # eager in a fresh context
    assert torch.equal(other, ref)
    assert e.launch_count() - n0 == 4 * (n1 - n0)           # replays are counted like the launches they contain


def test_plugin_with_yoso_start_and_checkpoint_directory(bundle, tmp_path):
    """(1) yoso=True: the one-step initialiser (its own UNet + ControlNet) provides the start latent; (2) the same
    weights written as a hub-style checkpoint directory (<dir>/<net>/*.safetensors, <dir>/vae, prompt_embeds.pt) and
    loaded through ``weights=<dir>`` give bit-identical outputs; both against the oracle predictor."""
    import os
    from safetensors.torch import save_file
    from oracle.stablenormal import stablenormal_predict
    from unigeo_b200.model import StableNormal
    from unigeo_b200.weights import (controlnet_param_shapes, synthetic_state_dict, unet2d_param_shapes,
                                     vae2d_param_shapes)
    cfg, sn, _, _, _, d = bundle
    data = {"images": [f.transpose(2, 0, 1).astype(np.float32) for f in d["u8"]]}
    noise = torch.randn(Fr, 4, H // 8, W // 8, generator=torch.Generator().manual_seed(10))
    plug = StableNormal(config="tiny", weights="synthetic", num_inference_steps=2, weight_seed=0, yoso=True)
    out = plug.forward(data, init_noise=noise)
    sds = {"unet2d": synthetic_state_dict(unet2d_param_shapes(sn.unet2d), 3000),
           "controlnet": synthetic_state_dict(controlnet_param_shapes(sn.unet2d), 3010),
           "yoso_unet": synthetic_state_dict(unet2d_param_shapes(sn.unet2d), 3020),
           "yoso_controlnet": synthetic_state_dict(controlnet_param_shapes(sn.unet2d), 3030)}
    vsd = synthetic_state_dict(vae2d_param_shapes(sn.vae2d), 4000)
    prompt = torch.randn(sn.unet2d.context_len, sn.unet2d.cross_attention_dim,
                         generator=torch.Generator().manual_seed(5000))
    with torch.no_grad():
        ref_u8 = stablenormal_predict(sds["unet2d"], sds["controlnet"], vsd, sn, d["u8"], prompt[None], noise, 2,
                                      yoso_unet_sd=sds["yoso_unet"], yoso_ctrl_sd=sds["yoso_controlnet"])
    ref = StableNormal.postprocess(list(ref_u8))["pred_normals"]
    assert angular_deg(out["pred_normals"], ref).mean().item() <= 1.0
    plain = StableNormal(config="tiny", weights="synthetic", num_inference_steps=2, weight_seed=0).forward(data, init_noise=noise)
    assert angular_deg(plain["pred_normals"], out["pred_normals"]).mean().item() > 1.0     # the start latent matters
    root = str(tmp_path)
    for net, sd in sds.items():
        os.makedirs(os.path.join(root, net))
        save_file({k: v.contiguous() for k, v in sd.items()}, os.path.join(root, net, "diffusion_pytorch_model.safetensors"))
    os.makedirs(os.path.join(root, "vae"))
    save_file({k: v.contiguous() for k, v in vsd.items()}, os.path.join(root, "vae", "diffusion_pytorch_model.safetensors"))
    torch.save(prompt, os.path.join(root, "prompt_embeds.pt"))
    disk = StableNormal(config="tiny", num_inference_steps=2, weights=root, yoso=True).forward(data, init_noise=noise)
    assert torch.equal(disk["pred_normals"], out["pred_normals"])
