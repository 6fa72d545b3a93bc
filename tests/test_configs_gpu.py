"""BASELINE.json configs as GPU tests (parity-test cases, not bench lines):
  cfg1  8-frame 256x256 clip, 5 steps, through the plugin call (here on the tiny architecture so the
        CPU oracle finishes in seconds) -- Abs Rel / delta parity against the oracle arm;
  weights through the diffusers-directory loader (safetensors written with synthetic tensors);
  bench.py end to end on the tiny config (guards the contract keys)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg1_plumbing_8x256x256_5steps(cuda):
    from harness.synthetic import gt_label, make_clip
    from oracle import metrics as OM
    from oracle import postprocess as OP
    from oracle.pipeline import depthcrafter_pipeline
    from unigeo_b200.model import DepthCrafter
    from unigeo_b200.weights import synthetic_state_dict, unet_param_shapes, vae_param_shapes
    T, H, W = 8, 256, 256
    data = make_clip(T, H, W, seed=1)
    plug = DepthCrafter(config="tiny", dtype="fp16", weights="synthetic", clip="none")   # 5 steps = reference default
    assert plug.num_inference_steps == 5
    g = torch.Generator().manual_seed(9)
    enc = torch.randn(T, plug.cfg.clip_embed_dim, generator=g)
    aug = torch.randn(T, 3, H, W, generator=g)
    init = torch.randn(T, 4, H // 8, W // 8, generator=g)
    out = plug.forward(data, enc=enc, aug_noise=aug, init_noise=init)
    usd = synthetic_state_dict(unet_param_shapes(plug.cfg.unet), 1000)
    vsd = synthetic_state_dict(vae_param_shapes(plug.cfg.vae), 2000)
    with torch.no_grad():
        ref = depthcrafter_pipeline(usd, vsd, plug.cfg, torch.from_numpy(OP.prepare_input(data["images"])),
                                    enc[None], aug, init[None], 5).numpy()
    ref_depth = torch.from_numpy(OP.disparity_to_depth(ref)).float()
    gt = gt_label(data)
    a = OM.depth_evaluation(ref_depth, gt["gt_depths"], gt["gt_masks"])
    b = OM.depth_evaluation(out["pred_depths"], gt["gt_depths"], gt["gt_masks"])
    assert abs(a["Abs Rel"] - b["Abs Rel"]) <= 1e-3, (a, b)
    assert abs(a["delta < 1.25"] - b["delta < 1.25"]) <= 2e-3


def test_weights_from_diffusers_directory(cuda, tmp_path):
    """The `weights: pretrained` route: *.safetensors with diffusers keys -> ug_ctx_load_weight."""
    from safetensors.torch import save_file
    from unigeo_b200.config import tiny_config
    from unigeo_b200.engine import Engine
    from unigeo_b200.weights import load_diffusers_dir, synthetic_state_dict, vae_param_shapes
    cfg = tiny_config()
    sd = synthetic_state_dict(vae_param_shapes(cfg.vae), 4, torch.float16)
    d = tmp_path / "vae"
    d.mkdir()
    save_file({k: v.contiguous() for k, v in sd.items()}, str(d / "diffusion_pytorch_model.fp16.safetensors"))
    loaded = load_diffusers_dir(str(d))
    assert set(loaded) == set(sd)
    e1, e2 = Engine(cfg, "fp16"), Engine(cfg, "fp16")
    e1.load_state_dict("vae", sd)
    e2.load_state_dict("vae", loaded)
    lat = torch.randn(3, 4, 8, 16, generator=torch.Generator().manual_seed(0))
    assert torch.equal(e1.vae_decode(lat, 8), e2.vae_decode(lat, 8))


def test_missing_weight_is_reported(cuda):
    from unigeo_b200._lib import UgError
    from unigeo_b200.config import tiny_config
    from unigeo_b200.engine import Engine
    from unigeo_b200.weights import synthetic_state_dict, vae_param_shapes
    cfg = tiny_config()
    sd = synthetic_state_dict(vae_param_shapes(cfg.vae), 4)
    sd.pop("decoder.conv_out.bias")
    e = Engine(cfg, "fp16")
    e.load_state_dict("vae", sd)
    with pytest.raises(UgError) as ei:
        e.vae_decode(torch.zeros(2, 4, 8, 16))
    assert "decoder.conv_out.bias" in str(ei.value)


def test_bench_contract_keys_tiny(cuda):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "tiny", "--frames", "5",
                        "--height", "128", "--width", "256", "--steps", "2", "--warmup", "1", "--e2e-steps", "2",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in line, k
    assert line["metric"] == "denoising-steps/sec" and line["value"] > 0 and line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["roofline"]["bound"] == "tensor"
    assert line["vs_baseline"] is None and "workload" in line["config"]
