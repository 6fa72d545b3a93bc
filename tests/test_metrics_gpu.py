"""Metric kernels (ug_depth_metrics / ug_normal_metrics, csrc/metrics.cu) against
  * the golden values recorded from the UNMODIFIED reference functions (tests/golden/metrics_kat.npz),
  * the oracle restatement (bit-exact against that fixture) on seeded inputs, ragged sizes, with / without mask,
  * size-independent properties at the full 25x384x512 clip size.
Tolerances (fp32 per-pixel arithmetic in the reference's order; fp64 fixed-order sums here against the
reference's fp32 means; closed-form fp64 scale/shift against its fp32 SVD lstsq): depth metrics |d| <= 2e-5
(absolute, the values are O(0.1..1)), delta fractions <= 2 pixels' worth, normal mean / rmse <= 2e-4 degrees,
percentages <= 2 pixels' worth; counts, the empty-mask conventions and the MEDIAN are exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _clip(n_f, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(n_f, h, w, generator=g) * 12.0 - 1.0           # some <= 0 (invalid), none above max_depth...
    gt[0, :2] = 95.0                                                  # ...except these rows (> 80)
    pred = 0.37 * gt.clamp(min=0.05) + 0.8 + 0.1 * torch.randn(n_f, h, w, generator=g)
    mask = torch.rand(n_f, h, w, generator=g) > 0.3
    pn = torch.nn.functional.normalize(torch.randn(n_f, h, w, 3, generator=g), dim=-1)
    gn = torch.nn.functional.normalize(pn + 0.4 * torch.randn(n_f, h, w, 3, generator=g), dim=-1)
    return pred, gt, mask, pn, gn


def _check_depth(res, ref, n_valid):
    assert res["valid_pixels"] == ref["valid_pixels"]
    for k in ("Abs Rel", "Sq Rel", "RMSE", "Log RMSE"):
        assert abs(res[k] - ref[k]) <= 2e-5 * max(1.0, abs(ref[k])), (k, res[k], ref[k])
    for k in ("delta < 1.", "delta < 1.25", "delta < 1.25^2", "delta < 1.25^3"):
        assert abs(res[k] - ref[k]) <= 2.0 / max(n_valid, 1) + 1e-7, (k, res[k], ref[k])


def _check_normal(res, ref, n_valid):
    for k in ("normal mean", "normal rmse"):
        assert abs(res[k] - ref[k]) <= 2e-4, (k, res[k], ref[k])
    assert abs(res["normal median"] - ref["normal median"]) <= 1e-3       # acos ulp; exactness tested separately
    for k in ("angle < 5", "angle < 7.5", "angle < 11.25", "angle < 22.5", "angle < 30"):
        assert abs(res[k] - ref[k]) <= 200.0 / max(n_valid, 1) + 1e-5, (k, res[k], ref[k])


def test_golden_reference_values(cuda):
    from unigeo_b200.metrics import depth_evaluation, normal_evaluation
    k = np.load(os.path.join(G, "metrics_kat.npz"))
    res = depth_evaluation(k["pred"], k["gt"], custom_mask=k["mask"], align_with_lstsq=True)[0]
    ref = {str(a): float(b) for a, b in zip(k["depth_keys"], k["depth_vals"])}
    ref["valid_pixels"] = int(ref["valid_pixels"])
    _check_depth(res, ref, ref["valid_pixels"])
    nres = normal_evaluation(k["pn"], k["gn"], custom_mask=k["mask"])
    nref = {str(a): float(b) for a, b in zip(k["normal_keys"], k["normal_vals"])}
    _check_normal(nres, nref, int(k["mask"].sum()))


@pytest.mark.parametrize("shape,use_mask", [((3, 37, 53), True), ((1, 16, 16), True), ((4, 64, 96), False),
                                            ((2, 5, 7), True)])
def test_matches_oracle(cuda, shape, use_mask):
    from oracle import metrics as OM
    from unigeo_b200.metrics import depth_evaluation, normal_evaluation
    pred, gt, mask, pn, gn = _clip(*shape, seed=sum(shape))
    m = mask if use_mask else None
    res, err, aligned, gtv = depth_evaluation(pred, gt, custom_mask=m, align_with_lstsq=True)
    ref = OM.depth_evaluation(pred, gt, m)
    _check_depth(res, ref, ref["valid_pixels"])
    # the three maps of eval_depth.py:166-213
    valid = (gt > 0) & (gt < 80)
    assert err.shape == (shape[0] * shape[1], shape[2]) and err.is_cuda
    assert torch.equal(gtv.cpu().reshape(gt.shape), torch.where(valid, gt, torch.zeros_like(gt)))
    assert torch.all(err.cpu().reshape(gt.shape)[~valid] == 0)
    a = aligned.cpu().reshape(gt.shape)
    e_ref = torch.where(valid, (a - gt).abs() / gt, torch.zeros_like(gt))
    assert torch.equal(err.cpu().reshape(gt.shape), e_ref)
    # normals need a mask in the reference (err[None] would add an axis); full mask when unused
    nm = mask if use_mask else torch.ones(shape, dtype=torch.bool)
    nres = normal_evaluation(pn, gn, custom_mask=nm)
    _check_normal(nres, OM.normal_evaluation(pn, gn, nm), int(nm.sum()))
    if not use_mask:
        assert normal_evaluation(pn, gn, custom_mask=None) == nres


def test_median_is_exact_and_deterministic(cuda):
    """torch.median semantics (lower middle value), bit-exact on the kernel's own per-pixel errors, even / odd counts."""
    from unigeo_b200.metrics import _default_engine
    eng = _default_engine()
    for n_f, h, w, seed in [(3, 37, 53, 1), (2, 64, 64, 2), (1, 1, 2, 3), (1, 1, 1, 4)]:
        _, _, mask, pn, gn = _clip(n_f, h, w, seed)
        mask.view(-1)[0] = True
        vals, err = eng.normal_metrics(pn, gn, mask, with_map=True)
        e = err.cpu()[mask]
        assert np.float32(vals[1]) == torch.median(e).item()
        assert abs(vals[0] - e.double().mean().item()) < 1e-9 and int(round(vals[3] * e.numel() / 100)) == int((e < 5).sum())
        assert eng.normal_metrics(pn, gn, mask) == vals                 # fixed-order reductions: reruns identical


def test_empty_and_degenerate(cuda):
    from unigeo_b200.metrics import depth_evaluation, normal_evaluation
    pred, gt, mask, pn, gn = _clip(2, 16, 24, 9)
    res = depth_evaluation(pred, -gt.abs(), custom_mask=mask, align_with_lstsq=True, with_maps=False)[0]
    assert res["valid_pixels"] == 0 and all(v == 0 for v in res.values())      # eval_depth.py:217-227
    res = depth_evaluation(pred, gt, custom_mask=torch.zeros_like(mask), align_with_lstsq=True, with_maps=False)[0]
    assert res["valid_pixels"] == 0 and res["Abs Rel"] == 0
    nres = normal_evaluation(pn, gn, custom_mask=torch.zeros_like(mask))
    assert all(np.isnan(v) for v in nres.values())                             # torch: mean of an empty tensor
    with pytest.raises(NotImplementedError):
        depth_evaluation(pred, gt)                                            # median scaling: not on the device


def test_full_size_properties(cuda):
    """25x384x512 (BASELINE cfg2), tensors already on the device: affine invariance of the aligned metrics,
    identity / antipodal normals, exact counts."""
    from unigeo_b200.metrics import depth_evaluation, normal_evaluation
    g = torch.Generator(device=cuda).manual_seed(11)
    shape = (25, 384, 512)
    gt = torch.rand(shape, generator=g, device=cuda) * 10 - 0.5
    pred = 1.0 / (0.2 * gt.clamp(min=0.1) + 0.05 * torch.rand(shape, generator=g, device=cuda) + 0.1)
    mask = torch.rand(shape, generator=g, device=cuda) > 0.25
    r1 = depth_evaluation(pred, gt, custom_mask=mask, align_with_lstsq=True, with_maps=False)[0]
    r2 = depth_evaluation(3.5 * pred - 2.0, gt, custom_mask=mask, align_with_lstsq=True, with_maps=False)[0]
    assert r1["valid_pixels"] == int(((gt > 0) & (gt < 80) & mask).sum())
    for k in ("Abs Rel", "Sq Rel", "RMSE", "Log RMSE"):
        assert abs(r1[k] - r2[k]) <= 1e-4 * max(1.0, r1[k]), (k, r1[k], r2[k])
    exact = depth_evaluation(0.25 * gt - 1.0, gt, custom_mask=mask, align_with_lstsq=True, with_maps=False)[0]
    assert exact["Abs Rel"] < 1e-5 and exact["delta < 1.25"] > 0.9999     # gt arbitrarily close to 0 in this draw
    n = torch.nn.functional.normalize(torch.randn(shape + (3,), generator=g, device=cuda), dim=-1)
    same = normal_evaluation(n, n, custom_mask=mask)
    # cos = 1 / (1 + 1e-6) in the reference's formula (eval_normal.py:15): acos gives 0.081 degrees, not 0
    assert same["normal mean"] < 0.15 and same["angle < 5"] == 100.0
    anti = normal_evaluation(n, -n, custom_mask=mask)
    assert anti["normal mean"] > 179.85 and anti["angle < 30"] == 0.0 and abs(anti["normal median"] - 180.0) < 0.15
    # median of a known distribution: pred tilted from gt = +z by theta in [0, 60) degrees, uniform
    theta = torch.rand(shape, generator=g, device=cuda) * (np.pi / 3)
    z = torch.zeros(shape + (3,), device=cuda)
    z[..., 2] = 1
    p = torch.stack([theta.sin(), torch.zeros_like(theta), theta.cos()], -1)
    r = normal_evaluation(p, z, custom_mask=torch.ones(shape, dtype=torch.bool, device=cuda))
    assert abs(r["normal median"] - torch.median(theta).item() * 180 / np.pi) < 2e-2
    assert abs(r["angle < 30"] - 50.0) < 0.2 and abs(r["normal mean"] - 30.0) < 0.05


def test_plugin_scored_on_device(cuda):
    """forward_device keeps the outputs on the GPU; scoring them there equals scoring the CPU tensors of forward."""
    from harness.synthetic import gt_label, make_clip
    from oracle import metrics as OM
    from unigeo_b200.metrics import depth_evaluation, normal_evaluation
    from unigeo_b200.model import DepthCrafter
    data = make_clip(5, 128, 256, seed=3)
    plug = DepthCrafter(config="tiny", dtype="fp16", weights="synthetic", num_inference_steps=2, clip="none", seed=1)
    enc = torch.randn(5, plug.cfg.clip_embed_dim, generator=torch.Generator().manual_seed(2))
    dev = plug.forward_device(data, enc=enc)          # seed=1: both calls draw the same noise
    cpu = plug.forward(data, enc=enc)
    assert dev["pred_depths"].is_cuda and torch.equal(dev["pred_depths"].cpu(), cpu["pred_depths"])
    assert torch.equal(dev["pred_normals"].cpu(), cpu["pred_normals"])
    gt = gt_label(data)
    res = depth_evaluation(dev["pred_depths"], gt["gt_depths"], custom_mask=gt["gt_masks"], align_with_lstsq=True,
                           engine=plug.engine, with_maps=False)[0]
    ref = OM.depth_evaluation(cpu["pred_depths"], gt["gt_depths"], gt["gt_masks"])
    _check_depth(res, ref, ref["valid_pixels"])
    nres = normal_evaluation(dev["pred_normals"], gt["gt_normals"], custom_mask=gt["gt_masks"], engine=plug.engine)
    _check_normal(nres, OM.normal_evaluation(cpu["pred_normals"], gt["gt_normals"], gt["gt_masks"]),
                  int(torch.as_tensor(gt["gt_masks"]).sum()))


def test_edge_fixture_recorded_from_the_reference(cuda):
    """metrics_kat_edge.npz (recorded from the UNMODIFIED reference functions, tests/golden/make_golden.py) against the
    device kernels: aligned predictions below zero (the clamp(1e-5) / log branch of eval_depth.py:141-164), sparse masks
    with even and odd survivor counts (torch.median's lower-middle rule), custom_mask=None, and the three full-size maps.
    Same tolerances as above; valid_pixels and the gt_valid map exact, the error / aligned maps within the fp32 rounding
    of (s, t) (closed-form fp64 solve here, SVD lstsq in the reference)."""
    from unigeo_b200.metrics import depth_evaluation, normal_evaluation
    k = np.load(os.path.join(G, "metrics_kat_edge.npz"))
    for name in [str(c) for c in k["cases"]]:
        use_mask = bool(k[f"{name}_use_mask"])
        mask = k[f"{name}_mask"] if use_mask else None
        res, err, aligned, gtv = depth_evaluation(k[f"{name}_pred"], k[f"{name}_gt"], custom_mask=mask,
                                                  align_with_lstsq=True)
        ref = {str(a): float(b) for a, b in zip(k["depth_keys"], k[f"{name}_depth_vals"])}
        ref["valid_pixels"] = int(ref["valid_pixels"])
        _check_depth(res, ref, ref["valid_pixels"])
        nm = k[f"{name}_mask"] if use_mask else np.ones_like(k[f"{name}_mask"])
        nres = normal_evaluation(k[f"{name}_pn"], k[f"{name}_gn"], custom_mask=nm)
        nref = {str(a): float(b) for a, b in zip(k["normal_keys"], k[f"{name}_normal_vals"])}
        _check_normal(nres, nref, int(nm.sum()))
        if not use_mask:
            assert normal_evaluation(k[f"{name}_pn"], k[f"{name}_gn"], custom_mask=None) == nres
        if name == "neg":
            assert (k["neg_pred_aligned"] < 0).any()              # the case does exercise the clamp branch
            assert np.array_equal(gtv.cpu().numpy(), k["neg_gt_valid"])
            a, e = aligned.cpu().numpy(), err.cpu().numpy()
            assert a.shape == k["neg_pred_aligned"].shape
            assert np.abs(a - k["neg_pred_aligned"]).max() <= 2e-5 * max(1.0, np.abs(k["neg_pred_aligned"]).max())
            assert np.abs(e - k["neg_err_map"]).max() <= 2e-5 * max(1.0, np.abs(k["neg_err_map"]).max())
            assert (e[k["neg_gt_valid"] == 0] == 0).all()          # invalid pixels carry no error
