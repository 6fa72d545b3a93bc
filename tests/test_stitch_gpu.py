"""Scene-stitch kernels (ug_stitch_fit / ug_stitch_apply, csrc/stitch.cu) against the torch float64 statement of the
same arithmetic (unigeo_b200/sharding.py, the path the gloo tests run), at a small size and at BASELINE cfg4's
(8 clips x 25 frames x 384 x 512, overlap 5) through a size-independent property: clips cut from ONE disparity video
with per-clip min-max windows come back as clip 0's normalisation of the whole video."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(n_clips, T, ov, H, W, seed, dev):
    starts = [k * (T - ov) for k in range(n_clips)]
    N = starts[-1] + T
    g = torch.Generator(device=dev).manual_seed(seed)
    disp = torch.linspace(0.2, 3.0, N, device=dev).view(N, 1, 1) * (1.0 + 0.5 * torch.rand(N, H, W, generator=g, device=dev))
    clips = []
    for s0 in starts:
        wdw = disp[s0:s0 + T]
        clips.append((1.0 / ((wdw - wdw.min()) / (wdw.max() - wdw.min()) + 0.1)).float())
    x0 = (disp - disp[:T].min()) / (disp[:T].max() - disp[:T].min())
    return starts, N, clips, x0


@pytest.fixture(scope="module")
def eng(cuda):
    from unigeo_b200.metrics import _default_engine
    return _default_engine()


@pytest.mark.parametrize("space", ["disparity", "depth"])
def test_kernels_match_the_host_statement(eng, cuda, space):
    from unigeo_b200 import sharding as sh
    starts, N, clips, _ = _scene(5, 8, 3, 12, 20, 4, cuda)
    ids = list(range(5))
    got = sh.stitch_scene(clips, ids, 5, 3, engine=eng, space=space)
    ref = sh.stitch_scene([c.cpu() for c in clips], ids, 5, 3, space=space)
    for a, b in zip(got, ref):
        assert a.is_cuda and torch.allclose(a.cpu(), b, rtol=2e-6, atol=1e-6)
    again = sh.stitch_scene(clips, ids, 5, 3, engine=eng, space=space)
    assert all(torch.equal(a, b) for a, b in zip(got, again))            # fixed-order sums: reruns identical
    with pytest.raises(ValueError):
        sh.stitch_scene(clips, ids, 5, 3)                                 # CUDA tensors without the engine: no fallback


def test_cfg4_size_recovers_clip0_frame(eng, cuda):
    from unigeo_b200 import sharding as sh
    starts, N, clips, x0 = _scene(8, 25, 5, 384, 512, 7, cuda)
    st = sh.stitch_scene(clips, list(range(8)), 8, 5, engine=eng)
    video = sh.assemble_scene(st, starts, N, 5)
    assert video.shape == (165, 384, 512)
    want = 1.0 / (x0 + 0.1).clamp(min=1e-3)
    ok = (x0 + 0.1) > 2e-3
    rel = ((video.double() - want)[ok].abs() / want[ok]).max().item()
    assert rel <= 5e-4, rel
    # seam: the ramped head of clip k starts exactly at clip k-1's mapped tail
    for k in range(1, 8):
        assert torch.allclose(st[k][0], st[k - 1][-5], rtol=1e-5, atol=1e-6)


def test_constant_overlap_and_single_clip(eng, cuda):
    from unigeo_b200 import sharding as sh
    a = torch.full((6, 4, 4), 2.0, device=cuda)
    b = torch.full((6, 4, 4), 5.0, device=cuda)
    out = sh.stitch_scene([a, b], [0, 1], 2, 2, engine=eng, space="depth")
    assert torch.isfinite(out[1]).all() and torch.allclose(out[1], torch.full_like(b, 2.0))
    one = sh.stitch_scene([a], [0], 1, 2, engine=eng)
    assert torch.equal(one[0], a)
