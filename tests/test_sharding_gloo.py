"""N>1 host logic on CPU: two gloo ranks shard the clips of a scene and run the overlap
all-gather + stitch; the assembled video must equal the single-process result."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unigeo_b200 import sharding as sh
    T, ov, n = 8, 2, 5
    starts = [k * (T - ov) for k in range(n)]
    N = starts[-1] + T
    g = torch.Generator().manual_seed(0)
    scene = torch.linspace(1, 4, N).view(N, 1, 1) + torch.rand(N, 4, 6, generator=g) * 0.3
    clips = [scene[s:s + T] * (1.0 - 0.1 * k) + 0.05 * k for k, s in enumerate(starts)]
    mine = sh.clips_of_rank(n, rank, world)
    st = sh.stitch_scene([clips[k] for k in mine], mine, n, ov, rank, world, space="depth")
    # collect everything on rank 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, [t.numpy() for t in st]))
    if rank == 0:
        full = {}
        for ids, ts in gathered:
            for k, t in zip(ids, ts):
                full[k] = torch.from_numpy(t)
        video = sh.assemble_scene([full[k] for k in range(n)], starts, N, ov)
        single = sh.assemble_scene(sh.stitch_scene(clips, list(range(n)), n, ov, space="depth"), starts, N, ov)
        q.put((torch.allclose(video, single, atol=1e-6), torch.allclose(video, scene, atol=1e-4)))
    # per-clip metric rows -> every rank, in clip order (5 clips on 2 ranks: rank 1 pads one row)
    rows = torch.tensor([[float(k), 10.0 * k + 0.5] for k in mine], dtype=torch.float64)
    table = sh.gather_metric_rows(mine, rows, n, rank, world)
    want = torch.tensor([[float(k), 10.0 * k + 0.5] for k in range(n)], dtype=torch.float64)
    assert torch.equal(table, want), table
    assert torch.equal(sh.average_row(table), want.mean(0))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stitch_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, exact = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and exact


def _disparity_scene(n_clips, T, ov, seed=1):
    """One true disparity video; every clip is min-max normalised over ITS OWN window and turned into depth the way the
    adapter does (model/depthcrafter.py:95-96): x = (disp - min) / (max - min), depth = 1 / (x + 0.1)."""
    starts = [k * (T - ov) for k in range(n_clips)]
    N = starts[-1] + T
    g = torch.Generator().manual_seed(seed)
    disp = torch.linspace(0.2, 3.0, N).view(N, 1, 1) * (1.0 + 0.5 * torch.rand(N, 6, 8, generator=g))
    clips = []
    for s0 in starts:
        w = disp[s0:s0 + T]
        clips.append(1.0 / ((w - w.min()) / (w.max() - w.min()) + 0.1))
    x0 = (disp - disp[:T].min()) / (disp[:T].max() - disp[:T].min())           # the whole video in clip 0's window
    return starts, N, clips, x0


def test_disparity_space_fit_recovers_clip0_frame():
    """per-clip min-max windows of the SAME disparity are related by an affine map in normalised disparity,
    a projective one in depth.  Fitting in disparity space reproduces clip 0's normalisation of the whole video exactly;
    fitting the depths leaves a systematic residual."""
    from unigeo_b200 import sharding as sh
    starts, N, clips, x0 = _disparity_scene(4, 8, 3)
    ids = list(range(4))
    st = sh.stitch_scene(clips, ids, 4, 3)                                     # default: space="disparity"
    video = sh.assemble_scene(st, starts, N, 3)
    want = 1.0 / (x0 + 0.1).clamp(min=1e-3)
    ok = (x0 + 0.1) > 1e-3
    assert torch.allclose(video[ok], want[ok], rtol=1e-4, atol=1e-5)
    dep = sh.assemble_scene(sh.stitch_scene(clips, ids, 4, 3, space="depth"), starts, N, 3)
    assert (dep[ok] - want[ok]).abs().max() > 100 * (video[ok] - want[ok]).abs().max()


def test_constant_overlap_keeps_the_scale():
    """det = 0 (a constant overlap region): no NaN / inf scale down the chain -- s = 1, t = mean difference."""
    from unigeo_b200 import sharding as sh
    s, t = sh.fit_scale_shift(torch.full((50,), 2.0), torch.full((50,), 5.0))
    assert float(s) == 1.0 and float(t) == 3.0
    a = torch.full((6, 4, 4), 2.0)
    b = torch.full((6, 4, 4), 5.0)
    out = sh.stitch_scene([a, b], [0, 1], 2, 2, space="depth")
    assert all(torch.isfinite(o).all() for o in out) and torch.allclose(out[1], torch.full_like(b, 2.0))


def _worker_sparse(rank, world, port, q):
    """More ranks than clips: the ranks without a clip still enter every collective (no deadlock, no IndexError)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unigeo_b200 import sharding as sh
    n, T, ov = 2, 8, 3
    starts, N, clips, x0 = _disparity_scene(n, T, ov, seed=3)
    mine = sh.clips_of_rank(n, rank, world)
    st = sh.stitch_scene([clips[k] for k in mine], mine, n, ov, rank, world)
    rows = torch.tensor([[float(k), 2.0 * k] for k in mine], dtype=torch.float64)
    table = sh.gather_metric_rows(mine, rows, n, rank, world)
    assert torch.equal(table, torch.tensor([[0.0, 0.0], [1.0, 2.0]], dtype=torch.float64)), table
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, [t.numpy() for t in st]))
    if rank == 0:
        full = {k: torch.from_numpy(t) for ids, ts in gathered for k, t in zip(ids, ts)}
        video = sh.assemble_scene([full[k] for k in range(n)], starts, N, ov)
        single = sh.assemble_scene(sh.stitch_scene(clips, list(range(n)), n, ov), starts, N, ov)
        q.put((len(mine), torch.allclose(video, single, atol=1e-6)))
    dist.barrier()
    dist.destroy_process_group()


def test_more_ranks_than_clips():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sparse, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    n_mine, same = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_mine == 1 and same
