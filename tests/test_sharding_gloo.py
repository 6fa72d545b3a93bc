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
    st = sh.stitch_scene([clips[k] for k in mine], mine, n, ov, rank, world)
    # collect everything on rank 0
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, [t.numpy() for t in st]))
    if rank == 0:
        full = {}
        for ids, ts in gathered:
            for k, t in zip(ids, ts):
                full[k] = torch.from_numpy(t)
        video = sh.assemble_scene([full[k] for k in range(n)], starts, N, ov)
        single = sh.assemble_scene(sh.stitch_scene(clips, list(range(n)), n, ov), starts, N, ov)
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
