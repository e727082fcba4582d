"""CPU, world_size 2, gloo: the gradient bucketing / all-reduce logic of the data-parallel path."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mau_b200  # noqa: F401
    from mau_b200.parallel import GradReducer, shard_tiles
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    flat = torch.randn(10_000)
    mine = flat.clone()
    red = GradReducer(flat, bucket_numel=2_500)
    # readiness arrives in backward order: high offsets first, in uneven layer-sized pieces
    edges = [10_000, 9_990, 8_000, 7_999, 5_000, 1_200, 64, 0]
    for hi, lo in zip(edges[:-1], edges[1:]):
        red.ready(lo, hi)
    red.finish()
    gathered = [torch.zeros(10_000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    want = sum(gathered) / world
    ok = torch.allclose(flat, want, atol=1e-6) and red.launched >= 2 and not red.pending
    tiles = [list(shard_tiles(10, r, world)) for r in range(world)]
    ok = ok and sorted(sum(tiles, [])) == list(range(10))
    q.put((rank, bool(ok), red.launched))
    dist.destroy_process_group()


def test_grad_reducer_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_grad_reducer_single_process_is_noop():
    import mau_b200  # noqa: F401
    from mau_b200.parallel import GradReducer
    flat = torch.arange(100, dtype=torch.float32)
    red = GradReducer(flat.clone(), bucket_numel=10)
    red.ready(50, 100); red.ready(0, 50); red.finish()
    assert torch.equal(red.flat, flat) and red.launched == 0
