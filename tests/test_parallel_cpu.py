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
    # bf16 wire format: every rank ends with the SAME values (the average of the bf16-rounded contributions, re-rounded)
    flat2 = mine.clone()
    red2 = GradReducer(flat2, bucket_numel=2_500, grad_dtype="bf16")
    for hi, lo in zip(edges[:-1], edges[1:]):
        red2.ready(lo, hi)
    red2.finish()
    want2 = (sum(g.to(torch.bfloat16).float() for g in gathered) / world).to(torch.bfloat16).float()
    g2 = [torch.zeros(10_000) for _ in range(world)]
    dist.all_gather(g2, flat2)
    ok = ok and torch.equal(flat2, want2) and all(torch.equal(g2[0], t) for t in g2) and torch.allclose(flat2, want, rtol=1e-2, atol=1e-2)
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


def _dp_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mau_b200
    from mau_b200.parallel import DataParallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(1000 + rank)                      # different initial weights per rank
    m = mau_b200.UrbanPredictor("unet", 23, 828, 16, 8, 8, 32, 2, base_filters=8,
                                temporal_embeddings=False, metadata_embeddings=True)
    dp = DataParallel(m, sync_bn=True)                  # broadcasts rank 0's parameters and buffers
    w = m.model.conv2_0.conv1.weight.detach().clone()
    gathered = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    rm = m.state_dict()["model.conv0_0.bn1.running_mean"]
    rm.fill_(float(rank))                               # ranks drift apart in local-BN mode ...
    dp.sync_buffers()                                   # ... and are averaged before a checkpoint
    ok = same and torch.allclose(rm, torch.full_like(rm, (world - 1) / 2)) and m.model._dp is dp and dp.sync_bn
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_data_parallel_broadcast_and_buffer_sync_gloo_world2():
    """Host side of the data-parallel wrapper on CPU tensors: rank-0 broadcast at construction, running-stat
    averaging, registration on the module (the kernels themselves need a GPU: tools/dp_parity.py)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_flat_gradient_views_are_detached_before_reuse():
    """DataParallel hands autograd views of one flat buffer; a gradient kept across backward passes must not be
    overwritten by the next pass (host logic of make_grads, exercised with a stand-in plan on CPU)."""
    import mau_b200  # noqa: F401
    from mau_b200.parallel import DataParallel

    class _Plan:
        device = torch.device("cpu")
        num_state = 3
        def set_grad_hook(self, fn):
            self.hook = fn

    dp = DataParallel.__new__(DataParallel)
    dp.group, dp.bucket_numel, dp.sync_bn = None, 1 << 20, False
    plan = _Plan()
    params = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 3))]
    full, outs = dp.make_grads(plan, [0, 2], [p.shape for p in params], params)
    assert full[1] is None and outs[0].untyped_storage().data_ptr() == outs[1].untyped_storage().data_ptr()
    outs[0].fill_(1.0); outs[1].fill_(2.0)
    for p, g in zip(params, outs):
        p.grad = g                                   # what autograd does when it steals the incoming gradient
    full2, outs2 = dp.make_grads(plan, [0, 2], [p.shape for p in params], params)
    outs2[0].fill_(5.0); outs2[1].fill_(7.0)         # the next backward writes into the flat buffer ...
    assert torch.all(params[0].grad == 1.0) and torch.all(params[1].grad == 2.0)   # ... kept gradients survive
