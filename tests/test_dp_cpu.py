"""Data-parallel host logic on CPU with gloo (world_size 2): gradient averaging over equal shards,
clip-count weighting over unequal shards, and that sharded batches partition the global batch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from lirec_b200 import dp
    r, w, _ = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # equal shards: per-rank mean gradients average to the global-batch mean gradient
    g_all = torch.arange(8, dtype=torch.float32).view(8, 1) * torch.ones(8, 5)      # per-clip gradients
    a, b = dp.shard_range(8, rank, world)
    flat = g_all[a:b].mean(0).clone()
    scale = dp.allreduce_flat_grad(flat)
    ok1 = torch.allclose(flat * scale, g_all.mean(0))
    # unequal shards (last short batch of 7 clips): weight by local / global clip counts
    a, b = dp.shard_range(7, rank, world)
    flat = g_all[a:b].mean(0).clone()
    scale = dp.allreduce_flat_grad(flat, local_clips=b - a, global_clips=7)
    ok2 = torch.allclose(flat * scale, g_all[:7].mean(0))
    # in-place averaging variant used with torch.optim.Adam
    flat = torch.full((5,), float(rank + 1))
    s = dp.allreduce_flat_grad(flat, average_in_place=True)
    ok3 = s == 1.0 and torch.allclose(flat, torch.full((5,), 1.5))
    p = torch.full((3,), float(rank))
    dp.broadcast_params(p, src=0)
    ok4 = bool((p == 0).all())
    q.put((rank, ok1, ok2, ok3, ok4))
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, *oks in res:
        assert all(oks), (rank, oks)


def test_sharded_loader_partitions_global_batches():
    import sys
    sys.argv = sys.argv[:1]
    from lirec_b200 import dp
    n, bs, world = 37, 8, 2
    order = list(range(n))
    seen = []
    for s in range(0, n, bs):
        idx = order[s:s + bs]
        parts = [idx[slice(*dp.shard_range(len(idx), r, world))] for r in range(world)]
        assert sum(parts, []) == idx
        seen += idx
    assert seen == order
