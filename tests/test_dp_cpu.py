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


def test_every_rank_takes_a_step_for_every_global_batch():
    """a last global batch with fewer clips than ranks (37 clips, batch 8, 8 ranks: 5 left)
    used to be dropped by the ranks whose share was empty, un-pairing the collectives of the step.  Every
    rank now gets an entry per global batch (an empty one -> EmptyShard), and the shares partition the batch."""
    import sys
    sys.argv = sys.argv[:1]
    from lirec_b200.mixed_utils.classification_dataloader import EmptyShard, _IndexView, plan_batches
    n, bs, world = 37, 8, 8
    order = list(range(n))
    plans = [plan_batches(n, bs, order, r, world) for r in range(world)]
    assert len({len(p) for p in plans}) == 1 and len(plans[0]) == 5
    for step in range(len(plans[0])):
        shares = [plans[r][step][0] for r in range(world)]
        assert sum(shares, []) == order[step * bs:(step + 1) * bs]
        assert {plans[r][step][1] for r in range(world)} == {len(sum(shares, []))}
    assert [len(plans[r][-1][0]) for r in range(world)] == [1, 1, 1, 1, 1, 0, 0, 0]

    class _DS:
        def __getitem__(self, j):
            raise AssertionError("an empty share must not touch the dataset")
    item = _IndexView(_DS(), plans[7])[4]
    assert isinstance(item, EmptyShard) and item.B == 0 and item.global_clips == 5 and item.host is item


def _rel_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.argv = sys.argv[:1]
    from types import SimpleNamespace
    from lirec_b200 import dp
    from lirec_b200.mlp import model as M
    dp.init_from_env(backend="gloo")
    # global batch: 7 rows; row terms g_i, rows 1 and 5 are labelled None; shards 4 / 3
    g = torch.tensor([1.0, 9.0, 2.0, 4.0, 8.0, 9.0, 16.0])
    labelled = torch.tensor([1, 0, 1, 1, 1, 0, 1], dtype=torch.bool)
    a, b = dp.shard_range(7, rank, world)
    loc_g, loc_l = g[a:b], labelled[a:b]
    pb = SimpleNamespace(device="cpu", B=b - a, host=SimpleNamespace(global_clips=7))
    mod = SimpleNamespace(_dp_world=world)

    def run(scale):          # what the fused kernel returns: per-row terms and gradients times `scale`
        t = torch.where(loc_l, loc_g, torch.zeros_like(loc_g)) * scale
        return t, t.clone()
    t, d = M._rel_term(mod, pb, int(loc_l.sum()), run)
    # the exchange weights rank r by B_r / B and sums
    contrib = d.sum() * (b - a) / 7.0
    dist.all_reduce(contrib)
    ok = abs(float(contrib) - float(g[labelled].sum() / labelled.sum())) < 1e-6
    # a rank without any labelled row still joins the collective and contributes nothing
    mod2 = SimpleNamespace(_dp_world=world)
    n_sel = 0 if rank == 1 else int(loc_l.sum())
    t2, d2 = M._rel_term(mod2, pb, n_sel, run)
    ok2 = (t2 is None) if rank == 1 else (t2 is not None)
    q.put((rank, ok, ok2))
    dist.destroy_process_group()


def test_relationship_term_is_normalised_over_the_global_batch():
    """MultiTaskMaxMargin / MultiTaskCrossEntropyLoss average their relationship term over
    the non-None rows; under data parallelism that count is global (one tiny all_reduce), so that the
    clip-weighted gradient sum equals the single-process gradient."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rel_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, *oks in res:
        assert all(oks), (rank, oks)
