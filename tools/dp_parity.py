"""N-rank check of the in-switch reduce + Adam kernel against NCCL all_reduce + the flat Adam kernel
(run under torchrun on >= 2 B200s with NVSwitch):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_parity.py
Both replicas start from the same weights and take 3 steps on per-rank batches; parameters, Adam state
and the bf16 shadow must agree to fp32 rounding of the sum order, and all ranks must hold identical
parameters afterwards."""
import contextlib
import io
import os
import sys

sys.argv = sys.argv[:1]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lirec_b200 import dp  # noqa: E402
from lirec_b200.mixed_utils import synthetic  # noqa: E402
from lirec_b200.utils.arg_pars import opt  # noqa: E402


def relationship_term_parity(rank, world):
    """int_rels (MultiTaskMaxMargin): the clip-weighted sum of the per-rank gradients equals the single-process
    gradient of the global batch — the relationship term is normalised by the GLOBAL number of labelled rows
    (ADVICE r1).  Eval mode (the dropout masks are keyed by local row position)."""
    for k, v in dict(tr_maximize=False, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                     rels_n_clips=18, mod_check=False, device="cuda", fused_adam=1, lr=1e-3).items():
        setattr(opt, k, v)
    import lirec_b200.mlp.model as M
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model, loss_fn, optimizer = M.create_model(101, n_rels=15)
    model.eval()
    dp.broadcast_params(model._flat)
    per = 5 + 0                                                  # clips per rank; rank 0 gets one more (unequal shards)
    counts = [per + (1 if r == 0 else 0) for r in range(world)]
    clips = [synthetic.make_clip(900 * 1000003 + i, preset="int_rels") for i in range(sum(counts))]
    a = sum(counts[:rank])
    mine = synthetic.pack_clips(clips[a:a + counts[rank]])
    mine.global_clips = sum(counts)
    loss_fn._dp_world = world
    M.train_step(model, loss_fn, mine.to_device("cuda"))
    loss_fn._dp_world = 1
    g = model._flat_grad.clone()
    scale = dp.allreduce_flat_grad(g, local_clips=counts[rank], global_clips=sum(counts))
    g.mul_(scale)
    M.train_step(model, loss_fn, synthetic.pack_clips(clips).to_device("cuda"))
    ref = model._flat_grad
    err = float((g - ref).abs().max() / (ref.abs().max() + 1e-30))
    if rank == 0:
        print("int_rels: DP gradient vs single-process gradient of the global batch, max-norm relative: %.2e" % err)
    return err < 2e-5


def main():
    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                     rels_n_clips=18, mod_check=False, device="cuda", fused_adam=1, lr=1e-3).items():
        setattr(opt, k, v)
    import lirec_b200.mlp.model as M
    import lirec_b200.mlp.train as TR
    pbs = [synthetic.make_batch(48, seed=100 * rank + i).to_device("cuda") for i in range(3)]
    finals = []
    # the same three steps through lirec_b200.mlp.train.train_step with (a) ncclAllReduce + Adam, (b) the
    # in-switch exchange + Adam after backward, (c) the same with the gate + head bucket overlapped with backward
    for mode in ("nccl", "shard", "shard_overlap", "bucket", "bucket_overlap"):
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model, loss_fn, optimizer = M.create_model(101, n_rels=15)
        model.train()
        model.set_rank(rank)
        dp.broadcast_params(model._flat)
        fused = dp.SwitchReduceAdam.attach(model, optimizer, mode=mode.split("_")[0]) if mode != "nccl" else None
        if mode != "nccl" and fused is None:
            if rank == 0:
                print("SKIP: no NVSwitch multicast support on this box")
            return 0
        if fused is not None:
            fused.overlap = mode.endswith("_overlap")
        for i, pb in enumerate(pbs):
            TR.train_step(model, loss_fn, optimizer, pb, world, fused)
        torch.cuda.synchronize()
        if fused is not None:
            fused.gather_moments()                     # 'shard' mode keeps 1/world of the moments per rank
        finals.append(dict(p=model._flat.clone(), m=optimizer._m.clone(), v=optimizer._v.clone(),
                           pb=model._flat_bf16.float().clone(), g=model._flat_grad.clone()))
        if fused is not None:
            fused.detach()
    ok = True
    for j, mode in ((1, "shard"), (2, "shard_overlap"), (3, "bucket"), (4, "bucket_overlap")):
        for k in (("p", "m", "v", "pb") if mode.startswith("shard") else ("g", "p", "m", "v", "pb")):   # 'shard' never sums g in place
            a, b = finals[0][k], finals[j][k]
            err = float((a - b).abs().max() / (a.abs().max() + 1e-30))
            tol = 4e-3 if k == "pb" else 2e-6
            if rank == 0:
                print("%-3s max-norm relative difference %s vs nccl: %.2e" % (k, mode, err))
            ok = ok and err < tol
    if not torch.equal(finals[3]["p"], finals[4]["p"]) and rank == 0:
        print("note: overlapped and non-overlapped in-switch steps differ in the last bits (shard boundaries move "
              "with the bucket cut, so the in-switch summation order does)")
    ok = relationship_term_parity(rank, world) and ok
    # every rank holds the same replica
    ref = finals[1]["p"].clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, finals[1]["p"]))
    flag = torch.tensor([1 if (ok and same) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("replicas identical across ranks:", same)
        print("DP PARITY", "OK" if int(flag.item()) else "FAILED")
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
