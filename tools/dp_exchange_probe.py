"""The exchange step of the data-parallel path, timed alone (report the gradient sum against
NCCL on the same box).  Under torchrun on N >= 2 GPUs of one NVSwitch box:

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/dp_exchange_probe.py

times, with CUDA events and the max over ranks, per step:
  (a) dp.SwitchReduceAdam.step(): mode 'shard' = lirec_dp_reduce_adam_bcast (in-switch sum of the rank's shard + Adam on
      it + multicast of the new parameters); mode 'bucket' (LIREC_DP_MODE=bucket) = lirec_dp_exchange + lirec_adam_flat;
  (b) ncclAllReduce(sum, fp32) of the same flat gradient buffer, alone;
  (c) (b) + lirec_adam_flat, the two-launch form the fused kernel replaces;
  (d) lirec_adam_flat alone.
Rank 0 prints one JSON line with the times and the all-reduce bus bandwidth 2 (N-1)/N * bytes / time
(NCCL's busbw convention) of (b) and of (a) - (d)."""
import contextlib
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = sys.argv[:1]
import torch
import torch.distributed as dist

from lirec_b200 import dp
from lirec_b200.utils.arg_pars import opt

rank, world, local = dp.init_from_env()
if world < 2:
    raise SystemExit("run under torchrun with at least 2 ranks")
torch.cuda.set_device(local)
for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                 rels_n_clips=18, mod_check=False, device="cuda", fused_adam=1).items():
    setattr(opt, k, v)
import lirec_b200.mlp.model as M

torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model, loss_fn, optimizer = M.create_model(101, n_rels=15)
dp.broadcast_params(model._flat)
fused = dp.SwitchReduceAdam.attach(model, optimizer, mode=os.environ.get('LIREC_DP_MODE', 'shard'))
n = model._flat.numel()
grad = model._flat_grad
grad.normal_(generator=torch.Generator(device=grad.device).manual_seed(rank))
grad.mul_(1e-3)
ITERS, WARM = 50, 5


def timed(fn):
    for _ in range(WARM):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / ITERS], dtype=torch.float64, device=grad.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def nccl_only():
    dist.all_reduce(grad, op=dist.ReduceOp.SUM)
    grad.mul_(1.0 / world)            # keep the magnitudes bounded over the repeats (not part of a real step)


def scale_only():
    grad.mul_(1.0 / world)


def nccl_adam():
    dist.all_reduce(grad, op=dist.ReduceOp.SUM)
    optimizer.step(grad_scale=1.0 / world)
    grad.mul_(1.0 / world)


def switch_adam():
    fused.step()                      # leaves the SUM over ranks in `grad`
    grad.mul_(1.0 / world)


def switch_only():
    from lirec_b200 import ops
    ops.dp_exchange(fused._mc("grad"), 0, n, rank, world, fused._flag_ptrs.data_ptr(), 0)
    grad.mul_(1.0 / world)


def switch_two_buckets():
    from lirec_b200 import ops
    ops.dp_exchange(fused._mc("grad"), fused.split, n - fused.split, rank, world, fused._flag_ptrs.data_ptr(), 0)
    ops.dp_exchange(fused._mc("grad"), 0, fused.split, rank, world, fused._flag_ptrs.data_ptr(), 1)
    grad.mul_(1.0 / world)


res = {"world": world, "params": n, "bytes": 4 * n}
scale_ms = timed(scale_only)
res["adam_ms"] = timed(lambda: optimizer.step(grad_scale=1.0))
res["nccl_allreduce_ms"] = timed(nccl_only) - scale_ms
res["nccl_allreduce_plus_adam_ms"] = timed(nccl_adam) - scale_ms
if fused is not None:
    res["mode"] = fused.mode
    res["switch_reduce_adam_ms"] = timed(switch_adam) - scale_ms
    res["switch_exchange_only_ms"] = timed(switch_only) - scale_ms
    res["switch_exchange_two_buckets_ms"] = timed(switch_two_buckets) - scale_ms
    res["switch_exchange_only_busbw_GBs"] = 2.0 * (world - 1) / world * 4 * n / (res["switch_exchange_only_ms"] * 1e-3) / 1e9
bus = 2.0 * (world - 1) / world * 4 * n
res["nccl_busbw_GBs"] = bus / (res["nccl_allreduce_ms"] * 1e-3) / 1e9
if fused is not None:
    ex = res["switch_reduce_adam_ms"] - res["adam_ms"]
    res["switch_exchange_ms"] = ex
    res["switch_busbw_GBs"] = bus / (ex * 1e-3) / 1e9 if ex > 0 else None
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
