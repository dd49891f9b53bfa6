"""Data parallelism over clips: one process per GPU, NCCL allreduce of the flat gradient buffer.

The reference is single-process; clips are independent and every loss is a
per-clip term averaged over the batch (mlp/model.py:402, 418, 439, 493, 561-573), so the only
exchange step of the hot path is the gradient sum.  Because all parameters live in ONE flat
fp32 gradient buffer (lirec_b200/mlp/model.py), that is a single `all_reduce` over NVLink /
NVSwitch per step; the 1/world_size average is folded into the fused Adam kernel's grad_scale
(or applied in place for torch.optim.Adam).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_range(n_items, rank, world):
    """Contiguous, near-even split of n_items over ranks: [begin, end)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_flat_grad(flat_grad, local_clips=None, global_clips=None, average_in_place=False):
    """Sum the flat gradient over ranks.  Per-rank losses are means over LOCAL clips; with equal
    shards the global-batch gradient is the plain average (returned scale = 1/world).  For unequal
    shards (last short batch) gradients are pre-scaled by local/global clip counts and summed
    (returned scale = 1)."""
    world = world_size()
    if world == 1:
        return 1.0
    scale = 1.0 / world
    if local_clips is not None and global_clips is not None and local_clips * world != global_clips:
        flat_grad.mul_(float(local_clips) / float(global_clips))
        scale = 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    if average_in_place and scale != 1.0:
        flat_grad.mul_(scale)
        return 1.0
    return scale


def broadcast_params(flat_param, src=0):
    if world_size() > 1:
        dist.broadcast(flat_param, src=src)
