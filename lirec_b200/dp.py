"""Data parallelism over clips: one process per GPU, NCCL allreduce of the flat gradient buffer.

The reference is single-process; clips are independent and every loss is a
per-clip term averaged over the batch (mlp/model.py:402, 418, 439, 493, 561-573), so the only
exchange step of the hot path is the gradient sum.  Because all parameters live in ONE flat
fp32 gradient buffer (lirec_b200/mlp/model.py), that is a single `all_reduce` over NVLink /
NVSwitch per step; the 1/world_size average is folded into the fused Adam kernel's grad_scale
(or applied in place for torch.optim.Adam).

`SwitchReduceAdam` goes one step further on NVSwitch boxes: the gradient buffer lives in symmetric
memory mapped into a multicast object, and ONE kernel per rank (csrc/dp.cu:lirec_dp_allreduce_adam)
reduces the gradients inside the switch (multimem.ld_reduce / multimem.st) and runs Adam — no NCCL
call on the step's critical path.  torch.distributed._symmetric_memory only allocates and
rendezvous-es the buffers.
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


class SwitchReduceAdam:
    """In-switch gradient reduction fused with the flat Adam step (needs NVSwitch multicast).

        fused = SwitchReduceAdam.attach(model, optimizer)      # collective; None if unsupported
        ...
        lv.backward()
        fused.step()                                           # instead of all_reduce + optimizer.step
    """

    def __init__(self, model, optimizer, grad, hdl, flags, flag_hdl):
        self.model, self.optimizer = model, optimizer
        self.grad, self.hdl, self.flags, self.flag_hdl = grad, hdl, flags, flag_hdl
        self.ws = torch.zeros(4, dtype=torch.int32, device=grad.device)
        self.epoch = 0
        self.rank, self.world = hdl.rank, hdl.world_size

    @staticmethod
    def supported(device):
        try:
            from torch._C._autograd import DeviceType
            from torch._C._distributed_c10d import _SymmetricMemory
            return bool(_SymmetricMemory.has_multicast_support(DeviceType.CUDA, torch.device(device).index or 0))
        except Exception:
            return False

    @classmethod
    def attach(cls, model, optimizer):
        """Move the model's flat gradient buffer into symmetric memory (collective over the world group).
        Returns None — and leaves everything as it was — when there is a single rank or no multicast."""
        from lirec_b200.mlp.model import FlatAdam
        if world_size() < 2 or not isinstance(optimizer, FlatAdam):
            return None
        model._sync_flat()
        dev = model._flat.device
        ok = torch.tensor([1 if cls.supported(dev) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            return None
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD
        n = model._flat.numel()
        grad = symm_mem.empty(n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(grad, group)
        flags = symm_mem.empty(max(64, 2 * dist.get_world_size()), dtype=torch.int32, device=dev)
        flag_hdl = symm_mem.rendezvous(flags, group)
        if not hdl.multicast_ptr:
            return None
        grad.zero_()
        flags.zero_()
        model.use_grad_buffer(grad)
        torch.cuda.synchronize(dev)
        dist.barrier()
        return cls(model, optimizer, grad, hdl, flags, flag_hdl)

    @torch.no_grad()
    def step(self, local_clips=None, global_clips=None):
        from lirec_b200 import ops
        m, o = self.model, self.optimizer
        scale = 1.0 / self.world
        if local_clips is not None and global_clips is not None and local_clips * self.world != global_clips:
            self.grad.mul_(float(local_clips) / float(global_clips))      # unequal last shards: weighted sum
            scale = 1.0
        g = o.param_groups[0]
        o._t += 1
        self.epoch += 1
        ops.dp_allreduce_adam(m._flat, self.grad, self.hdl.multicast_ptr, o._m, o._v, m._flat_bf16, g["lr"],
                              g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], o._t, scale, self.rank,
                              self.world, self.flag_hdl.buffer_ptrs_dev, self.ws, self.epoch)
        m.mark_bf16_fresh()          # FlatAdam materialises the per-parameter step counters lazily


def reduce_and_step(model, optimizer, fused=None, local_clips=None, global_clips=None):
    """Gradient exchange + optimizer step of one data-parallel iteration: the in-switch fused kernel when
    `fused` (SwitchReduceAdam.attach) is available, else NCCL all_reduce + the optimizer's own step."""
    if fused is not None:
        fused.step(local_clips, global_clips)
        return
    if world_size() > 1:
        from lirec_b200.mlp.model import FlatAdam
        if isinstance(optimizer, FlatAdam):
            optimizer.step(grad_scale=allreduce_flat_grad(model._flat_grad, local_clips, global_clips))
        else:
            allreduce_flat_grad(model._flat_grad, local_clips, global_clips, average_in_place=True)
            optimizer.step()
    else:
        step = getattr(optimizer, "_step_impl", None)       # FlatAdam: skip the Optimizer.step hook wrapper
        if step is not None:
            step()
        else:
            optimizer.step()
