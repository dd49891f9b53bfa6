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


def main():
    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                     rels_n_clips=18, mod_check=False, device="cuda", fused_adam=1, lr=1e-3).items():
        setattr(opt, k, v)
    import lirec_b200.mlp.model as M
    pbs = [synthetic.make_batch(48, seed=100 * rank + i).to_device("cuda") for i in range(3)]
    finals = []
    for mode in ("nccl", "switch"):
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model, loss_fn, optimizer = M.create_model(101, n_rels=15)
        model.train()
        dp.broadcast_params(model._flat)
        fused = dp.SwitchReduceAdam.attach(model, optimizer) if mode == "switch" else None
        if mode == "switch" and fused is None:
            if rank == 0:
                print("SKIP: no NVSwitch multicast support on this box")
            return 0
        for i, pb in enumerate(pbs):
            lv = loss_fn(model(pb, seed=7 + i), {})
            optimizer.zero_grad()
            lv.backward()
            dp.reduce_and_step(model, optimizer, fused)
        torch.cuda.synchronize()
        finals.append(dict(p=model._flat.clone(), m=optimizer._m.clone(), v=optimizer._v.clone(),
                           pb=model._flat_bf16.float().clone(), g=model._flat_grad.clone()))
    ok = True
    for k in ("g", "p", "m", "v", "pb"):
        a, b = finals[0][k], finals[1][k]
        err = float((a - b).abs().max() / (a.abs().max() + 1e-30))
        tol = 4e-3 if k == "pb" else 2e-6
        if rank == 0:
            print("%-3s max-norm relative difference switch vs nccl: %.2e" % (k, err))
        ok = ok and err < tol
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
