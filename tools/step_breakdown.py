"""Warm per-kernel device time of the int_rel_ch train step (torch.profiler / CUPTI), averaged over steps.
    python tools/step_breakdown.py [steps] [batch]
A diagnostic: numbers taken under a profiler are not bench values."""
import contextlib
import io
import os
import sys
from collections import defaultdict

sys.argv, ARGS = sys.argv[:1], sys.argv[1:]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from lirec_b200.mixed_utils import synthetic  # noqa: E402
from lirec_b200.utils.arg_pars import opt  # noqa: E402

steps = int(ARGS[0]) if ARGS else 20
batch = int(ARGS[1]) if len(ARGS) > 1 else 1024
for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                 rels_n_clips=18, mod_check=False, device="cuda", fused_adam=1).items():
    setattr(opt, k, v)
import lirec_b200.mlp.model as M  # noqa: E402

torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model, loss_fn, optimizer = M.create_model(101, n_rels=15)
model.train()
pbs = [synthetic.make_batch(batch, seed=i).pin().to_device("cuda") for i in range(2)]


def step(pb):
    lv = loss_fn(model(pb), {})
    optimizer.zero_grad()
    lv.backward()
    optimizer.step()


for i in range(5):
    step(pbs[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        step(pbs[i % 2])
    torch.cuda.synchronize()
seq = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
seq.sort(key=lambda e: e.time_range.start)
per = len(seq) // steps
tot = defaultdict(float)
print("# %d kernels per step; launch order of the last step (us, mean over %d steps)" % (per, steps))
for j in range(per):
    evs = seq[j::per]
    us = sum(e.device_time for e in evs) / len(evs)
    name = evs[-1].name.split("(")[0][:70]
    tot[name] += us
    print("%2d %-72s %8.1f" % (j, name, us))
span = (seq[-1].time_range.end - seq[-per].time_range.start)
print("# sum of kernels %.1f us; last step span %.1f us" % (sum(tot.values()), span))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-72s %8.1f" % (k, v))
