"""How long does the HOST take to enqueue one train step, vs the device time of the step?
If host >= device the step is launch-bound and a CUDA graph (or a leaner host path) pays."""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_argv = sys.argv[1:]; sys.argv = sys.argv[:1]
import torch
from lirec_b200.utils.arg_pars import opt
from lirec_b200.mixed_utils import synthetic
for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                 rels_n_clips=18, mod_check=False, device="cuda", fused_adam=1).items():
    setattr(opt, k, v)
import lirec_b200.mlp.model as M
B = int(_argv[0]) if _argv else 1024
with contextlib.redirect_stdout(io.StringIO()):
    model, loss_fn, optimizer = M.create_model(101, n_rels=15)
model.train()
pbs = [synthetic.make_batch(B, seed=i, preset="int_rel_ch").pin().to_device("cuda") for i in range(2)]

def step(pb):
    out = model(pb); lv = loss_fn(out, {}); optimizer.zero_grad(); lv.backward(); optimizer.step(); return lv

for i in range(5): step(pbs[i % 2])
torch.cuda.synchronize()
import cProfile, pstats
N = 50
t0 = time.perf_counter()
for i in range(N): step(pbs[i % 2])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("B=%d host enqueue %.3f ms/step, total (host+drain) %.3f ms/step" % (B, 1e3 * (t1 - t0) / N, 1e3 * (t2 - t0) / N))
# host-only cost: profile with the GPU idle between steps
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for i in range(20):
    step(pbs[i % 2]); torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(25)
