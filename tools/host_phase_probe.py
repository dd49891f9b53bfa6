"""Where does the HOST time of one train step go?  Phase timers (no device sync inside the loop, so
at small batches where the device is faster than the host these are pure enqueue costs) plus the
time spent inside each C-ABI call.

  python tools/host_phase_probe.py [clips_per_step] [steps]
"""
import collections, contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_argv = sys.argv[1:]; sys.argv = sys.argv[:1]
import torch
from lirec_b200 import _ext
from lirec_b200.utils.arg_pars import opt
from lirec_b200.mixed_utils import synthetic
for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                 rels_n_clips=18, mod_check=False, device="cuda", fused_adam=1).items():
    setattr(opt, k, v)
import lirec_b200.mlp.model as M
B = int(_argv[0]) if _argv else 64
N = int(_argv[1]) if len(_argv) > 1 else 300
with contextlib.redirect_stdout(io.StringIO()):
    model, loss_fn, optimizer = M.create_model(101, n_rels=15)
model.train()
pbs = [synthetic.make_batch(B, seed=i, preset="int_rel_ch").pin().to_device("cuda") for i in range(4)]

c_time = collections.defaultdict(float)
c_calls = collections.defaultdict(int)


class TimedLib(object):
    def __init__(self, L):
        object.__setattr__(self, "_L", L)
        object.__setattr__(self, "_cache", {})

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._L, name)

            def fn(*a, _raw=raw, _name=name):
                t = time.perf_counter()
                r = _raw(*a)
                c_time[_name] += time.perf_counter() - t
                c_calls[_name] += 1
                return r
            self._cache[name] = fn
        return fn


phase = collections.defaultdict(float)


import lirec_b200.mlp.train as TR
NATIVE = os.environ.get("PROBE_NATIVE", "0") == "1"


def step(pb, timed):
    if NATIVE:
        t0 = time.perf_counter()
        TR.train_step(model, loss_fn, optimizer, pb)
        if timed:
            phase["native train_step"] += time.perf_counter() - t0
        return
    t0 = time.perf_counter()
    out = model(pb)
    t1 = time.perf_counter()
    lv = loss_fn(out, {})
    t2 = time.perf_counter()
    optimizer.zero_grad()
    t3 = time.perf_counter()
    lv.backward()
    t4 = time.perf_counter()
    optimizer.step()
    t5 = time.perf_counter()
    if timed:
        for k, v in (("forward", t1 - t0), ("loss", t2 - t1), ("zero_grad", t3 - t2), ("backward", t4 - t3),
                     ("optimizer.step", t5 - t4)):
            phase[k] += v


for i in range(10):
    step(pbs[i % 4], False)
torch.cuda.synchronize()

# 1) plain loop: enqueue time vs total
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(N):
    step(pbs[i % 4], True)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("B=%d  host enqueue %.3f ms/step   wall incl. drain %.3f ms/step   device span %.3f ms/step" % (
    B, 1e3 * (t1 - t0) / N, 1e3 * (t2 - t0) / N, e0.elapsed_time(e1) / N))
for k, v in phase.items():
    print("  phase %-16s %7.1f us/step" % (k, 1e6 * v / N))

# 2) time inside the C ABI
_ext._lib = TimedLib(_ext.lib())
for i in range(N):
    step(pbs[i % 4], False)
torch.cuda.synchronize()
tot = 0.0
for k in sorted(c_time, key=lambda k: -c_time[k]):
    print("  C call %-32s %6.1f us/step  (%.1f calls/step)" % (k, 1e6 * c_time[k] / N, c_calls[k] / N))
    tot += c_time[k]
print("  C calls total %.1f us/step" % (1e6 * tot / N))
_ext._lib = _ext._lib._L

# 3) device-only time of a step: sync before, events around one step
dev = []
for i in range(50):
    torch.cuda.synchronize()
    e0.record(); step(pbs[i % 4], False); e1.record(); torch.cuda.synchronize()
    dev.append(e0.elapsed_time(e1))
dev.sort()
print("  single-step device span (host-bound, includes launch gaps): median %.3f ms  min %.3f ms" % (dev[len(dev) // 2], dev[0]))
