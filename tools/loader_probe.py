"""Host-side batch assembly throughput of the index-only dataset (no GPU needed):
__getitem__ per clip and collate_indexed per batch on the synthetic annotation world.

    python tools/loader_probe.py
"""
import os, sys, time; sys.argv=['x']
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lirec_b200.utils.arg_pars import opt
for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                 rels_n_clips=18, mod_check=False, synthetic=2, world_movies=6, world_scenes=40).items():
    setattr(opt, k, v)
from lirec_b200.mixed_utils import classification_dataloader as cd, indexed_dataset as ix
ds = cd.MixedFeaturesDataset("train")
ds = ds.cache().init_relships() if hasattr(ds, "cache") else ds
print(type(ds).__name__, len(ds))
t0 = time.perf_counter()
if hasattr(ds, "warm_records"):
    ds.warm_records()          # what packed_loader does before forking workers: every item's decision tree
print("warm_records %.1f ms (%.2f ms/clip, once)" % (1e3 * (time.perf_counter() - t0), 1e3 * (time.perf_counter() - t0) / len(ds)))
def best(f, n=7):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return min(ts), r
for B in (64, 256, 1024):
    idx = [i % len(ds) for i in range(B)]
    tg, recs = best(lambda: [ds[i] for i in idx])
    for resident in (False, True):
        tc, pb = best(lambda: ix.collate_indexed(recs, ds, resident=resident))
        print("B=%d getitem %.1f us/clip  collate(resident=%s) %.2f ms/batch (%.1f us/clip) -> %.0f clips/s/core (best of 7)" % (
            B, 1e6 * tg / B, resident, 1e3 * tc, 1e6 * tc / B, B / (tg + tc)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
recs = [ds[i] for i in range(min(256, len(ds)))]; pb = ix.collate_indexed(recs, ds, resident=True)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
