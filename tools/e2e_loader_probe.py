"""End to end THROUGH the dataset: `packed_loader` (index-only items -> native collate in worker processes ->
pinned arena -> async H2D -> device row gather from the HBM-resident banks) feeding `train_step`, on the
synthetic annotation world (`--synthetic 2`, the path real MovieGraphs annotations take), against the same
batches staged ahead of time.  Needs a B200.

    python tools/e2e_loader_probe.py [--batch 1024] [--workers 8] [--movies 24] [--scenes 60] [--epochs 3]

Prints one JSON line: clips/s with the loader in the loop, clips/s with pre-staged batches, loader-only
clips/s on the host, and the batch statistics.  (bench.py's e2e legs time pre-collated batches of the
independent-clips workload; this probe answers whether the host side keeps up.)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--workers", type=int, default=8)
ap.add_argument("--movies", type=int, default=24)
ap.add_argument("--scenes", type=int, default=60)
ap.add_argument("--epochs", type=int, default=3)
ap.add_argument("--resident", type=int, default=1)
args = ap.parse_args()
sys.argv = sys.argv[:1]

import contextlib
import io

import numpy as np
import torch

from lirec_b200.utils.arg_pars import opt

for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                 rels_n_clips=18, mod_check=False, device="cuda", fused_adam=1, synthetic=2, world_movies=args.movies,
                 world_scenes=args.scenes, resident_banks=args.resident, batch_size=args.batch,
                 num_workers=args.workers, inter_class="all", merged=True, multilab_weights=True, rels=False,
                 soft_gt=False, native_step=1).items():
    setattr(opt, k, v)

from lirec_b200.mixed_utils import classification_dataloader as cd
import lirec_b200.mlp.model as M
import lirec_b200.mlp.train as TR

ds = cd.MixedFeaturesDataset("train").cache().init_relships()
t0 = time.perf_counter()
ds.warm_records()
warm_s = time.perf_counter() - t0
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model, loss_fn, optimizer = M.create_model(ds.n_classes, n_rels=ds.n_rels - 1)
model.train()
dev = torch.device("cuda", 0)


def epoch(loader_iter):
    n = 0
    last = None
    for pb in loader_iter:
        last = TR.train_step(model, loss_fn, optimizer, pb)
        n += pb.B
    return n, last


def loader():
    return cd.packed_loader(ds, args.batch, shuffle=True, num_workers=args.workers, device=dev, drop_last=True,
                            seed=0)


# host side alone: items + collate, no GPU work (one process)
t0 = time.perf_counter()
n_host = 0
for s in range(0, min(len(ds), 4 * args.batch), args.batch):
    idx = list(range(s, min(s + args.batch, len(ds))))
    ds.collate([ds[i] for i in idx])
    n_host += len(idx)
host_rate = n_host / (time.perf_counter() - t0)

n, last = epoch(loader())                 # warm-up epoch (allocator, worker start-up)
torch.cuda.synchronize()
t0 = time.perf_counter()
clips = 0
for _ in range(args.epochs):
    n, last = epoch(loader())
    clips += n
float(last.item())
torch.cuda.synchronize()
loader_rate = clips / (time.perf_counter() - t0)

staged = list(loader())                   # the same kind of batches, already on the device
torch.cuda.synchronize()
t0 = time.perf_counter()
clips2 = 0
for _ in range(args.epochs):
    n, last = epoch(iter(staged))
    clips2 += n
float(last.item())
torch.cuda.synchronize()
staged_rate = clips2 / (time.perf_counter() - t0)
pb = staged[0]
print(json.dumps({"dataset_items": len(ds), "clips_per_batch": args.batch, "workers": args.workers,
                  "resident_banks": args.resident, "warm_records_s": warm_s,
                  "candidate_rows_per_batch": pb.n_cand, "context_rows_per_batch": pb.n_ctx_rows,
                  "clip_rows": pb.n_clip, "track_rows": pb.n_track,
                  "e2e_with_loader_clips_s": loader_rate, "prestaged_clips_s": staged_rate,
                  "host_items_plus_collate_clips_s_per_core": host_rate, "host_cores": os.cpu_count()}))
