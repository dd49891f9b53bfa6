"""Dataset + packed dataloader (reference: mixed_utils/classification_dataloader.py).

The reference's `MixedFeaturesDataset.__getitem__` (:291-616) builds a dense zero-padded
float64 `[20, 19, 6912]` block per clip (21 MB, mostly np.tile copies) and torch's default collate
stacks them — 1.34 GB per 64-clip batch through worker IPC.  Here an item is a small clip RECORD
(its own pooled vectors plus index triples in the reference's slot order) and the collate function
packs records into one PackedBatch whose tables live in pinned memory; `packed_loader` copies batch
i+1 to the GPU on a side stream while batch i computes.

The MovieGraphs annotations / feature dump (~80 GB) are not available offline, and their parsing
(utils/util_functions.py, moviegraphs/py3loader) is outside the hot path:
without `opt.synthetic` the dataset refuses to construct instead of pretending.
"""
import torch
from torch.utils.data import Dataset

from lirec_b200.mixed_utils import synthetic
from lirec_b200.utils.arg_pars import opt


def preset_from_opt():
    if opt.mod_check:
        return "modalities"
    if opt.tr_maximize:
        return "int_rel_ch" if (opt.ctx == 1 and opt.rels_multitask) else "int_ch"
    return "int_rels"


class MixedFeaturesDataset(Dataset):
    """Same constructor and attributes the loops use (`n_classes`, `n_rels`, `rels_list`, `cache()`,
    `init_relships()`, `epoch`), backed by the synthetic MovieGraphs-shaped generator."""
    SIZES = {"train": 4096, "val": 512, "test": 512}

    def __init__(self, mode="train", size=None):
        if not getattr(opt, "synthetic", 0):
            raise RuntimeError(
                "MixedFeaturesDataset needs the MovieGraphs feature dump and annotations (README.md:18-28 of "
                "the reference), which are not available offline; run with --synthetic 1 for the synthetic "
                "MovieGraphs-shaped dataset")
        self.mode = mode
        self.n_classes = synthetic.N_CLASSES
        self.rels_list = ["rel%02d" % i for i in range(synthetic.N_RELS)] + ["None"]
        self.n_rels = len(self.rels_list)
        self.interidx2mgdidx = list(range(self.n_classes))
        self._max_n_tripl = int(getattr(opt, "max_n_tripl", 20))
        self.rels_n_clips = opt.rels_n_clips if opt.rels_multi_clip else 18
        self.epoch = 0
        self._size = int(size or self.SIZES.get(mode, 512))
        self._base = {"train": 0, "val": 1, "test": 2}.get(mode, 3) * 10_000_019 + int(opt.seed) * 7919
        self.preset = preset_from_opt()

    def cache(self):
        """The reference precomputes pooled features here (:139-186); synthetic clips are generated on
        demand by the dataloader workers."""
        return self

    def init_relships(self):
        assert self.rels_list[-1] == "None"
        return self

    def __len__(self):
        return self._size

    def __getitem__(self, idx):
        return synthetic.make_clip(self._base + int(idx), preset=self.preset, max_n_tripl=self._max_n_tripl,
                                   rels_n_clips=self.rels_n_clips)


def collate_packed(records):
    """records -> pinned host PackedBatch (runs in the dataloader worker / main process)."""
    return synthetic.pack_clips(records)


def packed_loader(dataset, batch_size, shuffle, num_workers=0, device="cuda", rank=0, world=1, drop_last=False,
                  seed=0):
    """Iterate device-resident PackedBatches with one-batch-ahead async H2D prefetch.

    Data parallel: every rank iterates the same shuffled order and takes its contiguous share of each
    global batch (lirec_b200/dp.py:shard_range), so the global batch equals the single-GPU one."""
    from lirec_b200 import dp
    g = torch.Generator()
    g.manual_seed(int(seed) * 1000003 + int(getattr(dataset, "epoch", 0)))
    n = len(dataset)
    order = torch.randperm(n, generator=g).tolist() if shuffle else list(range(n))
    batches = []
    for s in range(0, n, batch_size):
        idx = order[s:s + batch_size]
        if drop_last and len(idx) < batch_size:
            break
        a, b = dp.shard_range(len(idx), rank, world)
        if b > a:
            batches.append((idx[a:b], len(idx)))
    loader = torch.utils.data.DataLoader(_IndexView(dataset, batches), batch_size=None, shuffle=False,
                                         num_workers=int(num_workers), collate_fn=None)
    copy_stream = torch.cuda.Stream(device=device)
    pending = None
    for host_pb in loader:
        with torch.cuda.stream(copy_stream):
            dev_pb = host_pb.pin().to_device(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        if pending is not None:
            prev, pev = pending
            torch.cuda.current_stream().wait_event(pev)
            yield prev.record_stream(torch.cuda.current_stream())
        pending = (dev_pb, ev)
    if pending is not None:
        prev, pev = pending
        torch.cuda.current_stream().wait_event(pev)
        yield prev.record_stream(torch.cuda.current_stream())


class _IndexView(Dataset):
    """One item = one packed (per-rank) batch, so workers do the packing too."""

    def __init__(self, dataset, batches):
        self.dataset, self.batches = dataset, batches

    def __len__(self):
        return len(self.batches)

    def __getitem__(self, i):
        idx, global_size = self.batches[i]
        pb = collate_packed([self.dataset[j] for j in idx])
        pb.global_clips = global_size
        return pb


def f_dataloader(mode="train"):
    """Reference: classification_dataloader.py:623-630."""
    print("load mixed features. mode: %s" % mode)
    dataset = MixedFeaturesDataset(mode)
    loader = packed_loader(dataset, opt.batch_size, shuffle=(mode == "train"), num_workers=opt.num_workers)
    return loader, dataset.n_classes
