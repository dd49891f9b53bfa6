"""Independent synthetic clips behind the reference's dataset surface, with the features CACHED in two
dataset-level banks — the shape of the reference's real pipeline: `MixedFeaturesDataset.cache()` pools every
clip / track vector once (mixed_utils/classification_dataloader.py:139-186, mixed_features.py:37-112) and
`__getitem__` (:291-616) then assembles rows out of cached vectors.

    --synthetic 3   `cache()` draws `size` clips with mixed_utils/synthetic.py:make_clip (the bench workload,
                    SURVEY.md §8d C1-C5) and stores their vectors once: clip bank = every clip's own row followed
                    by its context clips (+ one all-zero last row), track bank = row 0 all zero, then every
                    clip's character tracks and context tracks.  An item is the clip's index-only record in the
                    format of mixed_utils/indexed_dataset.py (candidate / context triples into the dataset
                    banks), so batches are built by the same native collate (`lirec_collate_tables`) and, with
                    `--resident_banks 1`, staged by the same device row gather as the annotation-world dataset.

This is the dataset `bench.py` drives its end-to-end leg through (`packed_loader` with worker processes)."""
import numpy as np
from torch.utils.data import Dataset

from lirec_b200.mixed_utils import synthetic
from lirec_b200.mixed_utils.indexed_dataset import collate_arrays, collate_indexed
from lirec_b200.packing import CLIP_DIM, TRACK_DIM
from lirec_b200.utils.arg_pars import opt


class CachedClipsDataset(Dataset):
    SIZES = {"train": 4096, "val": 512, "test": 512}

    def __init__(self, mode="train", size=None, preset=None, seed_base=None, max_n_tripl=None, rels_n_clips=None,
                 clip_kwargs=None):
        from lirec_b200.mixed_utils.classification_dataloader import preset_from_opt
        self.mode = mode
        self.n_classes = synthetic.N_CLASSES
        self.rels_list = ["rel%02d" % i for i in range(synthetic.N_RELS)] + ["None"]
        self.n_rels = len(self.rels_list)
        self.interidx2mgdidx = list(range(self.n_classes))
        self._max_n_tripl = int(max_n_tripl or getattr(opt, "max_n_tripl", 20))
        self.rels_n_clips = int(rels_n_clips or (opt.rels_n_clips if opt.rels_multi_clip else 18))
        self.epoch = 0
        self._size = int(size or self.SIZES.get(mode, 512))
        self._base = int(seed_base) if seed_base is not None else \
            {"train": 0, "val": 1, "test": 2}.get(mode, 3) * 10_000_019 + int(opt.seed) * 7919
        self.preset = preset or preset_from_opt()
        self.clip_kwargs = dict(clip_kwargs or {})
        self.records = None
        self.clip_bank = self.track_bank = None
        self.zero_clip = -1

    # ---- the reference's cache(): every vector pooled once ---------------------------------------------
    def cache(self):
        if self.records is not None:
            return self
        clips = [synthetic.make_clip(self._base + i, preset=self.preset, max_n_tripl=self._max_n_tripl,
                                     rels_n_clips=self.rels_n_clips, **self.clip_kwargs) for i in range(self._size)]
        has_ctx = clips[0]["ctx_rows"] is not None
        track_models = synthetic.PRESETS[self.preset]["enumerate_tracks"]
        n_clip = sum(c["clip_vecs"].shape[0] if has_ctx else 1 for c in clips)
        n_track = sum((c["track_vecs"].shape[0] if has_ctx else c["n_own_tracks"]) - 1 for c in clips)
        self.clip_bank = np.zeros((n_clip + 1, CLIP_DIM), dtype=np.float32)       # last row: the all-zero clip
        self.track_bank = np.zeros((n_track + 1, TRACK_DIM), dtype=np.float32)    # row 0: the all-zero track
        self.zero_clip = n_clip
        self.records = []
        c0, t0 = 0, 1
        for c in clips:
            nc = c["clip_vecs"].shape[0] if has_ctx else 1
            nt = (c["track_vecs"].shape[0] if has_ctx else c["n_own_tracks"]) - 1
            self.clip_bank[c0:c0 + nc] = c["clip_vecs"][:nc]
            self.track_bank[t0:t0 + nt] = c["track_vecs"][1:1 + nt]

            def remap(rows, c0=c0, t0=t0):
                out = np.empty(rows.shape, dtype=np.int32)
                out[:, 0] = c0 + rows[:, 0]
                out[:, 1:] = np.where(rows[:, 1:] == 0, 0, t0 + rows[:, 1:] - 1)     # local 0 = "no track"
                return out
            rec = {"labels": int(c["label"]), "cand_rows": remap(c["cand_rows"]),
                   "multilab_weights": c["multilab"].astype(np.float64), "n_ctx_slots": c["n_ctx_slots"],
                   "n_names": c["n_names"], "just_zeros": c["just_zeros"]}
            if has_ctx:
                rec["ctx_cat"] = remap(c["ctx_rows"])
                rec["ctx_counts"] = c["ctx_counts"].astype(np.int32)
                rec["ctx_rows"] = None                       # marks a context batch (the blocks are ctx_cat / ctx_counts)
                rec["rels_label"] = c["rels_label"] if track_models else int(c["rels_label"][0])
            if track_models:
                rec["gt_tracks"] = c["gt_tracks"]
            self.records.append(rec)
            c0, t0 = c0 + nc, t0 + nt
        self._flatten()
        return self

    def _flatten(self):
        """Every record's arrays back to back in dataset-level tables (CSR by clip, and by candidate for the
        context blocks): a batch is then a handful of vectorised gathers instead of ~10 numpy calls per clip."""
        R = self.records
        has_ctx, track_models = "ctx_cat" in R[0], "gt_tracks" in R[0]
        n_cand = np.fromiter((len(r["cand_rows"]) for r in R), dtype=np.int64, count=len(R))
        self._cand_off = np.concatenate(([0], np.cumsum(n_cand)))
        self._cand = np.ascontiguousarray(np.concatenate([r["cand_rows"] for r in R]), dtype=np.int32)
        self._labels = np.array([r["labels"] for r in R], dtype=np.int32)
        self._multilab = np.ascontiguousarray(np.stack([r["multilab_weights"] for r in R]) != 0).astype(np.uint8)
        self._n_names = np.array([r["n_names"] for r in R])
        self._just_zeros = np.array([r["just_zeros"] for r in R])
        self._gt = np.stack([r["gt_tracks"] for r in R]).astype(np.int64) if track_models else None
        self._ctx = self._ctx_cnt = self._ctx_off = self._rels = None
        if has_ctx:
            self._ctx_cnt = np.ascontiguousarray(np.concatenate([r["ctx_counts"] for r in R]), dtype=np.int32)
            self._ctx_off = np.concatenate(([0], np.cumsum(self._ctx_cnt, dtype=np.int64)))
            self._ctx = np.ascontiguousarray(np.concatenate([r["ctx_cat"] for r in R]), dtype=np.int32)
            self._rels = np.concatenate([np.asarray(r["rels_label"]).reshape(-1) for r in R]).astype(np.int64) \
                if track_models else np.array([r["rels_label"] for r in R], dtype=np.int64)
        self._track_models = track_models
        self._cand_off = np.ascontiguousarray(self._cand_off, dtype=np.int64)
        self._max_cand = int(n_cand.max())
        self._max_ctx = int(self._ctx_cnt.max()) if has_ctx else 0

    @staticmethod
    def _ranges(starts, counts):
        """Concatenation of arange(starts[i], starts[i] + counts[i])."""
        total = int(counts.sum())
        ends = np.cumsum(counts)
        return np.repeat(starts - (ends - counts), counts) + np.arange(total, dtype=np.int64)

    def arena_bound(self, batch_size):
        """Upper bound of the int32 arena a batch of `batch_size` clips needs (for pre-allocated pinned slots)."""
        from lirec_b200 import _ext
        mc = batch_size * self._max_cand
        return int(_ext.lib().lirec_collate_arena_bound(batch_size, mc, mc * self._max_ctx, int(self._ctx is not None)))

    def get_batch(self, indices, arena_out=None, multilab_out=None):
        """`collate([self[i] for i in indices])` in one go — the same records, gathered from the dataset-level
        tables (equal batches: tests/test_dataloader_cpu.py).  Thread-safe (read-only on the dataset; the native
        calls release the GIL).  arena_out / multilab_out: pinned torch buffers the batch is built in directly."""
        if self.records is None:
            self.cache()
        from lirec_b200 import _ext
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        B = len(idx)
        # native ragged gather (lirec_collate_gather): candidate / context triples of the batch, back to back
        max_cand = B * self._max_cand
        max_ctx = max_cand * self._max_ctx if self._ctx is not None else 0
        cand = np.empty((max_cand, 3), dtype=np.int32)
        counts = np.empty(B, dtype=np.int32)
        cpos = np.empty(max_cand, dtype=np.int64)
        ctx = np.empty((max(max_ctx, 1), 3), dtype=np.int32) if self._ctx is not None else None
        ctx_counts = np.empty(max_cand, dtype=np.int32) if self._ctx is not None else None
        n = np.zeros(2, dtype=np.int64)
        _ext.check(_ext.lib().lirec_collate_gather(
            self._cand_off.ctypes.data, self._cand.ctypes.data,
            self._ctx_off.ctypes.data if ctx is not None else None, self._ctx_cnt.ctypes.data if ctx is not None else None,
            self._ctx.ctypes.data if ctx is not None else None, idx.ctypes.data, B, cand.ctypes.data, counts.ctypes.data,
            cpos.ctypes.data, max_cand, ctx.ctypes.data if ctx is not None else None,
            ctx_counts.ctypes.data if ctx is not None else None, max_ctx, n[0:].ctypes.data, n[1:].ctypes.data))
        ni, nx = int(n[0]), int(n[1])
        cand, cpos = cand[:ni], cpos[:ni]
        rels = None
        if ctx is not None:
            ctx, ctx_counts = ctx[:nx], ctx_counts[:ni]
            rels = self._rels[cpos] if self._track_models else self._rels[idx]
        gt = self._gt[idx] if self._gt is not None else np.zeros((len(idx), 2), dtype=np.int64)
        extras = {"just_zeros": self._just_zeros[idx], "n_names": self._n_names[idx]}
        ml = self._multilab[idx]
        pb = collate_arrays(self, np.ascontiguousarray(cand), np.ascontiguousarray(counts, dtype=np.int32),
                            None if ctx is None else np.ascontiguousarray(ctx),
                            None if ctx is None else np.ascontiguousarray(ctx_counts), self._labels[idx], rels, gt,
                            ml, extras, self._max_n_tripl if self._track_models else 1,
                            self.records[0]["n_ctx_slots"], bool(getattr(opt, "resident_banks", 0)), arena_out=arena_out)
        if multilab_out is not None:
            import torch
            out = multilab_out[:ml.size].view(ml.shape)
            out.copy_(torch.from_numpy(ml))
            pb.multilab = out
        pb.preset = self.preset
        pb.kind = synthetic.PRESETS[self.preset]["kind"]
        return pb

    def init_relships(self):
        assert self.rels_list[-1] == "None"
        return self

    def warm_records(self):
        return self.cache()

    def __len__(self):
        return self._size

    def __getitem__(self, idx):
        if self.records is None:
            self.cache()
        return self.records[int(idx)]

    def collate(self, records):
        pb = collate_indexed(records, self, resident=bool(getattr(opt, "resident_banks", 0)))
        pb.preset = self.preset
        pb.kind = synthetic.PRESETS[self.preset]["kind"]
        return pb
