"""Synthetic MovieGraphs-shaped clips and batches (the 80 GB feature dump is not available offline).

`make_clip(seed)` generates ONE clip the way SURVEY.md §8d specifies and lays its candidate slots out
in the reference's order (mixed_utils/classification_dataloader.py:339-573):

  slot 0                ground-truth triplet (with its relationship context)
  then                  every ordered pair of the clip's characters except the GT pair
                        (itertools.permutations order); a bidirectional GT records
                        gt_tracks[1] = mem_counter - 1 when it meets the reversed pair (:451-453)
  then (1-person GT)    the GT person in the other slot (:513-540)
  then                  two single-person rows per character (:543-573), guarded by
                        mem_counter < max_n_tripl - 1 (:555)

Every row is an index triple into the clip's local clip / track vectors; context rows of a pair with
a relationship point at that pair's context clips, candidates without one carry the single self row
the reference tiles (:412-416, 477-478, 496-497, 531-533, 558-565).  `pack_clips(records)` is the
collate function: it concatenates clip records into one PackedBatch (lirec_b200/packing.py), batch
clips first in both banks so the ints branch works on a prefix.
"""
from itertools import permutations

import numpy as np

from ..packing import CLIP_DIM, TEXT_DIM, TRACK_DIM, PackedBatch

N_CLASSES, N_RELS = 101, 15

PRESETS = {
    # name: model kind, candidate tracks enumerated, context branch
    "modalities": dict(kind="modalities", enumerate_tracks=False, ctx=False),
    "int_rels": dict(kind="midfusion", enumerate_tracks=False, ctx=True),
    "int_ch": dict(kind="maxtracks", enumerate_tracks=True, ctx=False),
    "int_rel_ch": dict(kind="maxtracks", enumerate_tracks=True, ctx=True),
}
DEFAULT_N_CHARS = {1: 0.1, 2: 0.6, 3: 0.2, 4: 0.1}


def _features(rng, n, dim, nonneg):
    x = rng.standard_normal((n, dim), dtype=np.float32)
    return np.abs(x, out=x) if nonneg else x


def make_clip(seed, preset="int_rel_ch", max_n_tripl=20, rels_n_clips=18, n_chars_probs=None, p_rel=0.7,
              p_zero_track=0.1, p_bi=0.3, p_single_gt=0.1, n_classes=N_CLASSES, n_rels=N_RELS):
    """One clip record (host numpy).  Local index spaces: clip vectors — 0 is the clip itself, 1.. its
    context clips; track vectors — 0 is the clip's all-zero 'no track' row, then its characters, then
    the tracks of the context clips."""
    cfg = PRESETS[preset]
    rng = np.random.default_rng(seed)
    probs = n_chars_probs or DEFAULT_N_CHARS
    n_opts = np.array(sorted(probs))
    n_p = np.array([probs[k] for k in n_opts], dtype=np.float64)
    n = int(rng.choice(n_opts, p=n_p / n_p.sum()))
    T, S, NONE = int(max_n_tripl), int(rels_n_clips), n_rels

    # every clip owns one all-zero track row: a single shared zero row would be referenced by
    # thousands of candidate rows and serialise the backward scatter-reduce
    tr, n_own = [], 1
    for _ in range(n):
        if rng.random() < p_zero_track:
            tr.append(0)
        else:
            tr.append(n_own)
            n_own += 1
    single = (n == 1) or (rng.random() < p_single_gt)
    if single:
        gt, gt_pos = (int(rng.integers(n)),), int(rng.integers(2))
    else:
        i, j = rng.choice(n, size=2, replace=False)
        gt, gt_pos = (int(i), int(j)), 0
    bi = bool(rng.random() < p_bi)

    slots, gt_idx = [], [0, 0]               # (person or None, person or None) per candidate slot
    if len(gt) == 2:
        slots.append(gt)
        for (i, j) in permutations(range(n), 2):
            if (i, j) == gt:
                continue
            if bi and (i, j) == (gt[1], gt[0]):
                gt_idx[1] = len(slots) - 1                      # the reference's off-by-one (:453)
            if len(slots) < T:
                slots.append((i, j))
    else:
        g = gt[0]
        slots.append((g, None) if gt_pos == 0 else (None, g))
        for (i, j) in permutations(range(n), 2):
            if len(slots) < T:
                slots.append((i, j))
        if len(slots) < T:
            if bi:
                gt_idx[1] = len(slots)
            slots.append((None, g) if gt_pos == 0 else (g, None))
    if cfg["enumerate_tracks"]:
        for i in range(n):
            if len(gt) == 1 and i == gt[0]:
                continue
            if len(slots) < T - 1:
                slots.append((i, None))
                slots.append((None, i))
    else:
        slots, gt_idx = slots[:1], [0, 0]

    cand_rows, rels_label, ctx_rows, ctx_counts = [], [], [], []
    pair_ctx, n_ctx_clips, n_ctx_tracks = {}, 0, 0
    for (i, j) in slots:
        row = (0, 0 if i is None else tr[i], 0 if j is None else tr[j])
        cand_rows.append(row)
        lab, rows = NONE, [row]                                 # default: the tiled self row
        if cfg["ctx"] and i is not None and j is not None:
            key = (min(i, j), max(i, j))
            if key not in pair_ctx:
                L = int(rng.integers(1, S + 1)) if rng.random() < p_rel else 0
                pair_ctx[key] = dict(L=L, clip0=n_ctx_clips, tr0=n_ctx_tracks,
                                     lab={True: int(rng.integers(n_rels)), False: int(rng.integers(n_rels))})
                n_ctx_clips += L
                n_ctx_tracks += 2 * L
            pc = pair_ctx[key]
            if pc["L"] > 0:
                lab, rows = pc["lab"][i < j], []
                for l in range(pc["L"]):
                    ta = n_own + pc["tr0"] + 2 * l              # track of min(i,j) in context clip l
                    rows.append((1 + pc["clip0"] + l, ta if i < j else ta + 1, ta + 1 if i < j else ta))
        rels_label.append(lab)
        if cfg["ctx"]:
            ctx_rows += rows
            ctx_counts.append(len(rows))

    clip_vecs = np.empty((1 + n_ctx_clips, CLIP_DIM), dtype=np.float32)
    clip_vecs[:, :TEXT_DIM] = _features(rng, 1 + n_ctx_clips, TEXT_DIM, nonneg=False)
    clip_vecs[:, TEXT_DIM:] = _features(rng, 1 + n_ctx_clips, CLIP_DIM - TEXT_DIM, nonneg=True)
    track_vecs = _features(rng, n_own + n_ctx_tracks, TRACK_DIM, nonneg=True)
    track_vecs[0] = 0.0
    return dict(
        clip_vecs=clip_vecs, track_vecs=track_vecs, n_own_tracks=n_own,
        cand_rows=np.asarray(cand_rows, dtype=np.int64).reshape(-1, 3),
        ctx_rows=np.asarray(ctx_rows, dtype=np.int64).reshape(-1, 3) if cfg["ctx"] else None,
        ctx_counts=np.asarray(ctx_counts, dtype=np.int64) if cfg["ctx"] else None,
        rels_label=np.asarray(rels_label, dtype=np.int64), gt_tracks=np.asarray(gt_idx, dtype=np.int64),
        label=int(rng.integers(n_classes)), multilab=(rng.random(n_classes) < 0.95).astype(np.uint8),
        n_names=n, just_zeros=bool(all(t == 0 for t in tr)), n_slots=T if cfg["enumerate_tracks"] else 1,
        n_ctx_slots=S if cfg["ctx"] else 0, preset=preset)


def pack_clips(records):
    """Collate clip records into one host PackedBatch (the packed replacement of torch's default
    collate over the reference's dense items, mlp/train.py:33-37)."""
    B = len(records)
    has_ctx = records[0]["ctx_rows"] is not None
    # bank layout: [clip 0 .. clip B-1 | context clips]   [own tracks of all clips | context tracks]
    own_tr = np.array([r["n_own_tracks"] for r in records])
    ctx_cl = np.array([r["clip_vecs"].shape[0] - 1 for r in records])
    ctx_tr = np.array([r["track_vecs"].shape[0] - r["n_own_tracks"] for r in records])
    own_tr_off = np.concatenate(([0], np.cumsum(own_tr)))
    n_clip_ints, n_track_ints = B, int(own_tr_off[-1])
    ctx_cl_off = n_clip_ints + np.concatenate(([0], np.cumsum(ctx_cl)))
    ctx_tr_off = n_track_ints + np.concatenate(([0], np.cumsum(ctx_tr)))
    n_clip = int(ctx_cl_off[-1]) if has_ctx else n_clip_ints
    n_track = int(ctx_tr_off[-1]) if has_ctx else n_track_ints
    clip_bank = np.empty((n_clip, CLIP_DIM), dtype=np.float32)
    track_bank = np.empty((n_track, TRACK_DIM), dtype=np.float32)
    cand_rows, ctx_rows, ctx_counts, cand_counts = [], [], [], []
    for b, r in enumerate(records):
        no = r["n_own_tracks"]
        clip_bank[b] = r["clip_vecs"][0]
        track_bank[own_tr_off[b]:own_tr_off[b + 1]] = r["track_vecs"][:no]
        if has_ctx:
            clip_bank[ctx_cl_off[b]:ctx_cl_off[b + 1]] = r["clip_vecs"][1:]
            track_bank[ctx_tr_off[b]:ctx_tr_off[b + 1]] = r["track_vecs"][no:]

        def remap(rows, b=b, no=no):
            out = np.empty_like(rows)
            c, t = rows[:, 0], rows[:, 1:]
            out[:, 0] = np.where(c == 0, b, ctx_cl_off[b] + c - 1)
            out[:, 1:] = np.where(t < no, own_tr_off[b] + t, ctx_tr_off[b] + t - no)
            return out
        cand_rows.append(remap(r["cand_rows"]))
        cand_counts.append(len(r["cand_rows"]))
        if has_ctx:
            ctx_rows.append(remap(r["ctx_rows"]))
            ctx_counts.append(r["ctx_counts"])
    cand_off = np.concatenate(([0], np.cumsum(cand_counts)))
    ctx_off = ctx_tbl = None
    if has_ctx:
        ctx_off = np.concatenate(([0], np.cumsum(np.concatenate(ctx_counts))))
        ctx_tbl = np.concatenate(ctx_rows)
    pb = PackedBatch.from_tables(
        clip_bank, track_bank, n_clip_ints, n_track_ints, cand_off, np.concatenate(cand_rows), ctx_off, ctx_tbl,
        [r["label"] for r in records], np.concatenate([r["rels_label"] for r in records]) if has_ctx else None,
        np.stack([r["gt_tracks"] for r in records]), np.stack([r["multilab"] for r in records]),
        n_slots=records[0]["n_slots"], n_ctx_slots=records[0]["n_ctx_slots"],
        extras={"n_names": np.array([r["n_names"] for r in records]),
                "just_zeros": np.array([r["just_zeros"] for r in records])})
    pb.preset = records[0]["preset"]
    pb.kind = PRESETS[pb.preset]["kind"]
    return pb


def make_batch(B, seed=0, preset="int_rel_ch", **kw):
    """B synthetic clips (clip seeds seed*1_000_003 + i) packed into one host PackedBatch."""
    return pack_clips([make_clip(seed * 1000003 + i, preset=preset, **kw) for i in range(B)])


def stress_batch(B, seed=0):
    """Config 5: 4x context rows and 4x candidate slots per clip."""
    probs = {k: 1.0 / 7 for k in range(2, 9)}
    return make_batch(B, seed=seed, preset="int_rel_ch", max_n_tripl=80, rels_n_clips=72, n_chars_probs=probs)
