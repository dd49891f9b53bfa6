"""Synthetic MovieGraphs-shaped batches (the 80 GB feature dump is not available offline).

Generates clips the way SURVEY.md §8d specifies and lays their candidate slots out in the
reference's order (mixed_utils/classification_dataloader.py:339-573):

  slot 0                ground-truth triplet (with its relationship context)
  then                  every ordered pair of the clip's characters except the GT pair
                        (itertools.permutations order); a bidirectional GT records
                        gt_tracks[1] = mem_counter - 1 when it meets the reversed pair (:451-453)
  then (1-person GT)    the GT person in the other slot (:513-540)
  then                  two single-person rows per character (:543-573), guarded by
                        mem_counter < max_n_tripl - 1 (:555)

Every row is an index triple into the clip / track banks (lirec_b200/packing.py); context rows
of a pair with a relationship point at that pair's context clips, candidates without one carry
the single self row the reference tiles (:412-416, 477-478, 496-497, 531-533, 558-565).
"""
from itertools import permutations

import numpy as np

from ..packing import CLIP_DIM, TEXT_DIM, TRACK_DIM, PackedBatch

N_CLASSES, N_RELS = 101, 15

PRESETS = {
    # name: (model kind, tracks enumerated, context branch)
    "modalities": dict(kind="modalities", enumerate_tracks=False, ctx=False),
    "int_rels": dict(kind="midfusion", enumerate_tracks=False, ctx=True),
    "int_ch": dict(kind="maxtracks", enumerate_tracks=True, ctx=False),
    "int_rel_ch": dict(kind="maxtracks", enumerate_tracks=True, ctx=True),
}


def _features(rng, n, dim, nonneg):
    x = rng.standard_normal((n, dim), dtype=np.float32)
    return np.abs(x, out=x) if nonneg else x


def make_batch(B, seed=0, preset="int_rel_ch", max_n_tripl=20, rels_n_clips=18, n_chars_probs=None,
               p_rel=0.7, p_zero_track=0.1, p_bi=0.3, p_single_gt=0.1, n_classes=N_CLASSES, n_rels=N_RELS,
               return_scenes=False):
    """One synthetic batch of B clips as a host PackedBatch."""
    cfg = PRESETS[preset]
    rng = np.random.default_rng(seed)
    if n_chars_probs is None:
        n_chars_probs = {1: 0.1, 2: 0.6, 3: 0.2, 4: 0.1}
    n_opts = np.array(sorted(n_chars_probs))
    n_p = np.array([n_chars_probs[k] for k in n_opts], dtype=np.float64)
    n_p /= n_p.sum()
    T, S = int(max_n_tripl), int(rels_n_clips)
    NONE = n_rels

    # bank bookkeeping: every clip owns one all-zero "no track" row (a single shared zero row would
    # be referenced by thousands of candidate rows and serialise the backward scatter-reduce)
    n_ints_tracks = 0
    person_track = []            # per clip: list of bank rows of its characters
    zero_row = []                # per clip: its all-zero track row
    clips = []
    for b in range(B):
        n = int(rng.choice(n_opts, p=n_p))
        zero_row.append(n_ints_tracks)
        n_ints_tracks += 1
        rows = []
        for _ in range(n):
            if rng.random() < p_zero_track:
                rows.append(zero_row[b])
            else:
                rows.append(n_ints_tracks)
                n_ints_tracks += 1
        person_track.append(rows)
        single = (n == 1) or (rng.random() < p_single_gt)
        if single:
            gt = (int(rng.integers(n)),)
            gt_pos = int(rng.integers(2))          # which slot the single GT person occupies
        else:
            i, j = rng.choice(n, size=2, replace=False)
            gt, gt_pos = (int(i), int(j)), 0
        clips.append(dict(n=n, gt=gt, gt_pos=gt_pos, bi=bool(rng.random() < p_bi)))

    cand_rows, cand_off, rels_label, gt_tracks = [], [0], [], []
    ctx_specs = []               # per candidate: None (self row) or (pair key, direction)
    ctx_clip_rows = 0            # context clips get bank rows after the B batch clips
    ctx_track_rows = 0
    pair_ctx = {}                # (clip, unordered pair) -> dict(L, clip0, tr0) and labels per direction
    for b, c in enumerate(clips):
        tr = person_track[b]
        n = c["n"]
        slots = []               # (person index or None, person index or None)
        gt_idx = [0, 0]
        if len(c["gt"]) == 2:
            slots.append((c["gt"][0], c["gt"][1]))
            for (i, j) in permutations(range(n), 2):
                if (i, j) == c["gt"]:
                    continue
                if c["bi"] and (i, j) == (c["gt"][1], c["gt"][0]):
                    gt_idx[1] = len(slots) - 1                      # the reference's off-by-one
                if len(slots) < T:
                    slots.append((i, j))
        else:
            g = c["gt"][0]
            slots.append((g, None) if c["gt_pos"] == 0 else (None, g))
            for (i, j) in permutations(range(n), 2):
                if len(slots) < T:
                    slots.append((i, j))
            if len(slots) < T:
                if c["bi"]:
                    gt_idx[1] = len(slots)
                slots.append((None, g) if c["gt_pos"] == 0 else (g, None))
        if cfg["enumerate_tracks"]:
            for i in range(n):
                if len(c["gt"]) == 1 and i == c["gt"][0]:
                    continue
                if len(slots) < T - 1:
                    slots.append((i, None))
                    slots.append((None, i))
        else:
            slots = slots[:1]
            gt_idx = [0, 0]
        for (i, j) in slots:
            cand_rows.append((b, zero_row[b] if i is None else tr[i], zero_row[b] if j is None else tr[j]))
            lab, spec = NONE, None
            if cfg["ctx"] and i is not None and j is not None:
                key = (b, min(i, j), max(i, j))
                if key not in pair_ctx:
                    has = rng.random() < p_rel
                    L = int(rng.integers(1, S + 1)) if has else 0
                    pair_ctx[key] = dict(L=L, clip0=ctx_clip_rows, tr0=ctx_track_rows,
                                         lab={True: int(rng.integers(n_rels)), False: int(rng.integers(n_rels))})
                    ctx_clip_rows += L
                    ctx_track_rows += 2 * L
                pc = pair_ctx[key]
                if pc["L"] > 0:
                    lab, spec = pc["lab"][i < j], (key, i < j)
            rels_label.append(lab)
            ctx_specs.append(spec)
        cand_off.append(len(cand_rows))
        gt_tracks.append(gt_idx)

    Ni = len(cand_rows)
    cand_rows = np.asarray(cand_rows, dtype=np.int64)
    n_clip_ints, n_track_ints = B, n_ints_tracks
    ctx_off = ctx_rows = None
    if cfg["ctx"]:
        ctx_off, rows = [0], []
        for r in range(Ni):
            spec = ctx_specs[r]
            if spec is None:
                rows.append(tuple(cand_rows[r]))                    # the tiled self row, rels_mask[0] = 1
            else:
                key, fwd = spec
                pc = pair_ctx[key]
                for l in range(pc["L"]):
                    ta = n_track_ints + pc["tr0"] + 2 * l           # track of min(i,j) in context clip l
                    tb = ta + 1                                     # track of max(i,j)
                    rows.append((n_clip_ints + pc["clip0"] + l, ta if fwd else tb, tb if fwd else ta))
            ctx_off.append(len(rows))
        ctx_rows = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
    n_clip = n_clip_ints + (ctx_clip_rows if cfg["ctx"] else 0)
    n_track = n_track_ints + (ctx_track_rows if cfg["ctx"] else 0)

    clip_bank = np.empty((n_clip, CLIP_DIM), dtype=np.float32)
    clip_bank[:, :TEXT_DIM] = _features(rng, n_clip, TEXT_DIM, nonneg=False)
    clip_bank[:, TEXT_DIM:] = _features(rng, n_clip, CLIP_DIM - TEXT_DIM, nonneg=True)
    track_bank = _features(rng, n_track, TRACK_DIM, nonneg=True)
    track_bank[np.asarray(zero_row)] = 0.0
    labels = rng.integers(n_classes, size=B)
    multilab = (rng.random((B, n_classes)) < 0.95).astype(np.uint8)
    pb = PackedBatch.from_tables(clip_bank, track_bank, n_clip_ints, n_track_ints, cand_off, cand_rows, ctx_off,
                                 ctx_rows, labels, rels_label if cfg["ctx"] else None, gt_tracks, multilab,
                                 n_slots=T if cfg["enumerate_tracks"] else 1, n_ctx_slots=S if cfg["ctx"] else 0)
    pb.kind = cfg["kind"]
    pb.preset = preset
    if return_scenes:
        return pb, clips
    return pb


def stress_batch(B, seed=0):
    """Config 5: 4x context rows and 4x candidate slots per clip."""
    probs = {k: 1.0 / 7 for k in range(2, 9)}
    return make_batch(B, seed=seed, preset="int_rel_ch", max_n_tripl=80, rels_n_clips=72, n_chars_probs=probs)
