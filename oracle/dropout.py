"""numpy mirror of the counter-based dropout mask of liblirec_b200 (csrc/common.cuh:
fmix32 / drop_row_key / drop_keep).  The reference uses torch's global RNG for nn.Dropout
(mlp/model.py:52); masks cannot be matched to it, so train-mode parity is checked by feeding
THESE masks to the oracle (oracle/model.py `masks=`)."""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def _u32(x):
    return (np.asarray(x, dtype=np.uint64)) & _M32


def fmix32(h):
    h = _u32(h)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & _M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & _M32
    h ^= h >> np.uint64(16)
    return h


def row_key(seed, stream_id, rows):
    rows = _u32(rows)
    inner = fmix32((rows + np.uint64(0x7F4A7C15)) & _M32)
    k = np.uint64(seed & 0xFFFFFFFF) ^ ((np.uint64(stream_id) * np.uint64(0x9E3779B1)) & _M32) ^ inner
    return fmix32(k)


def keep_mask(seed, stream_id, rows, cols, p):
    """Boolean [len(rows), len(cols)] mask: True where element (row, col) is kept.  One 32-bit hash
    word covers two adjacent columns (16 bits each); an element is dropped iff its lane < p * 65536."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    if p <= 0:
        return np.ones((rows.size, cols.size), dtype=bool)
    rk = row_key(seed, stream_id, rows)[:, None]
    pair = _u32(cols >> 1)
    ck = ((pair * np.uint64(0x9E3779B1)) + np.uint64(0x632BE5AB)) & _M32
    word = fmix32(rk ^ ck[None, :])
    lane = (word >> (np.uint64(16) * _u32(cols & 1))[None, :]) & np.uint64(0xFFFF)
    thr = np.uint64(int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)))
    return lane >= thr


# dropout sites of lirec_model_forward (csrc/model.cu: DS_*)
DS_L1_INTS, DS_L1_CTX, DS_CAT_INTS, DS_CAT_CTX, DS_GATE = 1, 2, 3, 4, 5


def dense_masks(pb, seed, p, J=512, gate_dim=3072, kind="maxtracks", cat_width=None):
    """Masks of one step, laid out for the DENSE oracle (oracle/model.py `masks=`) from the packed
    tables of host PackedBatch `pb`: candidate r sits at dense row (clip, slot), context row x at
    (clip, slot, position).  Empty slots get all-ones masks (they are masked out downstream)."""
    import torch
    t = pb.tables
    B, T, S = pb.B, pb.n_slots, pb.n_ctx_slots
    Ni = pb.n_cand
    dense_row = t["cand_clip"].astype(np.int64) * T + t["cand_slot"].astype(np.int64)
    slots = ("txt", "vis", "tracks1", "tracks2")
    masks = {}
    for s, name in enumerate(slots):
        m = np.ones((B * T, J), dtype=bool)
        m[dense_row] = keep_mask(seed, DS_L1_INTS, np.arange(Ni), s * J + np.arange(J), p)
        masks[("l1", "ints", name)] = torch.from_numpy(m)
    cw = 3 * J if cat_width is None else int(cat_width)      # Modalities without some slots: narrower concat
    m = np.ones((B * T, cw), dtype=bool)
    m[dense_row] = keep_mask(seed, DS_CAT_INTS, np.arange(Ni), np.arange(cw), p)
    masks[("cat", "ints")] = torch.from_numpy(m)
    if pb.has_ctx:
        Nx = pb.n_ctx_rows
        owner = t["ctx_owner"].astype(np.int64)
        pos = np.arange(Nx) - t["ctx_off"][:-1].astype(np.int64)[owner]
        for s, name in enumerate(slots):
            m = np.ones((B * T, S, J), dtype=bool)
            if Nx:
                m[dense_row[owner], pos] = keep_mask(seed, DS_L1_CTX, np.arange(Nx), s * J + np.arange(J), p)
            masks[("l1", "ctx", name)] = torch.from_numpy(m)
        m = np.ones((B * T, 3 * J), dtype=bool)
        m[dense_row] = keep_mask(seed, DS_CAT_CTX, np.arange(Ni), np.arange(3 * J), p)
        masks[("cat", "ctx")] = torch.from_numpy(m)
        m = np.ones((B * T, gate_dim), dtype=bool)
        m[dense_row] = keep_mask(seed, DS_GATE, np.arange(Ni), np.arange(gate_dim), p)
        masks[("gate",)] = torch.from_numpy(m)
    return masks


def cat_distr_uniform(seed, n_clips):
    """The uniform draw in [0, 1) the track-loss kernel uses for clip b under opt.tr_cat_distr
    (csrc/loss.cu: fmix32(seed ^ fmix32(b + 0x51ED270B)) >> 8, scaled by 2^-24)."""
    b = np.arange(n_clips, dtype=np.uint64)
    inner = fmix32((b + np.uint64(0x51ED270B)) & _M32)
    h = fmix32((np.uint64(int(seed) & 0xFFFFFFFF) ^ inner) & _M32)
    return (h >> np.uint64(8)).astype(np.float64) / 16777216.0
