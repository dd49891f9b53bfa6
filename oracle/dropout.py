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
    """Boolean [len(rows), len(cols)] mask: True where element (row, col) is kept."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    if p <= 0:
        return np.ones((rows.size, cols.size), dtype=bool)
    rk = row_key(seed, stream_id, rows)[:, None]
    ck = ((_u32(cols) * np.uint64(0x9E3779B1)) + np.uint64(0x632BE5AB)) & _M32
    h = fmix32(rk ^ ck[None, :])
    u = (h >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return u >= np.float32(p)
