"""Pooled features of one (movie, scene): temporal MAX pooling of variable-length visual / text /
track sequences (reference: mixed_utils/mixed_features.py:37-112) on the GPU.

The reference pools one sequence at a time with np.max on the CPU and caches .npy files.  Here all
sequences of a scene (or of a whole split) are pooled by ONE segmented-max launch over offset tables
(lirec_seg_reduce_f32), writing bf16 rows straight into the clip / track banks the model reads.
Empty sequences give zero rows, like the reference (text_features.py:171-178, mixed_features.py:89-93).
"""
import numpy as np
import torch

from lirec_b200 import ops
from lirec_b200.packing import CLIP_DIM, TEXT_DIM, TRACK_DIM, VISUAL_DIM


def _offsets(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.int32)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    return off


def pool_sequences(seqs, dim, device="cuda", out=None, pinned=True, mode="max", beta=1.0):
    """seqs: list of float arrays [len_i, dim] (len_i may be 0) -> bf16 [len(seqs), dim] on device.
    mode: 'max' — the reference's pooling (np.max, mixed_features.py:54, 61, 105) and the default; 'mean'; or
    'softmax' — softmax_r(beta * x[r, c])-weighted sum per channel (lirec_seg_softmax_pool_fwd; the reference has
    no such pooling: parity unpinned, its beta -> inf / beta = 0 limits are the other two modes)."""
    off = _offsets(seqs)
    total = int(off[-1])
    flat = np.concatenate([np.asarray(s, dtype=np.float32).reshape(-1, dim) for s in seqs]) if total else \
        np.zeros((0, dim), dtype=np.float32)
    host = torch.from_numpy(flat)
    if pinned and total:
        host = host.pin_memory()
    x = host.to(device, non_blocking=True)
    offd = torch.from_numpy(off).to(device, non_blocking=True)
    if out is None:
        out = torch.empty(len(seqs), dim, dtype=torch.bfloat16, device=device)
    if total == 0:
        out.zero_()
        return out
    if mode == "softmax":
        pooled, _ = ops.seg_softmax_pool(x, offd, beta=beta, need_lse=False)
        out.copy_(pooled)
    else:
        ops.seg_reduce(x, offd, mode, out_bf16=out)
    return out


class MixedFeatures:
    """Feature holder of one scene.  `visual` [T, 2048] (already spatially mean-pooled, reference
    visual_features.py:67-69), `text` [n_tokens, 768], `tracks` {name: [len, 2048]}."""

    def __init__(self, visual, text, tracks, device="cuda"):
        self.visual, self.text, self.tracks, self.device = visual, text, tracks, device
        self.cached, self.cached_tracks = {}, {}

    def get_features_by_time(self, frame_range=None, token_range=None, idx=None):
        """max over the clip's frames ‖ max over its tokens -> bf16 [1, 2816] (text first)."""
        if idx in self.cached:
            return self.cached[idx]
        f0, f1 = frame_range if frame_range else (0, len(self.visual))
        t0, t1 = token_range if token_range else (0, len(self.text))
        out = torch.empty(1, CLIP_DIM, dtype=torch.bfloat16, device=self.device)
        pool_sequences([self.text[t0:t1]], TEXT_DIM, self.device, out=out[:, :TEXT_DIM])
        pool_sequences([self.visual[f0:f1]], VISUAL_DIM, self.device, out=out[:, TEXT_DIM:])
        if idx is not None:
            self.cached[idx] = out
        return out

    def get_features_by_track(self, name=None, idx=None):
        if idx in self.cached_tracks:
            return self.cached_tracks[idx]
        seq = self.tracks.get(name, np.zeros((0, TRACK_DIM), dtype=np.float32))
        out = pool_sequences([seq], TRACK_DIM, self.device)
        if idx is not None:
            self.cached_tracks[idx] = out
        return out
