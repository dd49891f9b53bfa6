"""numpy restatement of the reference's temporal / spatial pooling (TEST INFRASTRUCTURE).

  segmented_max()   np.max over a variable-length sequence, zeros for an empty one:
                    mixed_utils/mixed_features.py:54, 61, 105; text_utils/text_features.py:171-178;
                    mixed_utils/mixed_features.py:89-93
  segmented_mean()  plain mean over a segment (spatial / ROI mean, visual_utils/visual_features.py:67-69,
                    133-134, applied to already flattened positions)
"""
import numpy as np


def segmented_max(x, seg_off):
    out = np.zeros((len(seg_off) - 1, x.shape[1]), dtype=x.dtype)
    for s in range(len(seg_off) - 1):
        a, b = int(seg_off[s]), int(seg_off[s + 1])
        if b > a:
            out[s] = np.max(x[a:b], axis=0)
    return out


def segmented_mean(x, seg_off):
    out = np.zeros((len(seg_off) - 1, x.shape[1]), dtype=np.float64)
    for s in range(len(seg_off) - 1):
        a, b = int(seg_off[s]), int(seg_off[s + 1])
        if b > a:
            out[s] = np.mean(x[a:b].astype(np.float64), axis=0)
    return out.astype(np.float32)
