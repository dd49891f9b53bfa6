"""numpy restatement of the prediction arg-maxes of the reference's evaluation meters.

TEST INFRASTRUCTURE — see oracle/__init__.py.  Follows Precision.update_probs_max_tracks
(utils/evaluation.py:114-175) and Precision.update_probs_max_tracks_rels (:179-271) up to the
point where predictions are compared with the ground truth; the counters themselves are not part of
the hot path.  Dense inputs: inters [B,T,C], rels [B,T,R] or None,
mask [B,T]; float32 sigmoids, float64 sums (numpy concatenates the zero 'None' column as float64).
"""
import numpy as np
from scipy.special import expit


def predict_tracks(inters, rels, mask, labels, rels_label, gt_tracks):
    """Returns int array [B, 8]: pr_track, joint_t, joint_c, joint_r, cls_gt0, cls_gt1, rel_gt0, rel_gt1."""
    x = np.array(inters, dtype=np.float32, copy=True)
    B, T, C = x.shape
    live = np.asarray(mask) != 0
    x[~live] = -np.inf                                               # :124 / :193
    p_cl = expit(x)                                                  # float32
    b = np.arange(B)
    out = -np.ones((B, 8), dtype=np.int64)
    if rels is None:
        score = p_cl[b, :, labels].astype(np.float64)
        out[:, 0] = np.argmax(score, axis=1)                         # :137
        flat = np.argmax(p_cl.reshape(B, -1).astype(np.float64), axis=1)   # :144-147
        out[:, 1], out[:, 2] = flat // C, flat % C
    else:
        r = np.array(rels, dtype=np.float32, copy=True)
        R = r.shape[2]
        r[~live] = -np.inf                                           # :194
        p_r = np.concatenate((expit(r), np.zeros((B, T, 1))), axis=2)       # :219-220, float64
        gt_rel = np.asarray(rels_label)[:, 0]                        # :208
        out[:, 0] = np.argmax(p_cl[b, :, labels] + p_r[b, :, gt_rel], axis=1)   # :221-222
        joint = p_cl[:, :, :, None].astype(np.float64) + p_r[:, :, None, :]     # :229-231
        flat = np.argmax(joint.reshape(B, -1), axis=1)
        n = C * (R + 1)
        out[:, 1], out[:, 2], out[:, 3] = flat // n, (flat % n) // (R + 1), (flat % n) % (R + 1)
    for i in range(2):
        g = np.asarray(gt_tracks)[:, i]
        ok = live[b, g]
        out[ok, 4 + i] = np.argmax(x[b, g, :], axis=1)[ok]           # :152 / :241
        if rels is not None:
            out[ok, 6 + i] = np.argmax(r[b, g, :], axis=1)[ok]       # :243
    return out
