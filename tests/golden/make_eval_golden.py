"""Generate tests/golden/eval_meters.npz from the UNMODIFIED reference meters (build container only):
`Precision.update_probs_max_tracks`, `update_probs_max_tracks_rels`, `update_probs` and
`RelationshipsAcc` (utils/evaluation.py) over three batches of random logits with ragged candidate
masks, bidirectional ground truths, all-zero-track clips and None relationships, called the way
mlp/test.py:47-87 calls them.  Inputs and the resulting counters are stored;
tests/test_eval_cpu.py replays them through oracle/evaluation.py + lirec_b200/utils/evaluation.py."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.argv = sys.argv[:1]

from oracle import reference_shim as rs  # noqa: E402

C, R, T = 13, 6, 8


def batch(rng, B):
    counts = rng.choice([1, 2, 3, 6, 8], size=B)
    mask = (np.arange(T)[None, :] < counts[:, None])
    ints = (rng.standard_normal((B, T, C)) * 2).astype(np.float32)
    rels = (rng.standard_normal((B, T, R)) * 2).astype(np.float32)
    labels = rng.integers(C, size=B)
    rels_label = rng.integers(R + 1, size=(B, T)) * mask          # pad label 0
    gt = np.zeros((B, 2), dtype=np.int64)
    for b in range(B):
        if counts[b] > 1 and rng.random() < 0.5:
            gt[b, 1] = rng.integers(1, counts[b])
    just_zeros = rng.random(B) < 0.2
    # make some clips easy so every counter moves
    for b in range(0, B, 3):
        ints[b, gt[b, 0], labels[b]] += 6.0
        if rels_label[b, 0] != R:
            rels[b, gt[b, 0], rels_label[b, 0]] += 6.0
    hash_rel = rng.integers(0, 9, size=B)
    return dict(ints=ints, rels=rels, mask=mask.astype(np.float64), labels=labels, rels_label=rels_label, gt=gt,
                just_zeros=just_zeros, hash_rel=hash_rel)


def main():
    opt, _ = rs.load_dataloader()
    ev = rs._state["evaluation"]
    rng = np.random.default_rng(0)
    batches = [batch(rng, B) for B in (16, 9, 12)]
    opt.soft_gt = False
    out = {"n_batches": len(batches), "C": C, "R": R, "T": T}
    p_tr, p_trr, p_top = (ev.Precision(inter2mgd=None, n_rels=0) for _ in range(3))
    racc = ev.RelationshipsAcc(n_rels=R + 1)
    for i, b in enumerate(batches):
        for k, v in b.items():
            out["b%d_%s" % (i, k)] = v
        t = lambda a: torch.from_numpy(np.array(a, copy=True))
        p_tr.update_probs_max_tracks(t(b["ints"]), gt_tracks=t(b["gt"]), gt_classes=t(b["labels"]), n_names=None,
                                     mask=t(b["mask"]), just_zeros=t(b["just_zeros"]))
        rels_mask = torch.nonzero(t(b["rels_label"])[:, 0] - (R + 1) + 1)        # mlp/test.py:63, n_rels = R + 1
        p_trr.update_probs_max_tracks_rels(t(b["ints"]), t(b["rels"]), t(b["labels"]), t(b["rels_label"]),
                                           gt_tracks=t(b["gt"]), just_zeros=t(b["just_zeros"]), mask=t(b["mask"]),
                                           rels_mask=rels_mask)
        p_top.update_probs(t(b["ints"][:, 0]), t(b["labels"]), conf_mat=np.zeros((C, C)))
        sel = np.nonzero(b["rels_label"][:, 0] != R)[0]
        if len(sel):
            racc.update(t(b["rels"][sel, 0]), t(b["rels_label"][sel, 0]), t(b["hash_rel"][sel]))
    for name, p in (("tr", p_tr), ("trr", p_trr)):
        out["ref_" + name] = np.array([p.total, p.total_cl, getattr(p, "total_rels", 0), p._top1, p._cls_top1,
                                       p._trks_top1, p._rels_top1], dtype=np.int64)
    out["ref_top"] = np.array([p_top.total, p_top._top1, p_top._top3, p_top._top5], dtype=np.int64)
    racc.top1()
    out["ref_racc"] = np.array([racc.total, racc._top1, racc._top3], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "eval_meters.npz"), **out)
    print({k: out[k].tolist() for k in out if k.startswith("ref_")})


if __name__ == "__main__":
    main()
