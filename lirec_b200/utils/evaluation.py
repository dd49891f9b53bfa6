"""Accuracy meters of the evaluation loop (reference: utils/evaluation.py) over DEVICE-side predictions.

The reference moves the logits to the host every batch and runs `Precision.update_probs_max_tracks`
(:114-175), `update_probs_max_tracks_rels` (:179-271), `update_probs` (:68-107) and
`RelationshipsAcc.update` (:383-397) in numpy.  Here the arg-maxes come from `lirec_predict_tracks`
(int32 [B, 8] per batch, bit-exact with the numpy formulas) and the reference's counter bookkeeping —
which clips count (`just_zeros` clips are left out of the track / joint totals), the second chance a
bidirectional interaction gets through `gt_tracks[:, 1]`, the relationship counter that only runs
over clips whose ground-truth pair has a relationship — is replayed on integer tensors on the device;
one small tensor crosses to the host per evaluation.  Pinned against the reference's unmodified meters
(tests/golden/eval_meters.npz, tests/test_eval_cpu.py).
"""
import torch


class TrackMeters:
    """Counters of Precision.update_probs_max_tracks(_rels).  Accessors are the reference's:
    top1 = _top1 / total, trks_top1 = _trks_top1 / total, cls_top1 = _cls_top1 / total_cl and
    rels_top1 = _rels_top1 / total (the reference defines rels_top1 twice, :353 and :359; the second,
    dividing by `total`, is the one in effect)."""
    NAMES = ("total", "total_cl", "total_rels", "top1", "cls_top1", "trks_top1", "rels_top1")

    def __init__(self, device="cpu"):
        self.c = torch.zeros(len(self.NAMES), dtype=torch.long, device=device)

    def update(self, pred, labels, gt_tracks, just_zeros, gt_rel=None, rel_at_gt=None, n_rels=0):
        """pred [B, 8] from lirec_predict_tracks; labels [B]; gt_tracks [B, 2]; just_zeros [B] bool;
        gt_rel [B] relationship label of slot 0 (rels_label[:, 0]); rel_at_gt [B, 2] relationship labels
        at slots gt_tracks[:, 0] / [:, 1]; n_rels = R (index of None)."""
        pred = pred.long()
        y, gt = labels.long(), gt_tracks.long()
        nz = ~just_zeros.bool()
        bi = gt[:, 1] != 0
        pr_track, jt, jc, jr = pred[:, 0], pred[:, 1], pred[:, 2], pred[:, 3]
        # given the class (and relationship): which track pair?  second chance for bidirectional clips
        miss0 = pr_track != gt[:, 0]
        trk = (~miss0) | (bi & (pr_track == gt[:, 1]))
        # given the tracks: which class?  (slot gt1 is tried when slot gt0 was wrong, :165, :255)
        cls = (pred[:, 4] == y) | (pred[:, 5] == y)
        # nothing given: joint arg-max over (track, class[, relationship])
        j0 = (jc == y) & (jt == gt[:, 0])
        j1 = (jc == y) & (jt == gt[:, 1])
        with_rels = gt_rel is not None
        if with_rels:
            gr = gt_rel.long()
            j0, j1 = j0 & (jr == gr), j1 & (jr == gr)
        joint = j0 | (bi & miss0 & ~j0 & j1)
        rel_n = rel_ok = torch.zeros((), dtype=torch.long, device=pred.device)
        if with_rels:
            has = gr != n_rels                                            # mlp/test.py:63
            ra = rel_at_gt.long()
            rel = (pred[:, 6] == ra[:, 0]) | (pred[:, 7] == ra[:, 1])
            rel_n, rel_ok = has.sum(), (rel & has).sum()
        self.c += torch.stack([nz.sum(), torch.tensor(len(y), device=pred.device), rel_n, (joint & nz).sum(), cls.sum(),
                               (trk & nz).sum(), rel_ok])
        return self

    def counts(self):
        return dict(zip(self.NAMES, self.c.tolist()))

    @staticmethod
    def ratios(c):
        d = lambda a, b: float(a) / float(b) if b else 0.0
        return {"top1": d(c["top1"], c["total"]), "trks_top1": d(c["trks_top1"], c["total"]),
                "cls_top1": d(c["cls_top1"], c["total_cl"]), "rels_top1": d(c["rels_top1"], c["total"])}


class TopKMeters:
    """Precision.update_probs (:68-107): top-1 / top-3 / top-5 of row-wise scores."""

    def __init__(self, device="cpu"):
        self.c = torch.zeros(4, dtype=torch.long, device=device)      # total, top1, top3, top5

    def update(self, scores, gt):
        order = torch.argsort(-scores.float(), dim=1)[:, :5]
        hit = order == gt.long().view(-1, 1)
        self.c += torch.stack([torch.tensor(len(gt), device=scores.device), hit[:, :1].any(1).sum(),
                               hit[:, :3].any(1).sum(), hit[:, :5].any(1).sum()])
        return self

    def counts(self):
        return dict(zip(("total", "top1", "top3", "top5"), self.c.tolist()))


class RelationshipsAcc:
    """RelationshipsAcc (:367-417): sigmoid scores of all clips of the same (movie, pair, relationship)
    — `hash_rel` — are summed, then ranked once per pair."""

    def __init__(self, n_rels, n_hash, device="cpu"):
        self.sum = torch.zeros(n_hash, n_rels, dtype=torch.float32, device=device)
        self.gt = torch.full((n_hash,), -1, dtype=torch.long, device=device)
        self.order = torch.full((n_hash,), -1, dtype=torch.long, device=device)   # first-seen order, like the dict
        self.seen = 0

    def update(self, logits, gt, hashes):
        h = hashes.long()
        assert int((h < 0).sum()) == 0
        self.sum.index_add_(0, h, torch.sigmoid(logits.float()))
        first = self.gt[h] < 0
        # the first clip of a pair fixes its label (:392-396); duplicates inside the batch keep the earliest
        idx = torch.arange(len(h), device=h.device)
        earliest = torch.full_like(self.gt, len(h)).scatter_reduce_(0, h, idx, "amin")
        pick = first & (earliest[h] == idx)
        self.gt[h[pick]] = gt.long()[pick]
        return self

    def compute(self):
        live = self.gt >= 0
        order = torch.argsort(-self.sum[live], dim=1)
        g = self.gt[live].view(-1, 1)
        total = int(live.sum())
        top1 = int((order[:, :1] == g).any(1).sum())
        top3 = int((order[:, :3] == g).any(1).sum())
        return {"total": total, "top1": top1, "top3": top3}
