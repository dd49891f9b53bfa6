"""Evaluation meters: the prediction arg-maxes of oracle/evaluation.py fed
through lirec_b200/utils/evaluation.py must reproduce the counters of the reference's UNMODIFIED
Precision / RelationshipsAcc meters (tests/golden/eval_meters.npz, made by tests/golden/make_eval_golden.py).
This pins both the arg-max restatement and the counter bookkeeping; tests/test_loss_gpu.py then checks the
device arg-max kernel against the same restatement bit-exactly."""
import os

import numpy as np
import torch

from lirec_b200.utils import evaluation as ev
from oracle import evaluation as oe

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_meters.npz"))


def _batches():
    for i in range(int(G["n_batches"])):
        yield {k[len("b%d_" % i):]: G[k] for k in G.files if k.startswith("b%d_" % i)}


def replay(predict, device="cpu"):
    R = int(G["R"])
    tr, trr, top = ev.TrackMeters(device), ev.TrackMeters(device), ev.TopKMeters(device)
    racc = ev.RelationshipsAcc(R, 16, device)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(device)
    for b in _batches():
        B = len(b["labels"])
        ar = np.arange(B)
        gt = b["gt"]
        pred = predict(b, with_rels=False)
        tr.update(t(pred), t(b["labels"]), t(gt), t(b["just_zeros"]))
        pred = predict(b, with_rels=True)
        rel_at_gt = np.stack((b["rels_label"][ar, gt[:, 0]], b["rels_label"][ar, gt[:, 1]]), axis=1)
        trr.update(t(pred), t(b["labels"]), t(gt), t(b["just_zeros"]), gt_rel=t(b["rels_label"][:, 0]),
                   rel_at_gt=t(rel_at_gt), n_rels=R)
        top.update(t(b["ints"][:, 0]), t(b["labels"]))
        sel = np.nonzero(b["rels_label"][:, 0] != R)[0]
        if len(sel):
            racc.update(t(b["rels"][sel, 0]), t(b["rels_label"][sel, 0]), t(b["hash_rel"][sel]))
    return tr, trr, top, racc


def check(tr, trr, top, racc):
    names = ev.TrackMeters.NAMES
    got = tr.counts()
    ref = dict(zip(names, G["ref_tr"].tolist()))
    assert got == ref, (got, ref)
    got, ref = trr.counts(), dict(zip(names, G["ref_trr"].tolist()))
    assert got == ref, (got, ref)
    assert list(top.counts().values()) == G["ref_top"].tolist()
    r = racc.compute()
    assert [r["total"], r["top1"], r["top3"]] == G["ref_racc"].tolist()


def _oracle_predict(b, with_rels):
    return oe.predict_tracks(b["ints"], b["rels"] if with_rels else None, b["mask"], b["labels"],
                             b["rels_label"] if with_rels else None, b["gt"])


def test_meters_reproduce_the_reference_counters():
    check(*replay(_oracle_predict))
