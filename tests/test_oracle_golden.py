"""Pin the oracle: oracle/model.py + oracle/losses.py reproduce the golden vectors generated from the
UNMODIFIED reference (tests/golden/make_golden.py) — forward logits, the in-place -inf masking, the
loss and every parameter gradient, for all four presets and the loss variants."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from oracle import losses as ol, model as om

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("dataloader_", "pooling_", "eval_")))
DIMS = dict(text_dim=24, visual_dim=40, track_dim=40, joint_dim=16, mid_m_ints=6)
C, R = 11, 5


def test_golden_files_present():
    # 12 eval-mode cases + one train-mode (replayed dropout masks) case per preset + the relationship-only model
    # (opt.ints == 0, gates off) in eval and train mode
    assert len(GOLDEN) == 18


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference(path):
    z = np.load(path)
    preset, over = str(z["meta"][0]), dict(ast.literal_eval(str(z["meta"][1])))
    tr_correct, max_neg = bool(over.get("tr_correct")), bool(over.get("tr_max_neg"))
    sd = {k[2:]: torch.from_numpy(z[k]).clone().requires_grad_(True) for k in z.files if k.startswith("p_")}
    inp = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")}
    feats = inp["features"].float()
    kind = {"modalities": "modalities", "int_rels": "midfusion"}.get(preset, "maxtracks")
    ctx = preset in ("int_rels", "int_rel_ch")
    ints = int(over.get("ints", 1))
    cfg = om.default_cfg(ctx=int(ctx), gates=int(ctx and over.get("gates", 1)), ints=ints,
                         modality=over.get("modality", "m"), tracks=over.get("tracks", True), **DIMS)
    masks = None
    if over.get("train"):                      # the dropout masks the reference replayed (make_golden.py)
        cfg.dropout = float(z["dropout_p"])
        masks = {tuple(k[5:].split("/")): torch.from_numpy(z[k]) for k in z.files if k.startswith("mask_")}
        assert (("cat", "ints") in masks) == bool(ints)
    masked = {}
    if kind == "modalities":
        o = om.modalities_forward(sd, feats, cfg, masks)
        l = ol.max_margin_ce(o["inters"], inp["labels"], inp["multilab_weights"].float(), 0.101)
    elif kind == "midfusion":
        o = om.midfusion_forward(sd, feats, inp["rels_mask"], cfg, masks)
        l = ol.multitask_max_margin(o["inters"], o["rels"], inp["labels"], inp["rels_label"],
                                    inp["multilab_weights"].float(), 0.101, 1.0, R, ints=ints)
        assert (o["inters"] is None) == (not ints) == ("out_inters" not in z.files)
    else:
        o = om.maxtracks_forward(sd, feats, inp.get("rels_mask"), cfg, masks)
        if ctx:
            l, ts, xi, xr = ol.margin_track_rels(o["inters"], o["rels"], inp["labels"], inp["rels_label"],
                                                 inp["mem_mask"].float(), inp["multilab_weights"].float(),
                                                 inp["gt_tracks"], 0.101, 1.0, R, tr_correct=tr_correct, max_neg=max_neg)
        else:
            l, ts, xi = ol.margin_loss(o["inters"], inp["labels"], inp["mem_mask"].float(),
                                       inp["multilab_weights"].float(), inp["gt_tracks"], 0.101,
                                       tr_correct=tr_correct, max_neg=max_neg)
        masked["inters"] = xi
    l.backward()
    assert abs(l.item() - float(z["loss"])) < 2e-6 * max(1.0, abs(float(z["loss"])))
    for k in ("inters", "rels"):
        if "out_" + k not in z.files:
            continue
        ref = torch.from_numpy(z["out_" + k])
        got = masked.get(k, o[k]).detach()          # the reference stores logits AFTER its in-place masking
        assert got.shape == ref.shape
        assert torch.equal(torch.isinf(got), torch.isinf(ref)), "-inf masking differs for " + k
        fin = ~torch.isinf(ref)
        assert float((got[fin] - ref[fin]).abs().max()) < 2e-6
    for k in z.files:
        if k.startswith("g_"):
            ref = torch.from_numpy(z[k])
            got = sd[k[2:]].grad
            assert got is not None, k
            assert float((got - ref).abs().max()) <= 2e-6 + 2e-5 * float(ref.abs().max()), k
