"""Shared helpers of the parity tests."""
import contextlib
import io

import numpy as np
import torch

N_CLASSES, N_RELS = 101, 15
TOL = 1e-3   # north_star: logits, losses and gradients within 1e-3 relative (max-norm per tensor)


def rel_err(a, b):
    """Max-norm relative error ||a-b||_inf / ||b||_inf (elementwise relative error is undefined at
    zero crossings)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rounded_state_dict(model):
    """The model's parameters as the kernels see them: bf16-rounded weights, fp32 biases, as fp64
    leaves for the oracle (identical operand rounding on both sides)."""
    sd = {}
    for k, v in model.state_dict().items():
        v = v.detach().cpu()
        sd[k] = (v.to(torch.bfloat16) if k.endswith("weight") else v).double().requires_grad_(True)
    return sd


def make_model(n_classes=N_CLASSES, n_rels=N_RELS, seed=0):
    import lirec_b200.mlp.model as M
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return M.create_model(n_classes, n_rels=n_rels)


def oracle_forward_loss(pb, sd, preset, opt, masks, tape=None):
    """Dense fp64 oracle forward + loss for host PackedBatch `pb`. Returns (outputs, loss, extra)."""
    from lirec_b200.mixed_utils import synthetic
    from oracle import losses as ol, model as om
    kind = synthetic.PRESETS[preset]["kind"]
    dense = pb.to_dense(np.float64)
    cfg = om.default_cfg(ctx=int(opt.ctx), gates=int(opt.gates), dropout=opt.dropout)
    if kind == "modalities":
        cfg.modality, cfg.tracks = opt.modality, bool(opt.tracks)
    if tape is not None:
        cfg.tape = tape
    B = pb.B
    f = dense["features"]
    extra = {"dense": dense}
    if kind == "modalities":
        o = om.modalities_forward(sd, f.reshape(B, 1, -1), cfg, masks)
        l = ol.max_margin_ce(o["inters"], dense["labels"], dense["multilab_weights"], opt.margin)
        ragged = {"inters": o["inters"]}
    elif kind == "midfusion":
        o = om.midfusion_forward(sd, f.reshape(B, -1, f.shape[-1]), dense["rels_mask"].reshape(B, -1, 1), cfg, masks)
        l = ol.multitask_max_margin(o["inters"], o["rels"], dense["labels"].reshape(B, 1, 1),
                                    dense["rels_label"].reshape(B), dense["multilab_weights"], opt.margin,
                                    opt.lymbda, N_RELS)
        ragged = {"inters": o["inters"], "rels": o["rels"]}
    else:
        o = om.maxtracks_forward(sd, f, dense.get("rels_mask"), cfg, masks)
        mm = dense["mem_mask"].bool()
        max_neg = bool(opt.tr_max_neg and opt.tr_sum_max_flag)
        if opt.ctx:
            l, ts, xi, xr = ol.margin_track_rels(o["inters"], o["rels"], dense["labels"], dense["rels_label"],
                                                 dense["mem_mask"], dense["multilab_weights"], dense["gt_tracks"],
                                                 opt.tr_margin, opt.lymbda, N_RELS, tr_correct=opt.tr_correct,
                                                 max_neg=max_neg)
            ragged = {"inters": o["inters"][mm], "rels": o["rels"][mm]}
        else:
            l, ts, xi = ol.margin_loss(o["inters"], dense["labels"], dense["mem_mask"], dense["multilab_weights"],
                                       dense["gt_tracks"], opt.tr_margin, tr_correct=opt.tr_correct, max_neg=max_neg)
            ragged = {"inters": o["inters"][mm]}
        extra["assignment"] = ts
    return ragged, l, extra
