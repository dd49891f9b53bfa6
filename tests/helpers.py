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


def oracle_forward_loss(pb, sd, preset, opt, masks, tape=None, relu_gate=None):
    """Dense fp64 oracle forward + loss for host PackedBatch `pb`. Returns (outputs, loss, extra)."""
    from lirec_b200.mixed_utils import synthetic
    from oracle import losses as ol, model as om
    kind = synthetic.PRESETS[preset]["kind"]
    dense = pb.to_dense(np.float64)
    cfg = om.default_cfg(ctx=int(opt.ctx), gates=int(opt.gates), dropout=opt.dropout,
                         ints=1 if kind == "modalities" else int(opt.ints))
    if kind == "modalities":
        cfg.modality, cfg.tracks = opt.modality, bool(opt.tracks)
    if tape is not None:
        cfg.tape = tape
    if relu_gate is not None:
        cfg.relu_gate = relu_gate
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
                                    opt.lymbda, N_RELS, ints=int(opt.ints))
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


def kernel_relu_gates(model, pb_host, pb_dev, seed, train):
    """White-box: the on/off decision the CUDA path takes for every ReLU unit of one step, laid out like the
    dense oracle's pre-activations (oracle/model.py `cfg.relu_gate`).  Runs lirec_model_forward once and reads
    the workspace: r1 (relu(L1) of the unique bank rows, fp32 — backward gates on r1 > 0, csrc/rows.cu) and
    the gate output g2 (hi bf16 — backward gates on g2 > 0, csrc/gemm_tcgen05.cu POST_DRELU).  The oracle uses
    these ONLY for units whose fp64 pre-activation is within 1e-5 of zero (knife-edges)."""
    import ctypes as C
    from lirec_b200 import _ext
    slots = ("txt", "vis", "tracks1", "tracks2")
    training = bool(train) and model.dropout.p > 0
    with torch.no_grad():
        model._sync_flat()
        model._refresh_bf16()
        batch_c, ws, _, _ = model._run_forward(pb_dev, training, seed if training else 0)
    torch.cuda.synchronize()
    offs = (C.c_int64 * 40)()
    n = _ext.lib().lirec_model_workspace_layout(C.byref(model._cfg_c), C.byref(batch_c), offs, 40)
    assert n >= 31
    off = list(offs)[:n]
    t = pb_host.tables
    T, S, J = pb_host.n_slots, pb_host.n_ctx_slots, 512
    B, Ni = pb_host.B, pb_host.n_cand
    dense_row = torch.from_numpy(t["cand_clip"].astype(np.int64) * T + t["cand_slot"].astype(np.int64))
    col = (0, 0, 1, 2)
    gates = {}

    def r1(br, s, nu):
        o = off[br * 4 + s]
        if o < 0:
            return None
        return ws[o:o + nu * J * 4].view(torch.float32).view(nu, J).cpu() > 0

    cand_rows = torch.from_numpy(np.asarray(t["cand_rows"]).astype(np.int64))
    for s, name in enumerate(slots):
        nu = batch_c.n_clip_ints if s < 2 else batch_c.n_track_ints
        g = r1(0, s, nu)
        if g is None:
            continue
        d = torch.ones(B * T, J, dtype=torch.bool)
        d[dense_row] = g[cand_rows[:, col[s]]]
        gates["z1_%s_ints" % name] = d
    if model._ctx:
        Nx = pb_host.n_ctx_rows
        owner = torch.from_numpy(np.asarray(t["ctx_owner"]).astype(np.int64))
        pos = torch.arange(Nx) - torch.from_numpy(np.asarray(t["ctx_off"]).astype(np.int64))[:-1][owner]
        ctx_rows = torch.from_numpy(np.asarray(t["ctx_rows"]).astype(np.int64))
        for s, name in enumerate(slots):
            nu = batch_c.n_clip if s < 2 else batch_c.n_track
            g = r1(1, s, nu)
            d = torch.ones(B * T, S, J, dtype=torch.bool)
            if Nx:
                d[dense_row[owner], pos] = g[ctx_rows[:, col[s]]]
            gates["z1_%s_ctx" % name] = d
    if model._gates:
        Gd = model._gate_dim
        o = off[27]                                          # g2: [Ni, 2 * Gd] bf16, hi | lo
        g2 = ws[o:o + Ni * 2 * Gd * 2].view(torch.bfloat16).view(Ni, 2 * Gd)[:, :Gd].float().cpu() > 0
        d = torch.ones(B * T, Gd, dtype=torch.bool)
        d[dense_row] = g2
        gates["pre_gate"] = d
    return gates
