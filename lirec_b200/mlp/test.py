"""Evaluation loop (reference: mlp/test.py:17-145): model.eval(), no_grad, loss, accuracy counters,
returns {'total', 'ints', 'rels', 'tracks', 'joint'}.

The reference moves the logits to the host every batch and arg-maxes them in numpy / scipy
(utils/evaluation.py).  Here the prediction arg-maxes run on the GPU over the ragged logits
(lirec_predict_tracks, bit-exact against the numpy formulas — tests/test_loss_gpu.py) and only
integer counters come back, once per evaluation.  Counters follow the reference's definitions for
the first ground-truth slot and accept the second one for bidirectional interactions
(evaluation.py:150-175, 237-271); its per-movie bookkeeping and HTML dumps are not reproduced."""
import torch

from lirec_b200 import dp, ops
from lirec_b200.mixed_utils.classification_dataloader import packed_loader
from lirec_b200.utils.arg_pars import opt
from lirec_b200.utils.util_functions import Averaging


def _track_counters(out, pb, n_rels):
    """[n, trk_ok, cls_ok, n_rel, rel_ok, joint_ok] for one batch (device int64)."""
    pred = ops.predict_tracks(out.ragged_inters, out.ragged_rels, pb["cand_off"], pb["labels"],
                              pb["rels_label"] if n_rels else None, pb["gt_tracks"], n_rels).long()
    y, gt = pb["labels"].long(), pb["gt_tracks"].long()
    bi = gt[:, 1] != 0
    trk = (pred[:, 0] == gt[:, 0]) | (bi & (pred[:, 0] == gt[:, 1]))
    cls = (pred[:, 4] == y) | (bi & (pred[:, 5] == y))
    joint = (pred[:, 2] == y) & ((pred[:, 1] == gt[:, 0]) | (bi & (pred[:, 1] == gt[:, 1])))
    zero = torch.zeros((), dtype=torch.long, device=pred.device)
    n_rel, rel_ok = zero, zero
    if n_rels:
        first = pb["cand_off"][:-1].long()
        r0 = pb["rels_label"].long()[first]                       # relationship label of the GT slot
        has = r0 != n_rels
        r1 = pb["rels_label"].long()[first + gt[:, 1]]
        rel = (pred[:, 6] == r0) | (bi & (pred[:, 7] == r1))
        n_rel, rel_ok = has.sum(), (rel & has).sum()
        joint = joint & ((pred[:, 3] == r0) | ~has)
    return torch.stack([torch.tensor(len(y), device=pred.device), trk.sum(), cls.sum(), n_rel, rel_ok, joint.sum()])


def _topk_counters(logits, labels, ks=(1, 5)):
    top = logits.topk(max(ks), dim=1).indices
    hit = top == labels.view(-1, 1)
    return [hit[:, :k].any(1).sum() for k in ks]


def testing(test_dataset, model, loss, total_iter=1, mode="val", train_start_time=""):
    rank, world = (torch.distributed.get_rank(), dp.world_size()) if dp.world_size() > 1 else (0, 1)
    losses = Averaging()
    model.eval()
    n_rels = test_dataset.n_rels - 1 if opt.rels_multitask else 0
    dev = next(model.parameters()).device
    acc = torch.zeros(6, dtype=torch.long, device=dev)
    top = torch.zeros(3, dtype=torch.long, device=dev)            # n, top1, top5
    rel = torch.zeros(3, dtype=torch.long, device=dev)            # n, top1, top3
    loss_sum = torch.zeros((), device=dev)
    n_batches = 0
    with torch.no_grad():
        for pb in packed_loader(test_dataset, opt.batch_size, shuffle=False, num_workers=opt.num_workers,
                                device=dev, rank=rank, world=world):
            if getattr(pb.host, "global_clips", pb.B) == 1:     # reference skips batches of one (:38-39)
                continue
            out = model(pb)
            loss_sum += loss(out, {}) * pb.B
            n_batches += 1
            if opt.tr_maximize:
                acc += _track_counters(out, pb, n_rels if opt.ctx == 1 else 0)
            else:
                t1, t5 = _topk_counters(out.ragged_inters, pb["labels"].long())
                top += torch.stack([torch.tensor(pb.B, device=dev), t1, t5])
                if opt.rels_multitask and opt.ctx == 1:
                    lab = pb["rels_label"].long()
                    sel = lab != n_rels
                    if bool(sel.any()):
                        r1, r3 = _topk_counters(out.ragged_rels[sel], lab[sel], ks=(1, 3))
                        rel += torch.stack([sel.sum(), r1, r3])
    stats = torch.cat([acc, top, rel, loss_sum.view(1).long() * 0]).double()
    stats[-1] = loss_sum.double()
    if world > 1:
        torch.distributed.all_reduce(stats)                     # a few integers per evaluation
    s = stats.tolist()                                          # the only device->host read
    acc, top, rel, loss_total = s[0:6], s[6:9], s[9:12], s[12]
    n_clips = acc[0] if opt.tr_maximize else top[0]
    losses.update(loss_total / max(n_clips, 1), int(max(n_clips, 1)))

    def ratio(a, b):
        return float(a) / float(b) if b else 0.0
    out_val = out_val_ints = out_val_rels = out_val_tr = out_val_joint = 0.0
    print("%s loss: %f" % (mode.upper(), losses.avg))
    if opt.tr_maximize:
        out_val_tr, out_val_ints = ratio(acc[1], acc[0]), ratio(acc[2], acc[0])
        out_val_joint = ratio(acc[5], acc[0])
        print("%s pr@1: %f" % (mode.upper(), out_val_joint))
        print("%s pr@trks: %f" % (mode.upper(), out_val_tr))
        print("%s pr@cls: %f" % (mode.upper(), out_val_ints))
        out_val = out_val_joint + out_val_tr + out_val_ints
        if opt.ctx == 1:
            out_val_rels = ratio(acc[4], acc[3])
            print("%s pr@rels: %f" % (mode.upper(), out_val_rels))
            out_val += out_val_rels
    else:
        out_val_ints = out_val_joint = ratio(top[1], top[0])
        print("%s pr@1: %f" % (mode.upper(), out_val_ints))
        print("%s pr@5: %f" % (mode.upper(), ratio(top[2], top[0])))
        out_val = out_val_ints
        if opt.rels_multitask and opt.ctx == 1:
            out_val_rels = ratio(rel[1], rel[0])
            out_val += out_val_rels
            print("%s rels@top1: %f" % (mode.upper(), out_val_rels))
            print("%s rels@top3: %f" % (mode.upper(), ratio(rel[2], rel[0])))
            print("%s rel+int: %f" % (mode.upper(), out_val))
    out = {"total": out_val, "ints": out_val_ints}
    if opt.rels_multitask:
        out.update({"rels": out_val_rels})
    if opt.tr_maximize:
        out.update({"tracks": out_val_tr, "joint": out_val_joint})
    return out
