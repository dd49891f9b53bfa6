"""Evaluation loop (reference: mlp/test.py:17-145): model.eval(), no_grad, loss, accuracy counters,
returns {'total', 'ints', 'rels', 'tracks', 'joint'}.

The reference moves the logits to the host every batch and arg-maxes them in numpy / scipy
(utils/evaluation.py).  Here the prediction arg-maxes run on the GPU over the ragged logits
(lirec_predict_tracks, bit-exact against the numpy formulas — tests/test_loss_gpu.py) and only
integer counters come back, once per evaluation.  The counters are the reference's
(lirec_b200/utils/evaluation.py replays Precision / RelationshipsAcc bookkeeping on the device and is
pinned against the unmodified meters, tests/test_eval_cpu.py); HTML dumps and confusion matrices are
not reproduced."""
import torch

from lirec_b200 import dp, ops
from lirec_b200.mixed_utils.classification_dataloader import packed_loader
from lirec_b200.utils.arg_pars import opt
from lirec_b200.utils.evaluation import RelationshipsAcc, TopKMeters, TrackMeters
from lirec_b200.utils.util_functions import Averaging


def _update_track_meters(meters, out, pb, n_rels):
    """Device arg-maxes of one batch -> the reference's track / class / relationship / joint counters."""
    pred = ops.predict_tracks(out.ragged_inters, out.ragged_rels, pb["cand_off"], pb["labels"],
                              pb["rels_label"] if n_rels else None, pb["gt_tracks"], n_rels)
    dev = pred.device
    jz = pb.extras.get("just_zeros")
    jz = torch.zeros(pb.B, dtype=torch.bool, device=dev) if jz is None else torch.as_tensor(jz).to(dev).bool()
    gt_rel = rel_at_gt = None
    if n_rels:
        first = pb["cand_off"][:-1].long()
        gt = pb["gt_tracks"].long()
        rl = pb["rels_label"].long()
        gt_rel = rl[first]                                        # rels_label[:, 0] (evaluation.py:208)
        rel_at_gt = torch.stack((rl[first + gt[:, 0]], rl[first + gt[:, 1]]), dim=1)
    meters.update(pred, pb["labels"], pb["gt_tracks"], jz, gt_rel=gt_rel, rel_at_gt=rel_at_gt, n_rels=n_rels)


def testing(test_dataset, model, loss, total_iter=1, mode="val", train_start_time=""):
    rank, world = (torch.distributed.get_rank(), dp.world_size()) if dp.world_size() > 1 else (0, 1)
    losses = Averaging()
    model.eval()
    n_rels = test_dataset.n_rels - 1 if opt.rels_multitask else 0
    dev = next(model.parameters()).device
    tracks, top = TrackMeters(dev), TopKMeters(dev)
    n_hash = len(getattr(test_dataset, "hashidx_rels", ())) or None
    racc = RelationshipsAcc(max(n_rels, 1), n_hash, dev) if (n_hash and not opt.tr_maximize) else None
    rel = TopKMeters(dev)                                         # per-row relationship ranking (no pair ids)
    loss_sum = torch.zeros((), device=dev)
    n_batches = 0
    with torch.no_grad():
        for pb in packed_loader(test_dataset, opt.batch_size, shuffle=False, num_workers=opt.num_workers,
                                device=dev, rank=rank, world=world):
            if getattr(pb.host, "global_clips", pb.B) == 1:     # reference skips batches of one (:38-39)
                continue
            if pb.B == 0:                                        # EmptyShard: no per-batch collective in evaluation
                continue
            out = model(pb)
            loss_sum += loss(out, {}) * pb.B
            n_batches += 1
            if opt.tr_maximize:
                _update_track_meters(tracks, out, pb, n_rels if opt.ctx == 1 else 0)
            else:
                if opt.ints == 1:                                # mlp/test.py:74-79
                    top.update(out.ragged_inters, pb["labels"])
                else:                                            # relationship-only model: count the clips
                    top.c[0] += pb.B
                if opt.rels_multitask and opt.ctx == 1:
                    lab = pb["rels_label"].long()
                    sel = (lab != n_rels).nonzero().reshape(-1)          # mlp/test.py:80-87
                    if sel.numel():
                        rel.update(out.ragged_rels[sel], lab[sel])
                        h = pb.extras.get("hash_rel")
                        if racc is not None and h is not None:
                            racc.update(out.ragged_rels[sel], lab[sel], torch.as_tensor(h).to(dev)[sel])
    stats = torch.cat([tracks.c, top.c, rel.c]).double()
    if racc is not None and world > 1:
        torch.distributed.all_reduce(racc.sum)
        torch.distributed.all_reduce(racc.gt, op=torch.distributed.ReduceOp.MAX)
    stats = torch.cat([stats, loss_sum.double().view(1)])
    if world > 1:
        torch.distributed.all_reduce(stats)                     # a few integers per evaluation
    s = stats.tolist()                                          # the only device->host read
    tc = dict(zip(TrackMeters.NAMES, s[0:7]))
    top_c, rel_c, loss_total = s[7:11], s[11:15], s[15]
    n_clips = tc["total_cl"] if opt.tr_maximize else top_c[0]
    losses.update(loss_total / max(n_clips, 1), int(max(n_clips, 1)))

    def ratio(a, b):
        return float(a) / float(b) if b else 0.0
    out_val = out_val_ints = out_val_rels = out_val_tr = out_val_joint = 0.0
    if opt.ints == 1:                                           # the reference prints the loss with pr@1 (:102-103)
        print("%s loss: %f" % (mode.upper(), losses.avg))
    if opt.tr_maximize:
        r = TrackMeters.ratios(tc)                              # the reference's accessors (evaluation.py:329-360)
        out_val_joint, out_val_tr, out_val_ints = r["top1"], r["trks_top1"], r["cls_top1"]
        print("%s pr@1: %f" % (mode.upper(), out_val_joint))
        print("%s pr@trks: %f" % (mode.upper(), out_val_tr))
        print("%s pr@cls: %f" % (mode.upper(), out_val_ints))
        out_val = out_val_joint + out_val_tr + out_val_ints    # mlp/test.py:104-118
        if opt.ctx == 1:
            out_val_rels = r["rels_top1"]
            print("%s pr@rels: %f" % (mode.upper(), out_val_rels))
            out_val += out_val_rels
    else:
        if opt.ints == 1:                                       # mlp/test.py:102-109
            out_val_ints = out_val_joint = ratio(top_c[1], top_c[0])
            print("%s pr@1: %f" % (mode.upper(), out_val_ints))
            print("%s pr@5: %f" % (mode.upper(), ratio(top_c[3], top_c[0])))
            out_val = out_val_ints
        if opt.rels_multitask and opt.ctx == 1:
            if racc is not None:                                # pair-level ranking, as the reference (evaluation.py:399-417)
                rc = racc.compute()
                r1, r3 = ratio(rc["top1"], rc["total"]), ratio(rc["top3"], rc["total"])
            else:                                               # dataset without pair ids: per-row ranking
                r1, r3 = ratio(rel_c[1], rel_c[0]), ratio(rel_c[2], rel_c[0])
            out_val_rels = r1
            out_val += out_val_rels
            print("%s rels@top1: %f" % (mode.upper(), r1))
            print("%s rels@top3: %f" % (mode.upper(), r3))
            print("%s rel+int: %f" % (mode.upper(), out_val))
    out = {"total": out_val, "ints": out_val_ints}
    if opt.rels_multitask:
        out.update({"rels": out_val_rels})
    if opt.tr_maximize:
        out.update({"tracks": out_val_tr, "joint": out_val_joint})
    return out
