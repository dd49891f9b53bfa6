"""CPU restatement (plain torch, dense tensors) of the reference's losses.

TEST INFRASTRUCTURE — see oracle/__init__.py.  Follows:
  max_margin_ce()        MaxMarginCrossEntropyLoss.forward   mlp/model.py:427-441
  multitask_max_margin() MultiTaskMaxMargin.forward          mlp/model.py:387-419
  margin_loss()          MarginLoss.forward                  mlp/model.py:450-494
  margin_track_rels()    MarginTrackRelsLoss.forward         mlp/model.py:503-575
  multitask_ce()         MultiTaskCrossEntropyLoss.forward   mlp/model.py:367-378
Masks are boolean tensors (the reference's uint8 masks had torch-1.1 logical semantics,
SURVEY.md §0 "oracle hazard").  Like the reference, the two track losses overwrite the logits of
empty slots with -inf; here that is done on a copy and the masked logits are returned too.
tr_cat_distr (multinomial assignment, model.py:468-471, 540-543) is RNG-dependent and not
restated.
"""
import torch

NEG_INF = float("-inf")


def _hinge_rows(scores, target, neg_mask, margin):
    """sum_c relu(m - s[y] + s[c]) over the negatives of each row."""
    idx = torch.arange(scores.shape[0])
    pos = scores[idx, target]
    nm = neg_mask.to(scores.dtype)
    return (torch.relu((margin - pos).view(-1, 1) + scores * nm) * nm).sum(1)


def max_margin_ce(inters, labels, multilab_weights, margin):
    B, C = inters.shape
    neg = torch.ones(B, C, dtype=torch.bool)
    neg[torch.arange(B), labels] = False
    neg &= multilab_weights.bool()
    return _hinge_rows(torch.sigmoid(inters), labels, neg, margin).mean()


def multitask_max_margin(inters, rels, labels, rels_label, multilab_weights, margin, lymbda, n_rels,
                         ints=1, ctx=1):
    """inters [B, C] (or [B*k, C] viewed as [B, k, C]); labels [B, k, 1]; rels [B, R]."""
    loss = torch.zeros((), dtype=inters.dtype if inters is not None else rels.dtype)
    B = rels_label.shape[0]
    if ints:
        x = inters.view(B, -1, inters.shape[-1])[:, 0]
        y = labels[:, 0].reshape(-1)
        loss = loss + lymbda * max_margin_ce(x, y, multilab_weights, margin)
    if ctx:
        sel = (rels_label != n_rels).nonzero().reshape(-1)
        if sel.numel():
            y = rels_label[sel]
            r = torch.sigmoid(rels[sel])
            neg = torch.ones_like(r, dtype=torch.bool)
            neg[torch.arange(sel.numel()), y] = False
            loss = loss + _hinge_rows(r, y, neg, margin).mean()
    return loss


def _track_ints_part(x, labels, mem_mask, multilab_weights, gt_tracks, tr_correct):
    """Masked copy of the interaction logits and their negatives mask."""
    B, T, C = x.shape
    valid = mem_mask.bool().view(B, T, 1).expand(B, T, C)
    x = x.masked_fill(~valid, NEG_INF)                         # model.py:459-460 / 510-512
    neg = valid & multilab_weights.bool().view(B, 1, C)
    b = torch.arange(B)
    if tr_correct:
        neg[b, gt_tracks[:, 0], labels] = False                # model.py:463-465
        neg[b, gt_tracks[:, 1], labels] = False
    else:
        neg[b, :, labels] = False                              # model.py:467
    return x, neg


def _track_hinge(s, pos, neg, margin, max_neg):
    """s [B, T, C] sigmoid scores, pos [B], neg [B, T, C] -> per-clip loss [B]."""
    B = s.shape[0]
    nm = neg.to(s.dtype)
    if max_neg:                                                # model.py:483-486
        hardest = (s * nm).max(dim=2)[0]
        return torch.relu((margin - pos).view(-1, 1) + hardest).sum(1)
    flat = (s * nm).view(B, -1)
    return (torch.relu((margin - pos).view(-1, 1) + flat) * nm.view(B, -1)).sum(1)


def cat_distr_probs(x_masked, labels, r_masked=None, r0=None):
    """Sampling distribution of opt.tr_cat_distr (model.py:468-471, 538-543): softmax over the slots of the
    target-class logits, averaged with the softmax of the GT-relationship logits (NaN -> 0) if given."""
    b = torch.arange(x_masked.shape[0])
    p = torch.softmax(x_masked[b, :, labels], dim=1)
    if r_masked is not None:
        pr = torch.softmax(r_masked[b, :, r0], dim=1)
        pr = torch.where(pr != pr, torch.zeros_like(pr), pr)
        p = (p + pr) / 2
    return p


def margin_loss(inters, labels, mem_mask, multilab_weights, gt_tracks, margin, tr_correct=False,
                max_neg=False, assign=None):
    """Returns (loss, assignment t*, masked logits).  assign: forced t* (tr_cat_distr draws)."""
    B = inters.shape[0]
    x, neg = _track_ints_part(inters, labels, mem_mask, multilab_weights, gt_tracks, tr_correct)
    s = torch.sigmoid(x)
    b = torch.arange(B)
    if assign is not None:
        tstar = assign
    elif tr_correct:
        tstar = torch.zeros(B, dtype=torch.long)
    else:
        tstar = torch.argmax(s[b, :, labels] * mem_mask.to(s.dtype), dim=1)   # model.py:479
    pos = s[b, tstar, labels]
    return _track_hinge(s, pos, neg, margin, max_neg).mean(), tstar, x


def margin_track_rels(inters, rels, labels, rels_label, mem_mask, multilab_weights, gt_tracks, margin,
                      lymbda, n_rels, tr_correct=False, max_neg=False, assign=None):
    """Returns (loss, assignment t*, masked inters, masked rels with the appended None column)."""
    B, T, R = rels.shape
    b = torch.arange(B)
    x, neg_i = _track_ints_part(inters, labels, mem_mask, multilab_weights, gt_tracks, tr_correct)
    # relationships: append the None column, mask empty slots, None-labelled slots and the None column
    live = mem_mask.bool().view(B, T, 1) & (rels_label != n_rels).view(B, T, 1)   # model.py:516-520
    neg_r = torch.cat((live.expand(B, T, R), torch.zeros(B, T, 1, dtype=torch.bool)), dim=-1)
    r = torch.cat((rels, torch.zeros(B, T, 1, dtype=rels.dtype)), dim=-1)
    r = r.masked_fill(~neg_r, NEG_INF)                          # model.py:521-524
    neg_r = neg_r.clone()
    r0 = rels_label[b, gt_tracks[:, 0]]
    r1 = rels_label[b, gt_tracks[:, 1]]
    if tr_correct:
        flat = neg_r.view(-1, R + 1)
        flat[torch.arange(flat.shape[0]), rels_label.reshape(-1)] = False       # model.py:531-533
        neg_r = flat.view(B, T, R + 1)
    else:
        neg_r[b, :, r0] = False                                 # model.py:536-537
        neg_r[b, :, r1] = False
    s_i, s_r = torch.sigmoid(x), torch.sigmoid(r)
    if assign is not None:
        tstar = assign
    elif tr_correct:
        tstar = torch.zeros(B, dtype=torch.long)
    else:
        score = s_i[b, :, labels] + s_r[b, :, r0]
        tstar = torch.argmax(score * mem_mask.to(score.dtype), dim=1)           # model.py:552-553
    pos_i = s_i[b, tstar, labels]
    pos_r = s_r[b, tstar, r0]
    loss = lymbda * _track_hinge(s_i, pos_i, neg_i, margin, max_neg).mean() \
        + _track_hinge(s_r, pos_r, neg_r, margin, max_neg).mean()
    return loss, tstar, x, r


def multitask_ce(inters, rels, labels, rels_label, n_rels, weights=None):
    import torch.nn.functional as F
    loss = F.cross_entropy(inters, labels.reshape(-1), weight=weights)
    sel = (rels_label != n_rels).nonzero().reshape(-1)
    if sel.numel():
        loss = loss + F.cross_entropy(rels[sel], rels_label[sel])
    return loss
