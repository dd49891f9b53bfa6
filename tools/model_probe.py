"""GPU probe: full model + loss forward/backward vs the dense fp64 oracle on one synthetic batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = sys.argv[:1]
import numpy as np, torch
from lirec_b200.utils.arg_pars import opt
from lirec_b200.mixed_utils import synthetic
from oracle import model as omodel, losses as olosses, dropout as odrop


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def run(preset, B, train, seed=3):
    flags = dict(modalities=dict(mod_check=True, tr_maximize=False, ints=1, ctx=0, gates=0, rels_multitask=False),
                 int_rels=dict(mod_check=False, tr_maximize=False, ints=1, ctx=1, gates=1, rels_multitask=True),
                 int_ch=dict(mod_check=False, tr_maximize=True, ints=1, ctx=0, gates=0, rels_multitask=False),
                 int_rel_ch=dict(mod_check=False, tr_maximize=True, ints=1, ctx=1, gates=1, rels_multitask=True))[preset]
    for k, v in flags.items():
        setattr(opt, k, v)
    opt.tracks, opt.device, opt.modality = True, "cuda", "m"
    import lirec_b200.mlp.model as M
    import io, contextlib
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model, loss, _ = M.create_model(101, n_rels=15)
    pb = synthetic.make_batch(B, seed=seed, preset=preset)
    model.train(train)
    out = model(pb.to_device("cuda"), seed=1234)
    lv = loss(out, {})
    lv.backward()
    torch.cuda.synchronize()
    # ---- oracle on the dense batch with identically rounded operands ----
    sd = {}
    for k, v in model.state_dict().items():
        v = v.detach().cpu()
        sd[k] = (v.to(torch.bfloat16) if k.endswith("weight") else v).double().requires_grad_(True)
    dense = pb.to_dense(np.float64)
    cfg = omodel.default_cfg(ints=1, ctx=flags["ctx"], gates=flags["gates"], dropout=opt.dropout)
    kind = synthetic.PRESETS[preset]["kind"]
    masks = odrop.dense_masks(pb, 1234, opt.dropout, kind=kind) if train else None
    f = dense["features"]
    if kind == "modalities":
        o = omodel.modalities_forward(sd, f.reshape(B, 1, -1), cfg, masks)
        l = olosses.max_margin_ce(o["inters"], dense["labels"], dense["multilab_weights"], opt.margin)
        pairs = [("inters", out.ragged_inters, o["inters"])]
    elif kind == "midfusion":
        o = omodel.midfusion_forward(sd, f.reshape(B, f.shape[2], -1), dense["rels_mask"].reshape(B, -1, 1), cfg, masks)
        l = olosses.multitask_max_margin(o["inters"], o["rels"], dense["labels"].reshape(B, 1, 1).expand(B, 2, 1),
                                         dense["rels_label"].reshape(B), dense["multilab_weights"], opt.margin,
                                         opt.lymbda, 15)
        pairs = [("inters", out.ragged_inters, o["inters"]), ("rels", out.ragged_rels, o["rels"])]
    else:
        o = omodel.maxtracks_forward(sd, f, dense.get("rels_mask"), cfg, masks)
        mm = dense["mem_mask"].bool()
        if flags["ctx"]:
            l, ts, _, _ = olosses.margin_track_rels(o["inters"], o["rels"], dense["labels"], dense["rels_label"],
                                                    dense["mem_mask"], dense["multilab_weights"], dense["gt_tracks"],
                                                    opt.tr_margin, opt.lymbda, 15)
            pairs = [("inters", out.ragged_inters, o["inters"][mm]), ("rels", out.ragged_rels, o["rels"][mm])]
        else:
            l, ts, _ = olosses.margin_loss(o["inters"], dense["labels"], dense["mem_mask"], dense["multilab_weights"],
                                           dense["gt_tracks"], opt.tr_margin)
            pairs = [("inters", out.ragged_inters, o["inters"][mm])]
        print("   assignment equal:", bool((loss.last_assignment.cpu().long() == ts).all()))
    l.backward()
    print("== %s B=%d train=%s  Ni=%d Nx=%d n_clip=%d n_track=%d" % (preset, B, train, pb.n_cand, pb.n_ctx_rows, pb.n_clip, pb.n_track))
    for name, a, b in pairs:
        print("   %-8s rel err %.3e" % (name, rel(a.detach(), b.detach())))
    print("   loss ours %.6f oracle %.6f rel %.3e" % (lv.item(), l.item(), abs(lv.item() - l.item()) / abs(l.item())))
    worst = 0
    for k, p in model.named_parameters():
        e = rel(p.grad, sd[k].grad)
        worst = max(worst, e)
        if e > 1e-3 or os.environ.get("VERBOSE"):
            print("   grad %-28s rel err %.3e  |g|max %.3e" % (k, e, sd[k].grad.abs().max().item()))
    print("   worst grad rel err %.3e" % worst)
    sys.stdout.flush()


if __name__ == "__main__":
    for preset in ("int_rel_ch", "int_ch", "int_rels", "modalities"):
        for train in (False, True):
            try:
                run(preset, 6, train)
            except Exception as e:
                import traceback; traceback.print_exc()
