import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = sys.argv[:1]
import numpy as np, torch, io, contextlib
from lirec_b200.utils.arg_pars import opt
from lirec_b200.mixed_utils import synthetic
from lirec_b200 import _ext
from oracle import model as omodel, losses as olosses, dropout as odrop
for k, v in dict(mod_check=False, tr_maximize=True, ints=1, ctx=1, gates=1, rels_multitask=True).items(): setattr(opt, k, v)
opt.tracks, opt.device, opt.modality = True, "cuda", "m"
import lirec_b200.mlp.model as M

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()

def run(train):
    B = 6
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model, loss, _ = M.create_model(101, n_rels=15)
    pb = synthetic.make_batch(B, seed=3, preset="int_rel_ch")
    model.train(train)
    pbd = pb.to_device("cuda")
    out = model(pbd, seed=1234)
    fn = out.ragged_inters.grad_fn
    lv = loss(out, {})
    # grab ctx of _ModelFn before backward frees ws
    ws = fn.ws; batch_c = fn.batch_c
    lv.backward(); torch.cuda.synchronize()
    offs = (C.c_int64 * 40)()
    n = _ext.lib().lirec_model_workspace_layout(C.byref(model._cfg_c), C.byref(batch_c), offs, 40)
    names = ["r1_%d_%d" % (b, s) for b in range(2) for s in range(4)] + ["dz1_%d_%d" % (b, s) for b in range(2) for s in range(4)] + \
        ["a2_0", "a2_1", "f2_0", "f2_1", "dz2_0", "dz2_1", "da2_0", "da2_1", "flag_c", "flag_bf16", "ones", "g2", "dpreg2", "dli2", "dlr2"]
    off = dict(zip(names, list(offs)[:n]))
    Ni, J, F, Gd = pb.n_cand, 512, 1536, 3072
    def bf(name, rows, cols):
        t = ws[off[name]: off[name] + rows * cols * 2].view(torch.bfloat16).view(rows, cols).float()
        return t
    def merged(name, rows, w):
        t = bf(name, rows, 2 * w); return t[:, :w] + t[:, w:]
    sd = {}
    for k, v in model.state_dict().items():
        v = v.detach().cpu()
        sd[k] = (v.to(torch.bfloat16) if k.endswith("weight") else v).double().requires_grad_(True)
    dense = pb.to_dense(np.float64)
    cfg = omodel.default_cfg(dropout=opt.dropout); cfg.tape = {}
    masks = odrop.dense_masks(pb, 1234, opt.dropout) if train else None
    o = omodel.maxtracks_forward(sd, dense["features"], dense["rels_mask"], cfg, masks)
    l, ts, _, _ = olosses.margin_track_rels(o["inters"], o["rels"], dense["labels"], dense["rels_label"], dense["mem_mask"],
                                            dense["multilab_weights"], dense["gt_tracks"], opt.tr_margin, opt.lymbda, 15)
    o['inters'].retain_grad(); l.backward()
    mm = dense["mem_mask"].bool().reshape(-1)
    tp = cfg.tape
    print("== train", train)
    print("g2 vs relu/drop(pre_gate):", rel(merged("g2", Ni, Gd), (torch.relu(tp["pre_gate"]) * (masks[("gate",)].double() / 0.7 if train else 1.0))[mm].detach()))
    print("dpreg2 vs oracle:", rel(merged("dpreg2", Ni, Gd), tp["pre_gate"].grad[mm]))
    d = (merged("dpreg2", Ni, Gd).cpu().double() - tp["pre_gate"].grad[mm]).abs()
    print("   rows with err:", (d.max(1)[0] > 1e-3 * tp["pre_gate"].grad.abs().max()).nonzero().reshape(-1).tolist())
    ours = merged("dpreg2", Ni, Gd).cpu().double(); orc = tp["pre_gate"].grad[mm]
    nz = orc[1].abs() > 1e-9
    print("   row1 ours[:6]", ours[1][nz][:6].tolist()); print("   row1 orc [:6]", orc[1][nz][:6].tolist())
    print("   row0 ours[:6]", ours[0][nz][:6].tolist()); print("   row0 orc [:6]", orc[0][nz][:6].tolist())
    print("   mask mismatch row1:", ((ours[1] != 0) != (orc[1] != 0)).sum().item(), " rows0==rows1 feature identical:", bool((dense["features"][0,0,0] == dense["features"][0,1,0]).all()))
    dl = bf("dli2", Ni, 256); dlm = (dl[:, :128] + dl[:, 128:]).cpu().double()
    print("   d_ints row1 ours[:5]", dlm[1][:5].tolist())
    og = o["inters"].grad.reshape(-1, 101)[mm]
    print("   d_ints row1 orc [:5]", og[1][:5].tolist(), " d_ints rel err all rows:", rel(dlm[:, :101], og))
    print("dz2_ints vs oracle:", rel(merged("dz2_0", Ni, F), tp["z2_ints"].grad[mm]))
    print("dz2_ctx vs oracle:", rel(merged("dz2_1", Ni, F), tp["z2_ctx"].grad[mm]))
    d = (merged("dz2_1", Ni, F).cpu().double() - tp["z2_ctx"].grad[mm]).abs()
    print("   rows with err:", (d.max(1)[0] > 1e-3 * tp["z2_ctx"].grad.abs().max()).nonzero().reshape(-1).tolist())
    print("dli2 vs ours d_ints: (check pad) max pad", bf("dli2", Ni, 256)[:, 101:128].abs().max().item(), bf("dli2", Ni, 256)[:, 128 + 101:].abs().max().item())
    # row duplicates?
    f = dense["features"][dense["mem_mask"].bool()][:, 0]
    print("   cand_off", pb.tables["cand_off"].tolist())

run(False)
run(True)
