"""How many bf16 passes do the gradient GEMMs need for the 1e-3 bar?  CPU emulation on the fp64 oracle:
a hi/lo-split product dropped to fewer passes is the exact product with one operand rounded to bf16.
    python tools/precision_study.py [B]
"""
import os
import sys

import numpy as np
import torch

sys.argv, ARGS = sys.argv[:1], sys.argv[1:]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lirec_b200.mixed_utils import synthetic  # noqa: E402
from oracle import dropout as odrop, losses as ol, model as om  # noqa: E402


def bf(x):
    return x.to(torch.bfloat16).to(x.dtype)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    B = int(ARGS[0]) if ARGS else 32
    torch.manual_seed(0)
    cfg = om.default_cfg(dropout=0.3)
    sd = {k: (bf(v) if k.endswith("weight") else v).double().requires_grad_(True)
          for k, v in om.init_state_dict(cfg, 101, 15, "maxtracks", seed=0).items()}
    pb = synthetic.make_batch(B, seed=3, preset="int_rel_ch")
    dense = pb.to_dense(np.float64)
    masks = odrop.dense_masks(pb, 77, 0.3)
    cfg.tape = {}
    o = om.maxtracks_forward(sd, dense["features"], dense["rels_mask"], cfg, masks)
    l, *_ = ol.margin_track_rels(o["inters"], o["rels"], dense["labels"], dense["rels_label"], dense["mem_mask"],
                                 dense["multilab_weights"], dense["gt_tracks"], 0.101, 1.0, 15)
    l.backward()
    t = cfg.tape
    z, dpre = t["gate_in"].detach(), t["pre_gate"].grad
    gw = sd["gates_ints.fc_out.weight"].grad
    print("gate wgrad: exact-vs-autograd %.1e | drop x_lo: %.2e | drop dy_lo: %.2e | single pass: %.2e" % (
        rel(dpre.t() @ z, gw), rel(dpre.t() @ bf(z), gw), rel(bf(dpre).t() @ z, gw), rel(bf(dpre).t() @ bf(z), gw)))
    # gate dgrad: d z = dpre @ W ; single pass = bf16(dpre) @ W
    W = sd["gates_ints.fc_out.weight"].detach()
    dz = dpre @ W
    print("gate dgrad: drop dy_lo: %.2e" % rel(bf(dpre) @ W, dz))
    # first-layer wgrad (x is exact bf16): drop dy_lo
    for slot, sl in (("txt", slice(0, 768)), ("vis", slice(768, 2816)), ("tracks1", slice(2816, 4864))):
        for br in ("ints", "ctx"):
            dz1 = t["z1_%s_%s" % (slot, br)].grad
            x = dense["features"].reshape(-1, 19, 6912)
            x = x[:, 0, sl] if br == "ints" else x[:, 1:, sl]
            x = x.reshape(-1, x.shape[-1]).double()
            dz1 = dz1.reshape(-1, dz1.shape[-1])
            g = sd["%s_%s.weight" % (slot, br)].grad
            print("L1 wgrad %-8s %-4s exact %.1e | single pass (bf16 dy): %.2e" % (
                slot, br, rel(dz1.t() @ x, g), rel(bf(dz1).t() @ x, g)))


if __name__ == "__main__":
    main()
