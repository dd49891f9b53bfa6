"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py [file stem ...]

For each preset / loss variant a small random model of the reference (reduced dims so the fixtures
stay small) is built with the reference's own create_model, run forward + loss + backward on a
random dense batch, and inputs, parameters, outputs, loss, the in-place masked logits and all
parameter gradients are stored.  Thirteen cases run in eval mode; five more (`*_train.npz`, one per
preset and one for the relationship-only model, opt.ints == 0) run the reference in TRAIN mode with its nn.Dropout modules replaced by a replayer
(oracle/reference_shim.py:DropoutReplay) over random 0/1 masks that are stored with the case, which
pins the dropout sites of the restatement.  tests/test_oracle_golden.py checks oracle/ against them
everywhere (the GPU box has no /root/reference).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
ONLY_NAMES = sys.argv[1:]       # optional: file stems to (re)generate, e.g. int_rels_gates-0_ints-0; default all
sys.argv = sys.argv[:1]

from oracle import reference_shim as rs  # noqa: E402

DIMS = dict(text_dim=24, visual_dim=40, track_dim=40, joint_dim=16, mid_m_ints=6)
C, R, B, T, S = 11, 5, 4, 5, 3
D = DIMS["text_dim"] + DIMS["visual_dim"] + 2 * DIMS["track_dim"]

CASES = [
    ("modalities", {}), ("modalities", dict(modality="t", tracks=False)), ("modalities", dict(modality="v", tracks=False)),
    ("modalities", dict(tracks=False)), ("int_rels", {}), ("int_ch", {}), ("int_ch", dict(tr_correct=True)),
    ("int_ch", dict(tr_max_neg=True)), ("int_rel_ch", {}), ("int_rel_ch", dict(tr_correct=True)),
    ("int_rel_ch", dict(tr_max_neg=True)), ("int_rel_ch", dict(tr_correct=True, tr_max_neg=True)),
    # train mode (dropout masks replayed): one per preset
    ("modalities", dict(train=True)), ("int_rels", dict(train=True)), ("int_ch", dict(train=True)),
    ("int_rel_ch", dict(train=True)),
    # opt.ints == 0: the context branch + relationship head alone (mlp/model.py:102, 140, 151, 208; loss :391); the
    # reference's GatingUnit needs the interaction feature, so gates are off
    ("int_rels", dict(ints=0, gates=0)), ("int_rels", dict(ints=0, gates=0, train=True)),
]



def make_batch(preset, rng):
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    b = {}
    if preset == "modalities":
        b["features"] = torch.randn(B, 1, D, generator=g, dtype=torch.float64)
        b["labels"] = torch.randint(C, (B,), generator=g)
    elif preset == "int_rels":
        b["features"] = torch.randn(B, S + 1, D, generator=g, dtype=torch.float64)
        m = torch.zeros(B, S, 1, dtype=torch.long)
        for i in range(B):
            m[i, :int(rng.integers(1, S + 1))] = 1
        b["rels_mask"] = m
        b["labels"] = torch.randint(C, (B, S + 1, 1), generator=g)
        b["rels_label"] = torch.tensor([0, R, 2, 1])[:B]          # one None-labelled sample
    else:
        ctx = preset == "int_rel_ch"
        shape = (B, T, S + 1, D) if ctx else (B, T, D)
        b["features"] = torch.randn(*shape, generator=g, dtype=torch.float64)
        mem = torch.zeros(B, T, dtype=torch.float64)
        counts = [T, 2, 3, 1][:B]
        for i, n in enumerate(counts):
            mem[i, :n] = 1
        b["mem_mask"] = mem
        b["labels"] = torch.randint(C, (B,), generator=g)
        b["gt_tracks"] = torch.tensor([[0, 2], [0, 0], [0, 1], [0, 0]])[:B]
        if ctx:
            rm = torch.zeros(B, T, S, dtype=torch.long)
            for i, n in enumerate(counts):
                for t in range(n):
                    rm[i, t, :int(rng.integers(0 if t else 1, S + 1))] = 1     # includes empty contexts
            b["rels_mask"] = rm
            rl = torch.randint(R + 1, (B, T), generator=g)
            rl = torch.where(mem.bool(), rl, torch.zeros_like(rl))              # pad label 0
            b["rels_label"] = rl
    b["multilab_weights"] = (torch.rand(B, C, generator=g) < 0.85).double()
    return b


def main():
    opt, _ = rs.load()
    rng = np.random.default_rng(0)
    for idx, (preset, over) in enumerate(CASES):
        for k, v in DIMS.items():
            setattr(opt, k, v)
        opt.mlp_dim = D
        opt.modality, opt.tracks = "m", True
        over = dict(over)
        train = bool(over.pop("train", False))
        batch = make_batch(preset, rng)          # always drawn: the cases share one generator, in this order
        stem = "%s%s" % (preset, "".join("_" + (k if v is True else "%s-%s" % (k, v))
                                          for k, v in sorted(dict(over, **({"train": True} if train else {})).items())))
        if ONLY_NAMES and stem not in ONLY_NAMES:
            continue
        opt.ints, opt.gates = 1, 1               # presets set ctx / gates; a case may override ints / gates
        model, loss = rs.create_model(preset, C, R, seed=idx, **over)
        model.eval()
        inp = {k: v.clone() for k, v in batch.items()}
        masks = None
        if train:
            kind = {"modalities": "modalities", "int_rels": "midfusion"}.get(preset, "maxtracks")
            ctx = preset in ("int_rels", "int_rel_ch")
            rows = B * T if kind == "maxtracks" else B
            masks = rs.random_masks(kind, rows, S, DIMS["joint_dim"], DIMS["joint_dim"] * DIMS["mid_m_ints"], opt.dropout,
                                    torch.Generator().manual_seed(1000 + idx), ctx=ctx,
                                    gates=ctx and bool(over.get("gates", 1)), ints=bool(over.get("ints", 1)))
            queue = rs.replay_dropout(model, masks, opt.dropout)
        out = model(batch)                     # MaxTracks reshapes batch['features'] in place
        if train:
            assert not queue
        lv = rs.run_loss(loss, out, batch)     # track losses overwrite out[...] with -inf in place
        lv.backward()
        rec = {"loss": np.float64(lv.item())}
        for k, v in inp.items():
            rec["in_" + k] = v.numpy()
        for k, v in model.state_dict().items():
            rec["p_" + k] = v.numpy()
        for k, p in model.named_parameters():
            rec["g_" + k] = p.grad.numpy()
        for k, v in out.items():
            if v is not None:
                rec["out_" + k] = v.detach().numpy()
        if train:
            over["train"] = True
            rec["dropout_p"] = np.float64(opt.dropout)
            for k, m in masks.items():
                rec["mask_" + "/".join(k)] = m.numpy()
        rec["meta"] = np.array([preset, repr(sorted(over.items()))])
        name = "%s%s.npz" % (preset, "".join("_" + (k if over[k] is True else "%s-%s" % (k, over[k])) for k in sorted(over)))
        np.savez_compressed(os.path.join(HERE, name), **rec)
        print("wrote", name, "loss", lv.item())
    # restore the full-size dims for anything else importing the shim in this process
    opt.text_dim, opt.visual_dim, opt.track_dim, opt.joint_dim, opt.mlp_dim = 768, 2048, 2048, 512, 6912
    opt.modality, opt.tracks = "m", True
    opt.ints = 1


if __name__ == "__main__":
    main()
