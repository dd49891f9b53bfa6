"""CPU restatement (plain torch, dense tensors) of the reference model family.

TEST INFRASTRUCTURE — see oracle/__init__.py.  Follows, function by function:
  encode()              the modality encoder inlined at mlp/model.py:59-76, 152-167, 177-196,
                        279-294, 305-322
  modalities_forward()  Modalities.forward                    mlp/model.py:54-92
  midfusion_forward()   MidFusionMultiClip.forward            mlp/model.py:147-211
  maxtracks_forward()   MidFusionMultiClipMaxTracks.forward   mlp/model.py:265-339
  gating_unit()         GatingUnit.forward                    mlp/model.py:349-354
It is written over a `state_dict` with the reference's parameter names, on DENSE zero-padded
batches exactly as the reference dataloader emits them, so it is a restatement of the
reference's schedule — not of the packed/deduplicated schedule the CUDA path uses.

Dropout: the reference draws masks from torch's global RNG.  Here `masks` (a dict of 0/1
tensors, see keys below) makes train mode deterministic; masks=None means eval mode and
masks="rng" draws them from torch's RNG like the reference (used by the CPU baseline timing).
  ('l1', branch, slot)  branch in {'ints','ctx'}, slot in {'txt','vis','tracks1','tracks2'}
  ('cat', branch)       after tanh of the concatenated feature
  ('gate',)             after relu of the gating unit
"""
from types import SimpleNamespace

import torch

SLOTS = ("txt", "vis", "tracks1", "tracks2")
SECOND = {"txt": "txt2", "vis": "vis2", "tracks1": "tracks12", "tracks2": "tracks22"}


def default_cfg(**kw):
    cfg = SimpleNamespace(text_dim=768, visual_dim=2048, track_dim=2048, joint_dim=512, mid_m_ints=6,
                          ints=1, ctx=1, gates=1, dropout=0.3, modality="m", tracks=True)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def _lin(sd, name, x):
    return x @ sd[name + ".weight"].t() + sd[name + ".bias"]


def _drop(x, masks, key, p):
    if masks is None:
        return x
    if masks == "rng":      # train mode with torch's global RNG, as the reference runs it
        return torch.nn.functional.dropout(x, p=p, training=True)
    return x * masks[key].to(x.dtype) / (1.0 - p)


def _relu(cfg, name, z, dropped):
    """relu(dropped) where dropped = dropout(z) (or z itself).  With cfg.relu_gate = {name: bool tensor}
    the on/off decision of every unit whose pre-activation lies within cfg.relu_edge of zero is TAKEN FROM
    THAT TENSOR instead of from the sign of z: a unit that close to zero is a knife-edge on which an fp32 and
    an fp64 evaluation of the same formula may legitimately land on different sides (its activation is ~0
    either way; only d relu/dz flips), so the large-batch parity tests replay the decisions of the path under
    test there.  Everywhere else — and without cfg.relu_gate — this is torch.relu."""
    gates = getattr(cfg, "relu_gate", None)
    if not gates or name not in gates:
        return torch.relu(dropped)
    edge = z.detach().abs() < getattr(cfg, "relu_edge", 1e-5)
    g = torch.where(edge, gates[name].reshape(z.shape), z.detach() > 0)
    return dropped * g.to(dropped.dtype)


def _tape(cfg, name, t):
    """Optional white-box tape (cfg.tape = {}) used by the stage-level parity tests."""
    tape = getattr(cfg, "tape", None)
    if tape is not None:
        if t.requires_grad:
            t.retain_grad()
        tape[name] = t
    return t


def encode(sd, branch, x, cfg, masks, slots=SLOTS):
    """x[..., text|visual|track1|track2] -> list of the second-layer outputs of `slots`."""
    T, V, P = cfg.text_dim, cfg.visual_dim, cfg.track_dim
    parts = {"txt": x[..., :T], "vis": x[..., T:T + V],
             "tracks1": x[..., T + V:T + V + P], "tracks2": x[..., T + V + P:T + V + 2 * P]}
    outs = []
    for slot in slots:
        h = _tape(cfg, "z1_%s_%s" % (slot, branch), _lin(sd, "%s_%s" % (slot, branch), parts[slot]))
        h = _relu(cfg, "z1_%s_%s" % (slot, branch), h, _drop(h, masks, ("l1", branch, slot), cfg.dropout))  # relu(dropout(.)): model.py:62
        outs.append(_lin(sd, "%s_%s" % (SECOND[slot], branch), h))
    return outs


def gating_unit(sd, feat_ints, feat_ctx, cfg, masks):
    z = _tape(cfg, "gate_in", torch.cat((feat_ctx, feat_ints), dim=-1))   # (rels, inters) order: model.py:352
    pre = _tape(cfg, "pre_gate", _lin(sd, "gates_ints.fc_out", z))
    z = _relu(cfg, "pre_gate", pre, pre)
    return _drop(z, masks, ("gate",), cfg.dropout)             # dropout(relu(.)): model.py:353


def modalities_forward(sd, features, cfg, masks=None):
    """features [B, 1, D] -> inters [B, C].  cfg.modality in m / t / v and cfg.tracks select the slots
    (model.py:27-46, 57-86; 't' / 'v' only work without tracks in the reference)."""
    x = features[:, 0, :]
    slots = [s for s in SLOTS if (s == "txt" and cfg.modality in ("m", "t")) or (s == "vis" and cfg.modality in ("m", "v"))
             or (s.startswith("tracks") and cfg.tracks)]
    f = torch.cat(encode(sd, "ints", x, cfg, masks, slots), dim=-1)
    f = _drop(torch.tanh(f), masks, ("cat", "ints"), cfg.dropout)
    return {"inters": _lin(sd, "out_ints", f)}


def _ctx_feature(sd, ctx_rows, rels_mask, cfg, masks, guard_zero):
    """ctx_rows [N, S, D], rels_mask [N, S] -> dropout(tanh(cat(masked means))) [N, 3J]."""
    m = rels_mask.to(ctx_rows.dtype).unsqueeze(-1)             # [N, S, 1]
    div = m.sum(1)                                             # [N, 1]
    if guard_zero:
        div = torch.where(div == 0, torch.ones_like(div), div)  # model.py:303
    pooled = [(o * m).sum(1) / div for o in encode(sd, "ctx", ctx_rows, cfg, masks)]
    f = _tape(cfg, "z2_ctx", torch.cat(pooled, dim=-1))
    return _drop(torch.tanh(f), masks, ("cat", "ctx"), cfg.dropout)


def midfusion_forward(sd, features, rels_mask, cfg, masks=None):
    """features [B, S+1, D], rels_mask [B, S, 1] -> inters [B, C], rels [B, R]."""
    out_i = out_c = None
    if cfg.ints:
        f_i = torch.cat(encode(sd, "ints", features[:, 0, :], cfg, masks), dim=-1)
        f_i = _drop(torch.tanh(f_i), masks, ("cat", "ints"), cfg.dropout)
    if cfg.ctx:
        f_c = _ctx_feature(sd, features[:, 1:, :], rels_mask.reshape(features.shape[0], -1), cfg, masks,
                           guard_zero=False)                  # no divider guard: model.py:175
    if cfg.gates:
        f_i = gating_unit(sd, f_i, f_c, cfg, masks)
    if cfg.ctx:
        out_c = _lin(sd, "out_ctx", f_c)
    if cfg.ints:
        out_i = _lin(sd, "out_ints", f_i)
    return {"inters": out_i, "rels": out_c}


def maxtracks_forward(sd, features, rels_mask, cfg, masks=None):
    """features [B, T, S+1, D] (ctx) or [B, T, D] (no ctx); rels_mask [B, T, S].
    -> inters [B, T, C], rels [B, T, R] or None."""
    B, T = features.shape[0], features.shape[1]
    x = features.reshape(B * T, -1, features.shape[-1])       # model.py:272-274
    out_c = None
    out_i = f_i = None
    if cfg.ints:                                              # model.py:278
        f_i = _tape(cfg, "z2_ints", torch.cat(encode(sd, "ints", x[:, 0, :], cfg, masks), dim=-1))
        f_i = _drop(torch.tanh(f_i), masks, ("cat", "ints"), cfg.dropout)
    if cfg.ctx:
        f_c = _ctx_feature(sd, x[:, 1:, :], rels_mask.reshape(B * T, -1), cfg, masks, guard_zero=True)
    if cfg.gates:
        f_i = gating_unit(sd, f_i, f_c, cfg, masks)
    if cfg.ctx:
        out_c = _lin(sd, "out_ctx", f_c).reshape(B, T, -1)
    if cfg.ints:                                              # model.py:335
        out_i = _lin(sd, "out_ints", f_i).reshape(B, T, -1)
    return {"inters": out_i, "rels": out_c}


def init_state_dict(cfg, n_classes, n_rels, kind, seed=0, dtype=torch.float32):
    """Random-init parameters with the reference's names/shapes and nn.Linear's default init
    (kaiming-uniform weight, uniform bias).  kind in {'modalities','midfusion','maxtracks'}."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, out_f, in_f):
        bound = 1.0 / (in_f ** 0.5)
        sd[name + ".weight"] = ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound).to(dtype)
        sd[name + ".bias"] = ((torch.rand(out_f, generator=g) * 2 - 1) * bound).to(dtype)

    J = cfg.joint_dim
    ins = {"txt": cfg.text_dim, "vis": cfg.visual_dim, "tracks1": cfg.track_dim, "tracks2": cfg.track_dim}
    branches = (["ints"] if (kind == "modalities" or cfg.ints) else []) + (["ctx"] if (kind != "modalities" and cfg.ctx) else [])
    for br in branches:
        for slot in SLOTS:
            lin("%s_%s" % (slot, br), J, ins[slot])
        for slot in SLOTS:
            lin("%s_%s" % (SECOND[slot], br), J if slot in ("txt", "vis") else J // 2, J)
    width = 3 * J
    if kind != "modalities" and cfg.gates:
        width = J * cfg.mid_m_ints
        lin("gates_ints.fc_out", width, 6 * J)
    if kind == "modalities" or cfg.ints:
        lin("out_ints", n_classes, width)
    if kind != "modalities" and cfg.ctx:
        lin("out_ctx", n_rels, 3 * J)
    return sd
