"""Reference-style CPU training / inference step, timed by bench.py.

TEST / MEASUREMENT INFRASTRUCTURE — see oracle/__init__.py.  This is the oracle port of the
reference's own CPU path (the reference is pure Python/PyTorch and cannot travel to the GPU box):
the DENSE float64 batch the reference dataloader emits, `.float()` casts, fp32 torch CPU math
with dropout from torch's RNG, the reference loss, autograd backward and torch.optim.Adam with
the reference hyper-parameters (mlp/train.py:52-63, mlp/model.py:599-601).
"""
import time

import numpy as np
import torch

from . import losses as ol
from . import model as om


class CpuStep:
    def __init__(self, preset, n_classes=101, n_rels=15, seed=0, lr=3e-5, weight_decay=1e-5, dropout=0.3,
                 margin=0.101, lymbda=1.0):
        flags = {"modalities": ("modalities", 0, 0), "int_rels": ("midfusion", 1, 1),
                 "int_ch": ("maxtracks", 0, 0), "int_rel_ch": ("maxtracks", 1, 1)}[preset]
        self.kind = flags[0]
        self.cfg = om.default_cfg(ctx=flags[1], gates=flags[2], dropout=dropout)
        self.sd = om.init_state_dict(self.cfg, n_classes, n_rels, self.kind, seed=seed)
        for v in self.sd.values():
            v.requires_grad_(True)
        self.opt = torch.optim.Adam(list(self.sd.values()), lr=lr, weight_decay=weight_decay)
        self.n_rels, self.margin, self.lymbda = n_rels, margin, lymbda

    def _forward(self, dense, masks):
        f = dense["features"].float()                      # the reference casts slices with .float()
        B = f.shape[0]
        if self.kind == "modalities":
            return om.modalities_forward(self.sd, f.reshape(B, 1, -1), self.cfg, masks)
        if self.kind == "midfusion":
            return om.midfusion_forward(self.sd, f.reshape(B, -1, f.shape[-1]), dense["rels_mask"].reshape(B, -1, 1),
                                        self.cfg, masks)
        return om.maxtracks_forward(self.sd, f, dense.get("rels_mask"), self.cfg, masks)

    def _loss(self, o, dense):
        B = dense["features"].shape[0]
        if self.kind == "modalities":
            return ol.max_margin_ce(o["inters"], dense["labels"], dense["multilab_weights"], self.margin)
        if self.kind == "midfusion":
            return ol.multitask_max_margin(o["inters"], o["rels"], dense["labels"].reshape(B, 1, 1),
                                           dense["rels_label"].reshape(B), dense["multilab_weights"], self.margin,
                                           self.lymbda, self.n_rels)
        if self.cfg.ctx:
            return ol.margin_track_rels(o["inters"], o["rels"], dense["labels"], dense["rels_label"],
                                        dense["mem_mask"], dense["multilab_weights"], dense["gt_tracks"], self.margin,
                                        self.lymbda, self.n_rels)[0]
        return ol.margin_loss(o["inters"], dense["labels"], dense["mem_mask"], dense["multilab_weights"],
                              dense["gt_tracks"], self.margin)[0]

    def train_step(self, dense):
        o = self._forward(dense, "rng")
        loss = self._loss(o, dense)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.item())

    @torch.no_grad()
    def infer_step(self, dense):
        return self._forward(dense, None)


def time_train(preset, dense, steps=3, warmup=1, threads=None):
    """clips/s of the CPU path on dense batch `dense` (best of `steps` after `warmup`)."""
    if threads:
        torch.set_num_threads(threads)
    step = CpuStep(preset)
    B = dense["features"].shape[0]
    for _ in range(warmup):
        step.train_step(dense)
    best = float("inf")
    total = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        step.train_step(dense)
        dt = time.perf_counter() - t0
        best = min(best, dt)
        total += dt
    return {"clips_per_s_best": B / best, "clips_per_s_mean": B * steps / total, "s_per_step_best": best,
            "threads": torch.get_num_threads(), "clips": B}


def time_train_reference(preset, dense, steps=3, warmup=1, threads=None, lr=3e-5, weight_decay=1e-5):
    """The same measurement with the UNMODIFIED reference (mlp/model.py's create_model output: its model and loss
    classes, torch.optim.Adam with the reference's hyper-parameters, the loop body of mlp/train.py:57-63) — only
    where the reference tree is mounted (this build container; never the GPU box).  Same dense batch, same
    thread count, train mode with torch's own dropout RNG.  bench.py reports it as kind "reference"."""
    from . import reference_shim as rs
    if threads:
        torch.set_num_threads(threads)
    model, loss = rs.create_model(preset, 101, 15, seed=0)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)
    kind = {"modalities": "modalities", "int_rels": "midfusion"}.get(preset, "maxtracks")
    batch = {k: v for k, v in dense.items()}
    B = batch["features"].shape[0]
    if kind == "modalities":
        batch["features"] = batch["features"].reshape(B, 1, -1)
    elif kind == "midfusion":
        S1 = batch["features"].shape[2]
        batch["features"] = batch["features"].reshape(B, S1, -1)
        batch["rels_mask"] = batch["rels_mask"].reshape(B, -1, 1)
        batch["labels"] = batch["labels"].reshape(B, 1, 1).expand(B, S1, 1).contiguous()
        batch["rels_label"] = batch["rels_label"].reshape(B)

    def step():
        x = dict(batch)                       # MaxTracks re-points x['features'] at a reshaped view (model.py:272-274)
        out = model(x)
        lv = rs.run_loss(loss, out, x)
        opt.zero_grad()
        lv.backward()
        opt.step()
        return float(lv.item())

    for _ in range(warmup):
        step()
    best, total = float("inf"), 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        best, total = min(best, dt), total + dt
    return {"clips_per_s_best": B / best, "clips_per_s_mean": B * steps / total, "s_per_step_best": best,
            "threads": torch.get_num_threads(), "clips": B}
