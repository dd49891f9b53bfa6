"""The loops end to end on a tiny synthetic dataset: resume/int_rel_ch.py preset -> create_model ->
training (1 epoch, packed async loader, fused flat Adam) -> testing (device-side metrics) -> checkpoint
in the reference's format -> reload."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_int_rel_ch_train_eval_checkpoint(tmp_path, opt_preset, monkeypatch):
    opt = opt_preset("int_rel_ch", synthetic=1, epochs=1, batch_size=16, num_workers=0, test=True, test_fr=1,
                     save_model=True, save_model_often=False, store_root=str(tmp_path), resume=False,
                     resume_train=False, fused_adam=1, dp=0, lr=1e-3, tr_sum_max=False, rels_multi_clip=True)
    from lirec_b200.mixed_utils import classification_dataloader as cd
    monkeypatch.setattr(cd.SyntheticClipsDataset, "SIZES", {"train": 48, "val": 24, "test": 24})
    import lirec_b200.mlp.model as M
    import lirec_b200.mlp.test as T
    import lirec_b200.mlp.train as TR
    train_ds, val_ds = cd.MixedFeaturesDataset("train").cache().init_relships(), cd.MixedFeaturesDataset("val")
    torch.manual_seed(0)
    model, loss, optimizer = M.create_model(train_ds.n_classes, n_rels=len(train_ds.rels_list) - 1)
    assert isinstance(optimizer, M.FlatAdam)
    before = T.testing(val_ds, model, loss, mode="val")
    w0 = model.state_dict()["out_ints.weight"].clone()
    TR.training(train_ds, model=model, loss=loss, optimizer=optimizer, name="t", val_dataset=val_ds)
    after = T.testing(val_ds, model, loss, mode="val")
    assert set(after) == {"total", "ints", "rels", "tracks", "joint"}
    assert not torch.equal(w0, model.state_dict()["out_ints.weight"])
    assert all(0.0 <= v <= 4.0 for v in after.values())
    ckpt = torch.load(os.path.join(str(tmp_path), "0.pth.tar"), map_location="cpu", weights_only=False)
    assert set(ckpt) == {"epoch", "state_dict", "optimizer"} and len(ckpt["state_dict"]) == 38
    st = ckpt["optimizer"]["state"]
    assert len(st) == 38 and set(st[0]) == {"step", "exp_avg", "exp_avg_sq"}      # torch.optim.Adam layout
    model2, _, _ = M.create_model(train_ds.n_classes, n_rels=15)
    model2.load_state_dict(ckpt["state_dict"])
    pb = next(iter(cd.packed_loader(val_ds, 8, shuffle=False, device="cuda")))
    model.eval(), model2.eval()
    with torch.no_grad():
        assert torch.equal(model(pb).ragged_inters, model2(pb).ragged_inters)


def test_torch_adam_and_fused_adam_agree(opt_preset):
    """Drop-in optimizer (torch.optim.Adam on the flat-buffer parameters) and the fused flat Adam take the
    same step from the same state (the kernels are deterministic, so the gradients are bit-identical)."""
    from lirec_b200.mixed_utils import synthetic
    from helpers import make_model
    pb = synthetic.make_batch(8, seed=4).to_device("cuda")
    finals, grads = [], []
    for fused in (0, 1):
        opt_preset("int_rel_ch", fused_adam=fused, lr=1e-3)
        model, loss, optimizer = make_model(seed=3)
        model.train()
        lv = loss(model(pb, seed=50), {})
        optimizer.zero_grad()
        lv.backward()
        grads.append({k: p.grad.clone() for k, p in model.named_parameters()})
        optimizer.step()
        finals.append({k: v.clone() for k, v in model.state_dict().items()})
        # the bf16 shadow the next forward reads is fresh either way
        out1 = model(pb, seed=51).ragged_inters.clone()
        finals[-1]["__next_logits"] = out1
    for k in grads[0]:
        assert torch.equal(grads[0][k], grads[1][k]), k
    for k in finals[0]:
        tol = 1e-4 if k == "__next_logits" else 2e-6      # logits see the bf16 re-rounding of the weights
        assert float((finals[0][k] - finals[1][k]).abs().max()) < tol, k


def test_flat_adam_state_roundtrip_and_torch_adam_interop(opt_preset):
    """optimizer.state_dict() / load_state_dict() in torch.optim.Adam's layout (reference checkpoints,
    mlp/train.py:84-106): a FlatAdam resumed from its own state, and one resumed from a torch.optim.Adam
    state, continue exactly like the optimizer that kept running."""
    import copy
    from lirec_b200.mixed_utils import synthetic
    from helpers import make_model
    import lirec_b200.mlp.model as M
    pbs = [synthetic.make_batch(8, seed=s).to_device("cuda") for s in (4, 5, 6)]

    def run(model, loss, optimizer, pb, seed):
        lv = loss(model(pb, seed=seed), {})
        optimizer.zero_grad()
        lv.backward()
        optimizer.step()

    finals = {}
    for src in ("flat", "torch"):
        opt_preset("int_rel_ch", fused_adam=1 if src == "flat" else 0, lr=1e-3)
        model, loss, optimizer = make_model(seed=3)
        model.train()
        run(model, loss, optimizer, pbs[0], 60)
        run(model, loss, optimizer, pbs[1], 61)
        sd_m, sd_o = copy.deepcopy(model.state_dict()), copy.deepcopy(optimizer.state_dict())
        assert len(sd_o["state"]) == 38
        assert all(float(st["step"]) == 2.0 for st in sd_o["state"].values())
        run(model, loss, optimizer, pbs[2], 62)             # the run that simply keeps going
        finals[src + "_continued"] = {k: v.clone() for k, v in model.state_dict().items()}
        # resume into a fresh model + FlatAdam
        opt_preset("int_rel_ch", fused_adam=1, lr=1e-3)
        model2, loss2, optimizer2 = make_model(seed=11)
        assert isinstance(optimizer2, M.FlatAdam)
        model2.load_state_dict(sd_m)
        optimizer2.load_state_dict(sd_o)
        assert optimizer2._t == 2
        assert optimizer2.state[model2._param_list[0]]["exp_avg"].data_ptr() == optimizer2._m.data_ptr()
        model2.train()
        run(model2, loss2, optimizer2, pbs[2], 62)
        finals[src] = {k: v.clone() for k, v in model2.state_dict().items()}
        assert float(optimizer2.state_dict()["state"][0]["step"]) == 3.0
    for k in finals["flat_continued"]:
        d_flat = float((finals["flat_continued"][k] - finals["flat"][k]).abs().max())
        d_torch = float((finals["torch_continued"][k] - finals["torch"][k]).abs().max())
        assert d_flat == 0.0, (k, d_flat)
        # third step from torch.optim.Adam's state, taken by FlatAdam vs by torch.optim.Adam itself: the two
        # implementations agree to ~2e-6 per step (test above); a lost moment or step count would show as
        # ~lr = 1e-3.  (Whole trajectories are not comparable: Adam normalises, so a last-bit difference
        # in a near-zero gradient moves that element by a full lr.)
        assert d_torch < 4e-6, (k, d_torch)


@pytest.mark.parametrize("preset", ["modalities", "int_rels", "int_ch", "int_rel_ch"])
def test_native_train_step_equals_autograd_path(opt_preset, preset):
    """mlp.model.train_step (forward + loss + backward as three native calls, no autograd engine) leaves the
    same loss and bit-identical gradients as loss(model(x), x).backward() — for every model / loss pair."""
    from lirec_b200.mixed_utils import synthetic
    from helpers import make_model
    import lirec_b200.mlp.model as M
    opt_preset(preset, fused_adam=1, lr=1e-3)
    pb = synthetic.make_batch(12, seed=9, preset=preset).to_device("cuda")
    model, loss, optimizer = make_model(seed=5)
    model.train()
    lv = loss(model(pb, seed=77), {})
    optimizer.zero_grad()
    lv.backward()
    ref = {k: p.grad.clone() for k, p in model.named_parameters()}
    model._flat_grad.fill_(float("nan"))                       # the native step must rewrite every gradient
    lv2 = M.train_step(model, loss, pb, seed=77)
    assert not lv2.requires_grad and torch.equal(lv.detach(), lv2)
    for k, p in model.named_parameters():
        assert p.grad is not None and torch.equal(p.grad, ref[k]), k
    # and the loop helper takes the same step either way
    import lirec_b200.mlp.train as TR
    finals = []
    for native in (0, 1):
        o = opt_preset(preset, fused_adam=1, lr=1e-3, native_step=native)
        model, loss, optimizer = make_model(seed=5)
        model.train()
        for s in range(2):
            TR.train_step(model, loss, optimizer, pb)
        finals.append({k: v.clone() for k, v in model.state_dict().items()})
    for k in finals[0]:
        assert torch.equal(finals[0][k], finals[1][k]), k


@pytest.mark.parametrize("preset", ["int_rel_ch", "modalities"])
def test_overlapped_adam_takes_the_same_steps(opt_preset, preset):
    """--overlap_adam 1 (the single-GPU default): the gate + head parameters' Adam pass runs on a side stream behind
    backward's heads-final event, in co-resident CTAs; the encoder parameters follow after backward.  Same kernel,
    same elements: the trajectory is bit-identical to the plain optimizer step."""
    from lirec_b200 import dp
    from lirec_b200.mixed_utils import synthetic
    from helpers import make_model
    import lirec_b200.mlp.train as TR
    pbs = [synthetic.make_batch(16, seed=30 + i, preset=preset).to_device("cuda") for i in range(3)]
    finals = []
    for overlap in (False, True):
        opt_preset(preset, fused_adam=1, lr=1e-3)
        model, loss, optimizer = make_model(seed=5)
        model.train()
        fused = dp.SwitchReduceAdam.attach(model, optimizer, single_gpu=True) if overlap else None
        assert (fused is not None) == overlap
        for s in range(3):
            TR.train_step(model, loss, optimizer, pbs[s], 1, fused)
        torch.cuda.synchronize()
        finals.append(({k: v.clone() for k, v in model.state_dict().items()}, optimizer._m.clone(), optimizer._v.clone(),
                       model._flat_bf16.clone()))
        if fused is not None:
            fused.detach()
    for k in finals[0][0]:
        assert torch.equal(finals[0][0][k], finals[1][0][k]), k
    for a, b in zip(finals[0][1:], finals[1][1:]):
        assert torch.equal(a, b)
