"""HBM-resident feature banks + index-only batches on the GPU: the device row
gather is bit-exact, a batch staged from resident banks equals the batch streamed from the host, and
the train / eval loops run on the annotation-world dataset (--synthetic 2) in both modes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_gather_rows_is_bit_exact():
    from lirec_b200 import ops
    g = torch.Generator().manual_seed(0)
    for n_bank, dim, n in [(1000, 2816, 4097), (37, 2048, 5), (5000, 2048, 30000), (3, 8, 1)]:
        bank = torch.randn(n_bank, dim, generator=g).to(torch.bfloat16).cuda()
        idx = torch.randint(n_bank, (n,), generator=g, dtype=torch.int32).cuda()
        out = ops.gather_rows(bank, idx)
        assert torch.equal(out, bank[idx.long()])
    # strided bank view and out-of-range index -> zero row
    wide = torch.randn(64, 4096, generator=g).to(torch.bfloat16).cuda()
    view = wide[:, 1024:3072]
    idx = torch.tensor([3, -1, 63, 64, 0], dtype=torch.int32).cuda()
    out = ops.gather_rows(view, idx)
    ref = view[idx.clamp(0, 63).long()].clone()
    ref[1] = 0
    ref[3] = 0
    assert torch.equal(out, ref)
    with pytest.raises(RuntimeError):
        ops.gather_rows(wide[:, :12], idx)                         # dim not a multiple of 8


def _world_dataset(opt, mode="train", **kw):
    from lirec_b200.mixed_utils import indexed_dataset as ids, synthetic_world as sw
    world = sw.build_world(3, n_movies=3, n_scenes=10, n_inter_names=60, n_merged=101, **kw)
    ds = ids.IndexedMixedFeaturesDataset(sw.subset(world, mode), world, mode=mode)
    ds.cache()
    ds.init_relships()
    return ds


def test_resident_stage_equals_streamed_batch(opt_preset):
    opt = opt_preset("int_rel_ch", inter_class="all", merged=True, multilab_weights=True, rels=False, soft_gt=False)
    from lirec_b200.mixed_utils import indexed_dataset as ids
    from helpers import make_model
    ds = _world_dataset(opt)
    np.random.seed(1)
    records = [ds[i] for i in range(min(24, len(ds)))]
    streamed = ids.collate_indexed(records, ds).to_device("cuda")
    banks = ids.ResidentBanks(ds, "cuda")
    host = ids.collate_indexed(records, ds, resident=True)
    assert ids.ResidentBanks.h2d_bytes(host) < streamed.host.h2d_bytes() / 8
    staged = banks.stage(host)
    torch.cuda.synchronize()
    assert torch.equal(staged.clip_bank, streamed.clip_bank) and torch.equal(staged.track_bank, streamed.track_bank)
    for k in streamed.tables:
        assert torch.equal(staged.tables[k], streamed.tables[k]), k
    model, loss, _ = make_model(seed=0, n_rels=ds.n_rels - 1)
    model.train()
    a = model(streamed, seed=9)
    b = model(staged, seed=9)
    assert torch.equal(a.ragged_inters, b.ragged_inters) and torch.equal(a.ragged_rels, b.ragged_rels)
    la, lb = loss(a, {}), loss(b, {})
    assert torch.equal(la, lb) and bool(torch.isfinite(la))


@pytest.mark.parametrize("resident", [0, 1])
def test_world_dataset_trains_and_evaluates(resident, tmp_path, opt_preset):
    opt = opt_preset("int_rel_ch", synthetic=2, resident_banks=resident, world_movies=3, world_scenes=8, epochs=1,
                     batch_size=16, num_workers=0, test=True, test_fr=1, save_model=False, save_model_often=False,
                     store_root=str(tmp_path), resume=False, resume_train=False, fused_adam=1, dp=0, lr=1e-3,
                     tr_sum_max=False, inter_class="all", merged=True, multilab_weights=True, rels=False,
                     soft_gt=False, seed=0)
    from lirec_b200.mixed_utils import classification_dataloader as cd
    import lirec_b200.mlp.model as M
    import lirec_b200.mlp.test as T
    import lirec_b200.mlp.train as TR
    train_ds = cd.MixedFeaturesDataset("train").cache().init_relships()
    val_ds = cd.MixedFeaturesDataset("val").cache().init_relships()
    assert train_ds.n_classes == 101 and len(train_ds) > 16
    torch.manual_seed(0)
    model, loss, optimizer = M.create_model(train_ds.n_classes, n_rels=len(train_ds.rels_list) - 1)
    w0 = model.state_dict()["gates_ints.fc_out.weight"].clone()
    TR.training(train_ds, model=model, loss=loss, optimizer=optimizer, name="w", val_dataset=val_ds)
    res = T.testing(val_ds, model, loss, mode="val")
    assert set(res) == {"total", "ints", "rels", "tracks", "joint"}
    assert not torch.equal(w0, model.state_dict()["gates_ints.fc_out.weight"])


@pytest.mark.parametrize("case", ["relationship_only", "no_tracks"])
def test_world_dataset_rare_configurations(case, tmp_path, opt_preset):
    """The two configurations no released preset uses, through the whole loop (dataset -> packed loader -> train step
    -> evaluation): opt.ints == 0 (context branch + relationship head alone, reference mlp/model.py:102-210,
    mlp/test.py:74-109) and a dataset without person tracks (opt.tracks off, classification_dataloader.py:587-588)."""
    common = dict(synthetic=2, resident_banks=1, world_movies=3, world_scenes=8, epochs=1, batch_size=16, num_workers=0,
                  test=True, test_fr=1, save_model=False, save_model_often=False, store_root=str(tmp_path),
                  resume=False, resume_train=False, fused_adam=1, dp=0, lr=1e-3, tr_sum_max=False, inter_class="all",
                  merged=True, multilab_weights=True, rels=False, soft_gt=False, seed=0)
    if case == "relationship_only":
        opt = opt_preset("int_rels", ints=0, gates=0, **common)
    else:
        opt = opt_preset("modalities", tracks=False, **common)
    from lirec_b200.mixed_utils import classification_dataloader as cd
    import lirec_b200.mlp.model as M
    import lirec_b200.mlp.test as T
    import lirec_b200.mlp.train as TR
    train_ds, val_ds = cd.MixedFeaturesDataset("train").cache(), cd.MixedFeaturesDataset("val").cache()
    if opt.rels_multitask:
        train_ds.init_relships(), val_ds.init_relships()
    torch.manual_seed(0)
    model, loss, optimizer = M.create_model(train_ds.n_classes, n_rels=max(len(train_ds.rels_list) - 1, 0))
    names = [k for k, _ in model.named_parameters()]
    if case == "relationship_only":
        assert all("_ctx" in k for k in names)
        probe = "out_ctx.weight"
    else:
        assert not any("tracks" in k for k in names)
        rec = train_ds[0]
        assert rec["cand_rows"].shape == (1, 3) and not rec["cand_rows"][0, 1:].any()
        probe = "out_ints.weight"
    w0 = model.state_dict()[probe].clone()
    TR.training(train_ds, model=model, loss=loss, optimizer=optimizer, name="w", val_dataset=val_ds)
    res = T.testing(val_ds, model, loss, mode="val")
    assert not torch.equal(w0, model.state_dict()[probe])
    assert all(np.isfinite(v) for v in res.values())
    if case == "relationship_only":
        assert set(res) == {"total", "ints", "rels"} and res["ints"] == 0 and res["total"] == res["rels"]
    else:
        assert set(res) == {"total", "ints"}
