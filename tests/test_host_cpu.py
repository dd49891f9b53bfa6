"""Host-side logic that needs no GPU: dropout-hash mirror, packed batches, synthetic generator,
flag surface, data-parallel sharding."""
import numpy as np
import pytest
import torch


def test_dropout_mirror_bit_exact(built_lib):
    """oracle/dropout.py reproduces the kernels' counter hash bit for bit."""
    from lirec_b200 import _ext
    from oracle import dropout as od
    L = _ext.lib()
    rng = np.random.default_rng(0)
    for seed, stream, p in [(0, 1, 0.3), (123456789, 5, 0.3), (0xFFFFFFFF, 2, 0.5), (77, 3, 0.05)]:
        rows = rng.integers(0, 1 << 20, size=13)
        cols = rng.integers(0, 6144, size=17)
        m = od.keep_mask(seed, stream, rows, cols, p)
        for i, r in enumerate(rows):
            for j, c in enumerate(cols):
                assert bool(L.lirec_dropout_keep_host(seed, stream, int(r), int(c), p)) == bool(m[i, j])
    big = od.keep_mask(5, 1, np.arange(512), np.arange(2048), 0.3)
    assert abs(big.mean() - 0.7) < 5e-3
    assert abs(np.corrcoef(big[:-1].ravel(), big[1:].ravel())[0, 1]) < 5e-3


def test_flag_surface_matches_reference_defaults():
    from lirec_b200.utils.arg_pars import build_parser, opt
    d = vars(build_parser().parse_args([]))
    # a few reference defaults (utils/arg_pars.py:77,93,112,130,136,142,150-156) and quirks
    assert d["joint_dim"] == 512 and d["margin"] == 0.101 and d["tr_margin"] == 0.101
    assert d["rels_n_clips"] == 6 and d["lymbda"] == 1 and d["mid_m_ints"] == 6
    assert d["lr"] == 3e-5 and d["dropout"] == 0.3 and d["weight_decay"] == 1e-5 and d["batch_size"] == 64
    assert d["tr_sum_max_flag"] is True                      # store_false flag
    assert build_parser().parse_args(["--soft_gt", "False"]).soft_gt is True   # type=bool quirk
    assert opt.mlp_dim == 6912


@pytest.mark.parametrize("preset", ["modalities", "int_rels", "int_ch", "int_rel_ch"])
def test_synthetic_batch_structure(preset):
    from lirec_b200.mixed_utils import synthetic
    pb = synthetic.make_batch(32, seed=5, preset=preset)
    t = pb.tables
    counts = np.diff(t["cand_off"])
    assert pb.B == 32 and counts.min() >= 1 and counts.max() <= pb.n_slots
    if preset in ("int_ch", "int_rel_ch"):
        assert set(counts.tolist()) <= {2, 6, 7, 12, 13, 20}     # n^2+n for pairs GT, n^2+n-... for single GT
    else:
        assert (counts == 1).all()
    assert t["cand_rows"][:, 0].max() < pb.n_clip_ints and t["cand_rows"][:, 1:].max() < pb.n_track_ints
    assert (pb.track_bank[0] == 0).all()  # clip 0's no-track row
    if pb.has_ctx:
        n_ctx = np.diff(t["ctx_off"])
        assert n_ctx.min() >= 1 and n_ctx.max() <= pb.n_ctx_slots
        assert (t["rels_label"][n_ctx > 1] != 15).all()
        # every inverse CSR is a permutation consistent with its column
        for s in range(3):
            off, idx = t["inv_ctx_off%d" % s], t["inv_ctx_idx%d" % s]
            assert sorted(idx.tolist()) == list(range(pb.n_ctx_rows))
            for u in (0, len(off) // 2, len(off) - 2):
                assert (t["ctx_rows"][idx[off[u]:off[u + 1]], s] == u).all()
    for s in range(3):
        off, idx = t["inv_cand_off%d" % s], t["inv_cand_idx%d" % s]
        assert sorted(idx.tolist()) == list(range(pb.n_cand))
    gt = t["gt_tracks"]
    assert (gt[:, 0] == 0).all() and (gt[:, 1] < counts).all()


def test_dense_view_roundtrip():
    """unpack(packed) -> reference dense format -> pack_dense_batch -> same dense tensors."""
    from lirec_b200.mixed_utils import synthetic
    from lirec_b200.packing import pack_dense_batch
    pb = synthetic.make_batch(5, seed=2, preset="int_rel_ch")
    d = pb.to_dense(np.float32)
    assert d["features"].shape == (5, 20, 19, 6912) and d["rels_mask"].shape == (5, 20, 18)
    assert d["mem_mask"].sum() == pb.n_cand and d["rels_mask"].sum() == pb.n_ctx_rows
    pb2 = pack_dense_batch(d, "maxtracks")
    d2 = pb2.to_dense(np.float32)
    for k in ("features", "mem_mask", "rels_mask", "rels_label", "labels", "gt_tracks", "multilab_weights"):
        assert torch.equal(d[k], d2[k]), k
    # rows of an invalid candidate slot / invalid context row are all zero, like the reference padding
    mm = d["mem_mask"].bool()
    assert d["features"][~mm].abs().sum() == 0


def test_self_context_rows_are_tiled_candidate_rows():
    from lirec_b200.mixed_utils import synthetic
    pb = synthetic.make_batch(16, seed=9, preset="int_rel_ch")
    t = pb.tables
    n_ctx = np.diff(t["ctx_off"])
    none = t["rels_label"] == 15
    assert (n_ctx[none] == 1).all()
    first = t["ctx_rows"][t["ctx_off"][:-1]]
    assert (first[none] == t["cand_rows"][none]).all()


def test_shard_range_partitions():
    from lirec_b200 import dp
    for n in (1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_device_tables_are_lazy_views_of_the_arena():
    """PackedBatch.to_device exposes the integer tables as views built on first use; the dict protocol the
    rest of the code relies on (in / iter / len / keys / values / items / get) still sees every table."""
    import torch
    from lirec_b200.packing import PackedBatch, _DeviceTables
    arena = torch.arange(40, dtype=torch.int32)
    layout = {"cand_off": (0, 5, (5,)), "cand_rows": (5, 12, (4, 3)), "labels": (17, 4, (4,))}
    t = _DeviceTables(arena, layout)
    assert dict.__len__(t) == 0 and len(t) == 3 and "cand_rows" in t and "ctx_rows" not in t
    assert list(t) == list(layout) and list(t.keys()) == list(layout)
    assert t["cand_rows"].shape == (4, 3) and t["cand_rows"][1].tolist() == [8, 9, 10]
    assert t["cand_rows"] is t["cand_rows"] and dict.__len__(t) == 1          # materialised once
    assert t.get("labels").tolist() == [17, 18, 19, 20] and t.get("nope") is None
    assert [v.numel() for v in t.values()] == [5, 12, 4] and [k for k, _ in t.items()] == list(layout)
    with pytest.raises(KeyError):
        t["nope"]
    pb = PackedBatch()
    pb.tables = t
    assert pb.table_ptr("labels") == arena.data_ptr() + 4 * 17 == pb["labels"].data_ptr()


def test_install_aliases_exposes_the_reference_module_surface(monkeypatch):
    """`lirec_b200.install_aliases()` lets code written against the reference (`from utils.arg_pars import
    opt`, `import mlp.model`, `resume.int_rel_ch`) run on this package: same module objects, one `opt`,
    the reference's class and entry-point names."""
    import importlib
    import sys
    import lirec_b200
    names = ("utils", "utils.arg_pars", "utils.util_functions", "utils.model_saver", "mlp", "mlp.model", "mlp.train",
             "mlp.test", "mixed_utils", "mixed_utils.update_arg_pars", "mixed_utils.classification_dataloader",
             "mixed_utils.mixed_features", "resume", "resume.modalties", "resume.int_rels", "resume.int_ch",
             "resume.int_rel_ch")
    for n in names:                                   # restored by monkeypatch afterwards
        monkeypatch.setitem(sys.modules, n, sys.modules.get(n, None))
    lirec_b200.install_aliases()
    from utils.arg_pars import opt
    from lirec_b200.utils.arg_pars import opt as opt2
    assert opt is opt2
    model = importlib.import_module("mlp.model")
    for cls in ("Modalities", "MidFusionMultiClip", "MidFusionMultiClipMaxTracks", "GatingUnit",
                "MaxMarginCrossEntropyLoss", "MultiTaskMaxMargin", "MarginLoss", "MarginTrackRelsLoss",
                "MultiTaskCrossEntropyLoss", "create_model"):
        assert hasattr(model, cls), cls
    assert callable(importlib.import_module("mlp.train").training)
    assert callable(importlib.import_module("mlp.test").testing)
    assert callable(importlib.import_module("mixed_utils.classification_dataloader").MixedFeaturesDataset)
    assert callable(importlib.import_module("mixed_utils.update_arg_pars").update)
    for mod, fn in (("resume.modalties", "resume_modalities"), ("resume.int_rels", None), ("resume.int_ch", None),
                    ("resume.int_rel_ch", "resume_max_tracks")):
        m = importlib.import_module(mod)
        if fn is not None:
            assert callable(getattr(m, fn)), (mod, fn)


def test_model_saver_mirrors_the_reference_store(tmp_path):
    """per-metric folders, v<value>_ep<epoch>.pth.tar names, evicted checkpoints removed from
    disk, already-saved ones not rewritten, `ModelSaver(n=..., path=...)` — checked against the UNMODIFIED
    reference class when /root/reference is mounted, else against the expected listing."""
    import os
    from lirec_b200.utils.model_saver import ModelSaver
    seq = [(0, 0.50, 0.10), (1, 0.40, 0.30), (2, 0.60, 0.20), (3, 0.60, 0.05), (4, 0.10, 0.40), (5, 0.70, 0.40)]

    def drive(cls, root):
        ms = cls(n=2, path=str(root))
        listing = []
        for epoch, a, b in seq:
            val = {"total": a, "ints": b}
            if ms.check(val):
                ms.update(val, {"epoch": epoch}, epoch)
            if epoch % 2 == 1:
                ms.save()
                listing.append(sorted(os.path.relpath(os.path.join(d, f), str(root))
                                      for d, _, fs in os.walk(str(root)) for f in fs))
        return listing
    ours = drive(ModelSaver, tmp_path / "ours")
    # (update() books EVERY metric of a hit, so epoch 4 — a hit on 'ints' only — also enters 'total', evicting the
    # later of the two 0.60s; the listing below is what the unmodified reference class produces)
    assert ours[-1] == ["ints/v0.4000_ep4.pth.tar", "ints/v0.4000_ep5.pth.tar", "total/v0.6000_ep2.pth.tar",
                        "total/v0.7000_ep5.pth.tar"]
    assert all(len(l) <= 4 for l in ours)
    from oracle import reference_shim as rs
    if rs.available():
        rs.load_dataloader()
        ref_cls = rs._state["modules"]["utils.model_saver"].ModelSaver if "utils.model_saver" in rs._state["modules"] else None
        if ref_cls is None:
            import importlib.util
            spec = importlib.util.spec_from_file_location("_ref_model_saver", os.path.join(rs.REFERENCE_ROOT, "utils", "model_saver.py"))
            import sys
            saved = {k: sys.modules.get(k) for k in ("utils", "utils.util_functions")}
            sys.modules["utils"] = rs._state["modules"]["utils"]
            sys.modules["utils.util_functions"] = rs._state["modules"]["utils.util_functions"]
            try:
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
            finally:
                for k, v in saved.items():
                    if v is None:
                        sys.modules.pop(k, None)
                    else:
                        sys.modules[k] = v
            ref_cls = mod.ModelSaver
        assert drive(ref_cls, tmp_path / "ref") == ours
