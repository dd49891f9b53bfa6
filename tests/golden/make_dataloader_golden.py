"""Generate tests/golden/dataloader_*.npz from the UNMODIFIED reference dataloader (run in the build
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_dataloader_golden.py [preset alias ...]

A small synthetic annotation world (lirec_b200/mixed_utils/synthetic_world.py, reduced feature dims so
the fixtures stay small) is built out of the reference's own AnnotatedInter / Relationship classes and
handed to the reference's MixedFeaturesDataset: its __init__, cache(), cache_relationships(),
cache_None_rels(), init_relships() and __getitem__ run unmodified (only the file loaders the
constructor calls are replaced, oracle/reference_shim.py:reference_dataset).  Every item of the split
is stored; tests/test_dataloader_cpu.py rebuilds the same world with the product's classes and
requires lirec_b200's index-only dataset to reproduce every array bit-exactly, including the
train-mode context subsampling driven by the global numpy RNG.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
ONLY = sys.argv[1:]            # optional: preset aliases to (re)generate, e.g. `notracks`; default all
sys.argv = sys.argv[:1]

from oracle import reference_shim as rs  # noqa: E402
from lirec_b200.mixed_utils import synthetic_world as sw  # noqa: E402

DIMS = dict(text_dim=4, visual_dim=6, track_dim=6)
WORLD = {0: dict(DIMS), 1: dict(DIMS, n_scenes=14, n_chars=5, p_rel_node=0.6)}     # world seed -> build_world kwargs
# (preset, mode, rels_n_clips, world seed, numpy seed, extra opt flags)
CASES = [
    ("int_rel_ch", "train", 18, 0, 5, {}), ("int_rel_ch", "train", 3, 1, 6, {}),
    ("int_rel_ch", "test", 18, 0, 5, {}), ("int_rel_ch", "test", 3, 1, 6, {}),
    ("int_ch", "train", 18, 0, 5, {}), ("int_ch", "test", 18, 1, 6, {}),
    ("int_rels", "train", 18, 0, 5, {}), ("int_rels", "train", 3, 1, 6, {}),
    ("int_rels", "test", 18, 0, 5, {}), ("int_rels", "test", 3, 1, 6, {}),
    ("modalities", "train", 18, 0, 5, dict(soft_gt=True)), ("modalities", "test", 18, 1, 6, dict(soft_gt=True)),
    # opt.tracks off (classification_dataloader.py:75-77, 587-588): one item per interaction in train mode, the
    # clip's text|visual vector alone; stored as preset "notracks"
    ("modalities", "train", 18, 0, 5, dict(tracks=False, soft_gt=True), "notracks"),
    ("modalities", "test", 18, 1, 6, dict(tracks=False, soft_gt=True), "notracks"),
]


def case_name(preset, mode, n_clips, wseed):
    return "dataloader_%s_%s_s%d_w%d.npz" % (preset, mode, n_clips, wseed)


def main():
    opt, _ = rs.load_dataloader()
    uf = rs._state["util_functions"]
    for case in CASES:
        preset, mode, n_clips, wseed, seed, extra = case[:6]
        alias = case[6] if len(case) > 6 else preset
        if ONLY and alias not in ONLY:
            continue
        w = sw.build_world(wseed, inter_cls=uf.AnnotatedInter, rel_cls=uf.Relationship, **WORLD[wseed])
        opt.soft_gt = False
        ds = rs.reference_dataset(sw.subset(w, mode), w, mode, preset, rels_n_clips=n_clips, **extra)
        np.random.seed(seed)
        items = [ds[i] for i in range(len(ds))]
        out = {"n_items": len(items), "numpy_seed": seed, "world_seed": wseed, "rels_n_clips": n_clips,
               "world_kwargs": json.dumps(WORLD[wseed]), "n_classes": ds.n_classes, "n_rels": ds.n_rels, "keys": np.array(sorted(items[0].keys()))}
        for k in items[0]:
            arr = np.stack([np.asarray(it[k]) for it in items])
            if k == "features":
                assert np.array_equal(arr.astype(np.float16).astype(np.float64), arr)
                arr = arr.astype(np.float16)
            out["item_" + k] = arr
        path = os.path.join(HERE, case_name(alias, mode, n_clips, wseed))
        np.savez_compressed(path, **out)
        print("%-48s %3d items  %6.1f KB" % (os.path.basename(path), len(items), os.path.getsize(path) / 1e3))
    opt.soft_gt = False


if __name__ == "__main__":
    main()
