"""The device arg-max kernel + device meters against the counters of the reference's unmodified
Precision / RelationshipsAcc (tests/golden/eval_meters.npz)."""
import numpy as np
import pytest
import torch

from test_eval_cpu import G, check, replay

pytestmark = pytest.mark.gpu


def _device_predict(b, with_rels):
    from lirec_b200 import ops
    R = int(G["R"])
    mask = b["mask"] != 0
    counts = mask.sum(1)
    off = np.zeros(len(counts) + 1, dtype=np.int32)
    np.cumsum(counts, out=off[1:])
    dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).cuda()
    pred = ops.predict_tracks(dev(b["ints"][mask], np.float32), dev(b["rels"][mask], np.float32) if with_rels else None,
                              dev(off, np.int32), dev(b["labels"], np.int32),
                              dev(b["rels_label"][mask], np.int32) if with_rels else None, dev(b["gt"], np.int32),
                              R if with_rels else 0)
    return pred.cpu().numpy()


def test_device_meters_reproduce_the_reference_counters():
    check(*replay(_device_predict, device="cuda"))
