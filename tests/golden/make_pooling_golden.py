"""Generate tests/golden/pooling_visual.npz from the UNMODIFIED reference feature readers (build
container only).

    python tests/golden/make_pooling_golden.py

`VisualFeatures.get_features_by_time` / `get_features_by_track` (visual_utils/visual_features.py:60-143)
are run through `__new__` on synthetic I3D-shaped maps (reduced channel count so the fixture stays
small), with awkward face boxes (partly outside the frame, degenerate, the frame-index == T element the
reference skips), and `np.max(..., axis=0, keepdims=True)` is applied the way MixedFeatures does
(mixed_utils/mixed_features.py:54, 104-105).  tests/test_pooling_gpu.py checks the fused GPU kernel
against these outputs; tests/test_pooling_cpu.py checks the host-side (frame, box) arithmetic.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.argv = sys.argv[:1]

from oracle import reference_shim as rs  # noqa: E402


def make_world(seed=0, T=9, C=32, H=13, W=30):
    rng = np.random.RandomState(seed)
    feats = (rng.randint(0, 64, size=(T, C, H, W)) / 8.0).astype(np.float32)
    fps = 16                                               # 16 frames per second, sampling_fr = 1/16
    frame2time = {f: f // fps for f in range(T * fps)}
    dims = (360, 852)                                      # original video height, width
    time_nodes = [dict(start=0, end=3), dict(start=2, end=2), dict(start=4, end=9), dict(start=7, end=9),
                  dict(start=8, end=8)]
    tracks = []
    for k in range(12):
        n = int(rng.randint(1, 7))
        tr = []
        for _ in range(n):
            w = float(rng.randint(20, 300))
            h = float(rng.randint(20, 300))
            x = float(rng.randint(-200, 2 * dims[1] + 100))
            y = float(rng.randint(-150, 2 * dims[0] + 100))
            tr.append(dict(frame=int(rng.randint(0, T * fps)), x=x, y=y, w=w, h=h))
        tracks.append(tr)
    tracks[3][0]["frame"] = T * fps                        # frame index == T: left as a zero row (:130-131)
    tracks[5] = [dict(frame=T * fps, x=100.0, y=80.0, w=60.0, h=60.0)]   # a track made of that element only
    tracks[7][0].update(x=5000.0, y=10.0)                  # box right of the frame: empty -> NaN
    tracks.append([])                                      # empty track
    return feats, frame2time, dims, time_nodes, tracks


def text_world(seed=1, n_lines=9, dim=24):
    rng = np.random.RandomState(seed)
    times, ranges, start, tok = [], [], 0, 0
    for i in range(n_lines):
        dur = int(rng.randint(1, 5))
        gap = int(rng.randint(0, 4))
        times.append((start + gap, start + gap + dur))
        n = int(rng.randint(2, 9))
        ranges.append(list(range(tok, tok + n)))
        tok += n
        start += gap + dur - (1 if rng.rand() < 0.3 else 0)      # some lines overlap in time
    feats = (rng.randint(-40, 41, size=(tok, dim)) / 8.0).astype(np.float32)
    nodes = [dict(start=0, end=2), dict(start=5, end=6), dict(start=3, end=40), dict(start=200, end=210),
             dict(start=times[4][0], end=times[4][0]), dict(start=times[2][1], end=times[6][0])]
    return feats, times, ranges, nodes


def text_golden(opt):
    import contextlib
    import io
    tf_mod = rs._state["modules"]["text_utils.text_features"]
    feats, times, ranges, nodes = text_world()
    opt.text_dim, opt.contextualization = feats.shape[1], "second-to-last"
    t = tf_mod.TextFeatures.__new__(tf_mod.TextFeatures)
    t.features, t.video_idx = feats, "tt"
    t.times = [tf_mod.Time(a, b) for a, b in times]
    t.time_idx2token_range = ranges
    t.dialogs = []
    out = {"features": feats, "meta": json.dumps(dict(times=times, ranges=ranges, nodes=nodes))}
    for i, tn in enumerate(nodes):
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints a line per call (:145)
            rows = t.get_features_by_time(tn)
        out["rows_%d" % i] = np.asarray(rows, dtype=np.float32)
        out["max_%d" % i] = np.max(rows, axis=0).reshape(1, -1).astype(np.float32)      # mixed_features.py:61
    path = os.path.join(HERE, "pooling_text.npz")
    np.savez_compressed(path, **out)
    opt.text_dim = 768
    print(path, "%.1f KB" % (os.path.getsize(path) / 1e3), [out["rows_%d" % i].shape for i in range(len(nodes))])


def main():
    opt, _ = rs.load_dataloader()
    mods = rs._state["modules"]
    vf_mod = mods["visual_utils.visual_features"]
    feats, frame2time, dims, time_nodes, tracks = make_world()
    opt.sampling_fr, opt.tf_crop, opt.spat_pool, opt.visual_dim = 0.0625, True, True, feats.shape[1]
    v = vf_mod.VisualFeatures.__new__(vf_mod.VisualFeatures)
    v.features, v.dims, v.frame2time = feats, dims, frame2time
    v.time2frame = {}
    for f, t in frame2time.items():
        v.time2frame.setdefault(t, []).append(f)
    out = {"features": feats.astype(np.float16), "dims": np.array(dims), "meta": json.dumps(
        dict(time_nodes=time_nodes, tracks=tracks, frame2time_fps=16, sampling_fr=0.0625))}
    assert np.array_equal(out["features"].astype(np.float32), feats)
    with np.errstate(all="ignore"):
        import warnings
        warnings.simplefilter("ignore")
        for i, tn in enumerate(time_nodes):
            rows = v.get_features_by_time(tn)
            out["time_rows_%d" % i] = rows
            out["time_max_%d" % i] = np.max(rows, axis=0, keepdims=True)
        for i, tr in enumerate(tracks):
            if len(tr) == 0:
                out["track_max_%d" % i] = np.zeros((1, feats.shape[1]))       # mixed_features.py:89-93
                continue
            rows = v.get_features_by_track(tr)
            out["track_rows_%d" % i] = rows
            out["track_max_%d" % i] = np.max(rows, axis=0, keepdims=True)
    for k in list(out):
        if k.startswith(("time_", "track_")):
            a32 = out[k].astype(np.float32)
            assert np.array_equal(a32.astype(np.float64), out[k].astype(np.float64), equal_nan=True)
            out[k] = a32
    path = os.path.join(HERE, "pooling_visual.npz")
    np.savez_compressed(path, **out)
    text_golden(opt)
    print(path, "%.1f KB" % (os.path.getsize(path) / 1e3), "nan tracks:",
          [i for i in range(len(tracks)) if np.isnan(out["track_max_%d" % i]).any()])


if __name__ == "__main__":
    main()
