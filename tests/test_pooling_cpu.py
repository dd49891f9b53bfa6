"""Host-side index arithmetic of the visual pooling path (reference visual_features.py:76-94, 108-131):
the (frame, box) element lists must reproduce, through plain numpy means, the rows the reference's
unmodified VisualFeatures emitted (tests/golden/pooling_visual.npz) — same frames, same boxes."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden", "pooling_visual.npz")


class _HostVisual:
    """VisualFeatures without a device: only the index arithmetic is exercised here."""

    def __init__(self, feats, frame2time, dims):
        from lirec_b200.visual_utils.visual_features import VisualFeatures
        self.v = VisualFeatures.__new__(VisualFeatures)
        self.v.shape, self.v.dims = tuple(feats.shape), dims
        self.v.frame2time = frame2time
        self.v.time2frame = {}
        for f in sorted(frame2time):
            self.v.time2frame.setdefault(frame2time[f], []).append(f)


def load_world():
    g = np.load(G)
    meta = json.loads(str(g["meta"]))
    feats = g["features"].astype(np.float32)
    fps = meta["frame2time_fps"]
    frame2time = {f: f // fps for f in range(feats.shape[0] * fps)}
    return g, meta, feats, frame2time, tuple(int(x) for x in g["dims"])


def numpy_rows(feats, el):
    rows = np.zeros((len(el), feats.shape[1]), dtype=np.float64)
    with np.errstate(all="ignore"):
        for i, (f, y0, y1, x0, x1) in enumerate(el):
            if f < 0:
                continue
            rows[i] = feats[f][:, y0:y1, x0:x1].reshape(feats.shape[1], -1).astype(np.float64).mean(axis=1) \
                if (y1 > y0 and x1 > x0) else np.nan
    return rows


def test_frame_and_box_arithmetic_matches_the_reference(opt_preset):
    from lirec_b200.utils.arg_pars import opt
    opt.sampling_fr = 0.0625
    g, meta, feats, frame2time, dims = load_world()
    hv = _HostVisual(feats, frame2time, dims).v
    for i, tn in enumerate(meta["time_nodes"]):
        el = hv.frame_elements(tn)
        ref = g["time_rows_%d" % i]
        assert len(el) == len(ref), (i, len(el), len(ref))
        np.testing.assert_allclose(numpy_rows(feats, el), ref, rtol=2e-6, atol=0)
    for i, tr in enumerate(meta["tracks"]):
        if not tr:
            continue
        el = hv.track_elements(tr)
        ref = g["track_rows_%d" % i]
        got = numpy_rows(feats, el)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), i
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=0)
    # the frame-index == T element is a zero row, not a skipped one
    assert hv.track_elements(meta["tracks"][5])[0, 0] == -1


GT = os.path.join(HERE, "golden", "pooling_text.npz")


def load_text_world():
    g = np.load(GT)
    meta = json.loads(str(g["meta"]))
    return g, meta, g["features"].astype(np.float32)


def test_token_ranges_match_the_reference():
    """Subtitle-line overlap -> token index list (reference text_features.py:151-168): the gathered rows equal
    the rows the reference's unmodified TextFeatures returned, a clip without dialog gives one zero row."""
    from lirec_b200.text_utils.text_features import TextFeatures, Time
    g, meta, feats = load_text_world()
    t = TextFeatures.__new__(TextFeatures)
    t.times = [Time(a, b) for a, b in meta["times"]]
    t.time_idx2token_range = meta["ranges"]
    empty = 0
    for i, tn in enumerate(meta["nodes"]):
        rng = t.tokens_range(tn)
        ref = g["rows_%d" % i]
        if not rng:
            assert ref.shape == (1, feats.shape[1]) and not ref.any()
            empty += 1
        else:
            assert np.array_equal(feats[rng], ref)
    assert empty >= 1
