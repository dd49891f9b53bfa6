"""BASELINE config 5 — long-clip stress: sweep sequence lengths / candidate counts and report the achieved
HBM bandwidth of the ragged pooling and pair-scoring kernels plus the full int_rel_ch step (run on a B200):
    python tools/stress_sweep.py > profiles/r01_stress_sweep.txt
Bytes are ALGORITHMIC: every input element read once, every output written once."""
import contextlib
import io
import json
import os
import sys

sys.argv = sys.argv[:1]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lirec_b200 import _ext, ops  # noqa: E402
from lirec_b200.mixed_utils import synthetic  # noqa: E402
from lirec_b200.utils.arg_pars import opt  # noqa: E402

_ext.require_device()
PEAK = 6544.7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(iters):
        FLUSH.zero_()                                   # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def report(name, cfg, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print("%-22s %-42s %8.1f MB %8.3f ms %8.1f GB/s  %4.1f%% of %.0f" % (name, cfg, nbytes / 1e6, ms, gbs,
                                                                        100 * gbs / PEAK, PEAK))
    sys.stdout.flush()


def sweep_seg_max():
    rng = np.random.default_rng(0)
    for scale in (1, 2, 4):
        for dim, mean_len in ((2048, 32 * scale), (768, 64 * scale), (2048, 128 * scale)):
            nseg = 2048
            lens = rng.integers(0, 2 * mean_len + 1, size=nseg)
            off = np.zeros(nseg + 1, dtype=np.int32)
            np.cumsum(lens, out=off[1:])
            x = torch.randn(int(off[-1]), dim, device="cuda")
            offd = torch.from_numpy(off).cuda()
            out = torch.empty(nseg, dim, dtype=torch.bfloat16, device="cuda")
            ms = timeit(lambda: ops.seg_reduce(x, offd, "max", out_bf16=out))
            report("seg_max", "x%d dim=%d mean_len=%d nseg=%d" % (scale, dim, mean_len, nseg),
                   x.numel() * 4 + out.numel() * 2, ms)


def sweep_softmax_pool():
    """The softmax-weighted member of the segmented-reduction family (parity unpinned: the reference has none)."""
    rng = np.random.default_rng(2)
    for dim, mean_len in ((2048, 32), (768, 64), (2048, 128), (2048, 512)):
        nseg = 2048 if mean_len < 512 else 512
        lens = rng.integers(0, 2 * mean_len + 1, size=nseg)
        off = np.zeros(nseg + 1, dtype=np.int32)
        np.cumsum(lens, out=off[1:])
        x = torch.randn(int(off[-1]), dim, device="cuda")
        offd = torch.from_numpy(off).cuda()
        ms = timeit(lambda: ops.seg_softmax_pool(x, offd, beta=1.5))
        report("seg_softmax fwd", "dim=%d mean_len=%d nseg=%d" % (dim, mean_len, nseg), x.numel() * 4 + 2 * nseg * dim * 4, ms)
        out, lse = ops.seg_softmax_pool(x, offd, beta=1.5)
        dy = torch.randn_like(out)
        ms = timeit(lambda: ops.seg_softmax_pool_bwd(x, offd, 1.5, None, out, lse, dy))
        report("seg_softmax bwd", "dim=%d mean_len=%d nseg=%d" % (dim, mean_len, nseg), 2 * x.numel() * 4 + 3 * nseg * dim * 4, ms)
        # attention pooling: one score per row, shared by the columns
        sc = torch.randn(x.shape[0], device="cuda")
        ms = timeit(lambda: ops.seg_softmax_pool(x, offd, beta=1.5, scores=sc))
        report("seg_softmax fwd (row scores)", "dim=%d mean_len=%d nseg=%d" % (dim, mean_len, nseg),
               x.numel() * 4 + sc.numel() * 4 + nseg * dim * 4, ms)
        out, lse = ops.seg_softmax_pool(x, offd, beta=1.5, scores=sc)
        ms = timeit(lambda: ops.seg_softmax_pool_bwd(x, offd, 1.5, sc, out, lse, dy))
        report("seg_softmax bwd (row scores)", "dim=%d mean_len=%d nseg=%d" % (dim, mean_len, nseg),
               2 * x.numel() * 4 + 2 * sc.numel() * 4 + 2 * nseg * dim * 4, ms)


def sweep_roi():
    rng = np.random.default_rng(1)
    T, C, H, W = 64, 2048, 13, 30
    maps = torch.rand(T, C, H, W, device="cuda")
    for scale in (1, 2, 4):
        ntracks, mean_len = 64, 16 * scale
        lens = rng.integers(1, 2 * mean_len, size=ntracks)
        n = int(lens.sum())
        el = np.zeros((n, 5), dtype=np.int32)
        el[:, 0] = rng.integers(0, T, n)
        el[:, 1] = rng.integers(0, 4, n)
        el[:, 2] = rng.integers(9, H + 1, n)
        el[:, 3] = rng.integers(0, 10, n)
        el[:, 4] = rng.integers(18, W + 1, n)
        off = np.zeros(ntracks + 1, dtype=np.int32)
        np.cumsum(lens, out=off[1:])
        area = ((el[:, 2] - el[:, 1]) * (el[:, 4] - el[:, 3])).sum()
        eld, offd = torch.from_numpy(el).cuda(), torch.from_numpy(off).cuda()
        out = torch.empty(ntracks, C, dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: ops.roi_max_pool(maps, eld, offd, out_bf16=out))
        report("roi_max_pool (tracks)", "x%d tracks=%d mean_len=%d box~%dx%d" % (scale, ntracks, mean_len, 9, 18),
               int(area) * C * 4 + out.numel() * 2, ms)
        ms = timeit(lambda: ops.roi_max_pool(maps, eld, offd, out_bf16=out, two_stage=False))
        report("  single-pass form", "x%d" % scale, int(area) * C * 4 + out.numel() * 2, ms)
    for nfr in (8, 32, 64):
        el = np.zeros((nfr, 5), dtype=np.int32)
        el[:, 0], el[:, 2], el[:, 4] = np.arange(nfr), H, W
        eld = torch.from_numpy(el).cuda()
        offd = torch.tensor([0, nfr], dtype=torch.int32).cuda()
        out = torch.empty(1, C, dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: ops.roi_max_pool(maps, eld, offd, out_bf16=out))
        report("roi_max_pool (clip)", "frames=%d full %dx%d maps" % (nfr, H, W), nfr * C * H * W * 4, ms)
    # the cache() shape of the same pooling (visual_features.py:60-69 for every clip of a movie): many clips per
    # launch over distinct frames — a single 32-frame clip is 102 MB = 16 us at the HBM peak, less than the fixed
    # cost of timing one launch with a flushed L2 (~20 us on every kernel of this sweep)
    del maps
    T2 = 512
    maps2 = torch.rand(T2, C, H, W, device="cuda")
    for nclip, nfr in ((16, 32), (8, 64), (64, 8)):
        el = np.zeros((nclip * nfr, 5), dtype=np.int32)
        el[:, 0], el[:, 2], el[:, 4] = np.arange(nclip * nfr), H, W
        eld = torch.from_numpy(el).cuda()
        offd = torch.from_numpy(np.arange(nclip + 1, dtype=np.int32) * nfr).cuda()
        out = torch.empty(nclip, C, dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: ops.roi_max_pool(maps2, eld, offd, out_bf16=out))
        report("roi_max_pool (clips)", "%d clips x %d frames, full %dx%d maps" % (nclip, nfr, H, W),
               nclip * nfr * C * H * W * 4, ms)
    del maps2


def sweep_gather():
    bank = torch.randn(60000, 2816, device="cuda").to(torch.bfloat16)
    for n in (13000, 52000):
        idx = torch.randint(60000, (n,), dtype=torch.int32, device="cuda")
        out = torch.empty(n, 2816, dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: ops.gather_rows(bank, idx, out))
        report("gather_rows", "n=%d dim=2816" % n, 2 * out.numel() * 2, ms)


def sweep_step():
    for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True, rels_multi_clip=True,
                     mod_check=False, device="cuda", fused_adam=1).items():
        setattr(opt, k, v)
    import lirec_b200.mlp.model as M
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model, loss_fn, optimizer = M.create_model(101, n_rels=15)
    model.train()
    cases = [("base  T=20 S=18", lambda s: synthetic.make_batch(256, seed=s)),
             ("x4    T=80 S=72", lambda s: synthetic.stress_batch(256, seed=s))]
    for name, mk in cases:
        pbs = [mk(s).pin().to_device("cuda") for s in range(2)]
        it = [0]

        def step():
            pb = pbs[it[0] % 2]
            it[0] += 1
            lv = loss_fn(model(pb), {})
            optimizer.zero_grad()
            lv.backward()
            optimizer.step()
        ms = timeit(step, iters=12)
        pb = pbs[0]
        print("train step %-16s B=256: %6d candidate rows %7d context rows  %7.3f ms/step  %9.0f clips/s  %8.0f "
              "candidate rows/s" % (name, pb.n_cand, pb.n_ctx_rows, ms, 256 / ms * 1e3, pb.n_cand / ms * 1e3))
        sys.stdout.flush()
        # the pair-scoring loss alone
        out = model(pb)
        args = (out.ragged_inters.detach(), out.ragged_rels.detach(), pb["cand_off"], pb["labels"], pb["rels_label"],
                pb["gt_tracks"], pb.multilab, 0.101, 1.0, 15)
        ms = timeit(lambda: ops.loss_track(*args, max_slots=pb.n_slots))
        report("track_loss", "%s Ni=%d" % (name, pb.n_cand), 2 * pb.n_cand * (101 + 15) * 4, ms)


if __name__ == "__main__":
    print("# stress sweep on %s; HBM peak %.0f GB/s (MEASURED_PEAKS.json); L2 flushed between launches" % (
        torch.cuda.get_device_name(0), PEAK))
    only = [t for t in os.environ.get("LIREC_SWEEP_ONLY", "").split(",") if t]      # e.g. roi,softmax
    for name, fn in (("seg_max", sweep_seg_max), ("softmax", sweep_softmax_pool), ("roi", sweep_roi),
                     ("gather", sweep_gather), ("step", sweep_step)):
        if not only or name in only:
            fn()
