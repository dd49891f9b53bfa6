#!/usr/bin/env python
"""bench.py — headline benchmark of the LIReC hot path on B200.

metric  : train clips/sec, int_rel_ch forward + loss + backward (+ gradient allreduce) + Adam
workload: synthetic MovieGraphs-shaped packed batches, random-init weights
          of the reference architecture, bf16 tensor-core operands with hi/lo-split activations
          (fp32-grade products) and fp32 accumulation.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
  python bench.py --impl reference ...                      CPU arm: oracle port of the reference

One JSON line on stdout (rank 0).  `value` = device-resident throughput (inputs already in HBM),
`e2e` = the same step through the public Python API with the host->device copy of every packed
batch from pinned memory and a device->host read of the loss inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

_ARGV = sys.argv[1:]
sys.argv = sys.argv[:1]          # utils.arg_pars parses argv at import (reference behaviour)

import numpy as np  # noqa: E402
import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

# algorithmic MACs per encoder row / candidate row
E1 = 768 * 512 + 3 * 2048 * 512
E2 = 2 * 512 * 512 + 2 * 512 * 256
E = E1 + E2
G = 3072 * 3072
H_I = 3072 * 101
H_R = 1536 * 15


def algorithmic_flops(n_cand, n_ctx_rows):
    """fwd+bwd FLOPs of one int_rel_ch step over valid rows only, no credit for dedup (§8d)."""
    fwd = (n_cand + n_ctx_rows) * E + n_cand * (G + H_I + H_R)
    bwd = 2 * fwd - (n_cand + n_ctx_rows) * E1
    return 2.0 * (fwd + bwd)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="clips per GPU per step")
    ap.add_argument("--preset", default="int_rel_ch")
    ap.add_argument("--n_batches", type=int, default=4, help="distinct synthetic batches rotated per rank")
    ap.add_argument("--cpu_clips", type=int, default=64, help="clips per CPU-baseline step (bounded sample)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--dump_profile", default="", help="write the per-launch GEMM event timings to this file")
    ap.add_argument("--nccl_allreduce", action="store_true",
                    help="N > 1: use ncclAllReduce + Adam instead of the fused in-switch reduce + Adam kernel")
    ap.add_argument("--autograd_step", action="store_true",
                    help="drive the step through model()/loss()/backward()/optimizer.step() and the autograd engine "
                         "(the drop-in surface) instead of lirec_b200.mlp.train.train_step's native sequence")
    ap.add_argument("--ncu_window", action="store_true",
                    help="bracket the device-resident timed loop with cudaProfilerStart/Stop "
                         "(run under `ncu --profile-from-start off`; numbers printed under ncu are not bench values)")
    return ap.parse_args(_ARGV)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline_line(args, steps, warmup):
    """The reference's CPU path (oracle port) on a bounded sample of the same workload."""
    from lirec_b200.mixed_utils import synthetic
    from oracle import cpu_baseline as cb
    torch.set_num_threads(os.cpu_count() or 1)
    pb = synthetic.make_batch(args.cpu_clips, seed=0, preset=args.preset)
    dense = pb.to_dense(np.float64)
    r = cb.time_train(args.preset, dense, steps=steps, warmup=warmup)
    return {"value": r["clips_per_s_mean"], "unit": "clips/s", "cores": r["threads"], "kind": "port",
            "sample": "%d-clip dense float64 batch (reference dataloader format), %d timed train steps "
                      "(fwd+loss+bwd+Adam, dropout 0.3, fp32 torch CPU) after %d warm-up" % (args.cpu_clips, steps, warmup),
            "ms_per_step": 1e3 * args.cpu_clips / r["clips_per_s_mean"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    cb = cpu_baseline_line(args, steps, warmup)
    line = {"impl": "reference", "metric": "train clips/sec (int_rel_ch fwd+bwd)", "value": cb["value"],
            "unit": "clips/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "int_rel_ch train step, %d-clip dense batch per step on host CPU" % args.cpu_clips,
                       "preset": args.preset},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    from lirec_b200 import _ext, dp
    from lirec_b200.utils.arg_pars import opt
    from lirec_b200.mixed_utils import synthetic
    import torch.distributed as dist

    rank, world, local = dp.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    # resume/int_rel_ch.py preset
    for k, v in dict(tr_maximize=True, tracks=True, ints=1, ctx=1, gates=1, rels_multitask=True,
                     rels_multi_clip=True, rels_n_clips=18, mod_check=False, device="cuda", fused_adam=1).items():
        setattr(opt, k, v)
    if args.preset != "int_rel_ch":
        raise SystemExit("bench.py measures the int_rel_ch configuration")
    import contextlib
    import io
    import lirec_b200.mlp.model as M
    torch.manual_seed(opt.seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model, loss_fn, optimizer = M.create_model(101, n_rels=15)
    model.train()
    dp.broadcast_params(model._flat)
    fused = None if (world == 1 or args.nccl_allreduce) else dp.SwitchReduceAdam.attach(model, optimizer)

    # distinct synthetic batches per rank (seeded 1000*rank + i), pinned on the host
    host = [synthetic.make_batch(args.batch, seed=1000 * rank + i, preset=args.preset).pin()
            for i in range(args.n_batches)]
    resident = [h.to_device(dev) for h in host]
    torch.cuda.synchronize()
    in_bytes = sum(h.h2d_bytes() for h in host) / len(host)

    # the loop body a user runs: lirec_b200/mlp/train.py:train_step (forward + loss + backward as three
    # native calls, then gradient exchange + Adam); --autograd_step takes the reference's
    # model() / loss() / zero_grad() / backward() / step() sequence through the autograd engine instead
    import lirec_b200.mlp.train as TR
    opt.native_step = 0 if args.autograd_step else 1

    def step(pb):
        return TR.train_step(model, loss_fn, optimizer, pb, world, fused)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(resident[i % len(resident)])
    barrier()

    # ---------------- device-resident timed region ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _ext.launch_counter
    _ext.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.ncu_window:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for i in range(args.steps):
        step(resident[i % len(resident)])
    e1.record()
    barrier()
    if args.ncu_window:
        torch.cuda.cudart().cudaProfilerStop()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    prof = _ext.profile_end()
    launches = _ext.launch_counter - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clips_total = args.batch * world * args.steps
    value = clips_total / (ms_total / 1e3)

    # ---------------- end-to-end timed region (H2D of every batch + D2H of the loss) ----------------
    # The loop a user runs (lirec_b200/mlp/train.py via packed_loader): every step's packed batch is
    # copied from pinned host memory on a copy stream one step ahead of the compute stream, and every
    # step's loss is read back to pinned host memory (asynchronously; the region ends with a full sync).
    copy_stream = torch.cuda.Stream(device=dev)
    loss_host = torch.empty(args.steps, dtype=torch.float32).pin_memory()
    main = torch.cuda.current_stream()

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            pb = host[i % len(host)].to_device(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return pb, ev

    barrier()
    e0.record()
    nxt = prefetch(0)
    for i in range(args.steps):
        pb, ev = nxt
        main.wait_event(ev)
        pb.record_stream(main)
        if i + 1 < args.steps:
            nxt = prefetch(i + 1)
        lv = step(pb)
        loss_host[i:i + 1].copy_(lv.detach().reshape(1), non_blocking=True)
    e1.record()
    barrier()
    assert bool(torch.isfinite(loss_host).all()), "e2e: non-finite loss read back"
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = clips_total / (float(t.item()) / 1e3)

    # ---------------- end-to-end with HBM-resident dataset banks (index-only batches) ----------------
    # Same loop, but the pooled feature vectors of the whole (synthetic) dataset — here the union of the
    # rotating batches — were uploaded once, like the reference's dataset.cache(); every step copies only
    # the packed integer tables, multi-label bits and two row-index lists, and gathers its banks on the
    # device (lirec_gather_rows).  Reported next to `e2e`, which streams all features every step.
    from lirec_b200.mixed_utils.indexed_dataset import ResidentBanks
    banks = ResidentBanks(device=dev, clip=torch.cat([h.clip_bank for h in host]),
                          track=torch.cat([h.track_bank for h in host]))
    idx_host, c0, t0 = [], 0, 0
    for h in host:
        idx_host.append(h.without_banks(np.arange(c0, c0 + h.n_clip), np.arange(t0, t0 + h.n_track)).pin())
        c0, t0 = c0 + h.n_clip, t0 + h.n_track
    res_bytes = sum(ResidentBanks.h2d_bytes(h) for h in idx_host) / len(idx_host)

    def prefetch_res(i):
        with torch.cuda.stream(copy_stream):
            pb = banks.stage(idx_host[i % len(idx_host)])
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return pb, ev

    copy_stream.wait_stream(main)
    for i in range(2):                                  # warm the gather kernel / allocator on the copy stream
        pb, ev = prefetch_res(i)
        main.wait_event(ev)
        pb.record_stream(main)
        step(pb)
    barrier()
    e0.record()
    nxt = prefetch_res(0)
    for i in range(args.steps):
        pb, ev = nxt
        main.wait_event(ev)
        pb.record_stream(main)
        if i + 1 < args.steps:
            nxt = prefetch_res(i + 1)
        lv = step(pb)
        loss_host[i:i + 1].copy_(lv.detach().reshape(1), non_blocking=True)
    e1.record()
    barrier()
    assert bool(torch.isfinite(loss_host).all()), "e2e (resident banks): non-finite loss read back"
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_res_value = clips_total / (float(t.item()) / 1e3)

    # ---------------- inference: forward + device-side prediction arg-maxes (no_grad) ----------------
    from lirec_b200 import ops
    model.eval()
    with torch.no_grad():
        def infer(pb):
            out = model(pb)
            return ops.predict_tracks(out.ragged_inters, out.ragged_rels, pb["cand_off"], pb["labels"], pb["rels_label"],
                                      pb["gt_tracks"], 15)
        for i in range(3):
            infer(resident[i % len(resident)])
        barrier()
        e0.record()
        for i in range(args.steps):
            infer(resident[i % len(resident)])
        e1.record()
        barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    infer_value = clips_total / (float(t.item()) / 1e3)
    model.train()

    # ---------------- an HBM-bound kernel against the measured copy bandwidth: the fused Adam ----------------
    # 20 back-to-back optimizer steps (553 MB of parameter / moment traffic each, far beyond L2), CUDA events
    hbm_roof = None
    if world == 1:
        try:
            n_par = model._flat.numel()
            for _ in range(3):
                optimizer.step()
            barrier()
            e0.record()
            for _ in range(20):
                optimizer.step()
            e1.record()
            barrier()
            adam_ms = e0.elapsed_time(e1) / 20.0
            adam_bytes = n_par * (4 * 4 + 3 * 4 + 2)          # read p, g, m, v; write p, m, v + the bf16 shadow
            hbm_roof = {"bound": "hbm", "kernel": "adam_kernel", "achieved": adam_bytes / adam_ms / 1e6,
                        "unit": "GB/s", "bytes_per_launch": int(adam_bytes), "us_per_launch": 1e3 * adam_ms}
        except Exception as exc:                               # never lose the headline line over the extra one
            hbm_roof = {"error": repr(exc)[:200]}

    if rank != 0:
        return
    # ---------------- roofline of the dominant kernel (the tcgen05 GEMM) ----------------
    peaks_path = os.path.join(HERE, "MEASURED_PEAKS.json")
    peak, peak_src = 1590.0, "fallback (B200_PROFILING.md, burst)"
    hbm_peak, hbm_src = 7700.0, "fallback (nominal HBM3e)"
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            pk = json.load(f)
        peak, peak_src = float(pk["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained"
        if "hbm_gbs" in pk:
            hbm_peak, hbm_src = float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    if hbm_roof is not None and "achieved" in hbm_roof:
        hbm_roof.update(peak=hbm_peak, frac=hbm_roof["achieved"] / hbm_peak, peak_source=hbm_src)
    gemm_ms = sum(p[0] for p in prof)
    exec_flops = sum(p[1] for p in prof)
    nc = np.mean([h.n_cand for h in host])
    nx = np.mean([h.n_ctx_rows for h in host])
    alg_flops_step = algorithmic_flops(nc, nx)
    achieved = alg_flops_step * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    executed = exec_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None          # DRAM bytes of the GEMM launches of one step, from the committed ncu --set full capture
    tpath = os.path.join(HERE, "profiles", "r01_gemm_dram_traffic.json")
    if os.path.exists(tpath) and args.batch == 1024:
        with open(tpath) as f:
            traffic = int(json.load(f)["bytes_per_step"])
    roofline = {"bound": "tensor", "kernel": "lirec_gemm_tcgen05_pair_kernel", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "bytes per step over the 8 GEMM launches (profiles/r01_gemm_dram_traffic.json)",
                "peak_source": peak_src,
                "launches_per_step": len(prof) / float(args.steps), "gemm_ms_per_step": gemm_ms / args.steps,
                "executed_tflops": executed, "executed_frac": executed / peak,
                "note": "achieved = algorithmic fwd+bwd FLOPs of the step (SURVEY.md §8d, valid rows, no credit "
                        "for dedup) / summed CUDA-event duration of the GEMM launches; executed = MMA FLOPs "
                        "actually issued (layer 1 runs once per unique bank row; hi/lo split passes count 2-3x)"}
    if args.dump_profile:
        per = int(round(len(prof) / float(args.steps)))
        with open(args.dump_profile, "w") as f:
            f.write("# GEMM launches of one step (mean over %d steps): idx ms executed_GFLOP TFLOP/s tiles problems\n" % args.steps)
            for j in range(per):
                rows = prof[j::per]
                ms = float(np.mean([r[0] for r in rows]))
                fl = float(np.mean([r[1] for r in rows]))
                f.write("%d %.4f %.2f %.1f %d %d\n" % (j, ms, fl / 1e9, fl / ms / 1e9 if ms > 0 else 0, rows[0][2], rows[0][3]))
    cpu = None if args.no_cpu_baseline else cpu_baseline_line(args, steps=3, warmup=1)
    line = {"metric": "train clips/sec (int_rel_ch fwd+bwd)", "value": value, "unit": "clips/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "int_rel_ch (MidFusionMultiClipMaxTracks + MarginTrackRelsLoss + Adam) train step, "
                                   "synthetic MovieGraphs-shaped packed batches",
                       "clips_per_gpu": args.batch, "global_batch": args.batch * world,
                       "candidate_rows_per_step": float(nc), "context_rows_per_step": float(nx),
                       "parallelism": "dp%d" % world, "optimizer": "fused flat Adam",
                       "step_api": ("model()/loss()/backward()/optimizer.step() (autograd)" if args.autograd_step
                                    else "lirec_b200.mlp.train.train_step (native forward+loss+backward, no autograd)"),
                       "gradient_exchange": ("none (1 GPU)" if world == 1 else
                                             "in-switch multimem reduce fused with Adam (lirec_dp_allreduce_adam)"
                                             if fused is not None else "ncclAllReduce fp32 + Adam"),
                       "precision": "bf16 operands, hi/lo split activations, fp32 accumulate",
                       "l2": "inputs+workspace per step exceed L2 (%d distinct batches of %.0f MB rotate)" % (
                           len(host), in_bytes / 1e6)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": int(in_bytes),
                    "d2h_bytes_per_step": 4},
            "e2e_resident_banks": {"value": e2e_res_value, "unit": "clips/s", "h2d_bytes_per_step": int(res_bytes),
                                   "d2h_bytes_per_step": 4,
                                   "note": "dataset feature banks uploaded once and kept in HBM; per step only index "
                                           "tables cross PCIe and the batch banks are gathered on the device"},
            "inference": {"value": infer_value, "unit": "clips/s",
                          "what": "forward (eval, no_grad) + device-side track/class/relationship arg-maxes, inputs "
                                  "resident in HBM"},
            "gpu_launches": int(launches),
            "roofline": roofline}
    if hbm_roof is not None:
        line["roofline_hbm"] = hbm_roof
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: library chatter written to fd 1 while the job runs (NCCL prints
    # its version banner there) is routed to stderr, and the line is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
