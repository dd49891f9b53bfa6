#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run in the build container; no GPU needed).

  python tools/ncu_extract.py launches gpurun_out/launches.csv [steps]   # per-launch device time of one step
  python tools/ncu_extract.py full gpurun_out/x.ncu-rep                  # key metrics of a --set full capture
"""
import csv
import subprocess
import sys

FULL = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_%"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("sm__cycles_elapsed.max.per_second", "sm_clock"),
]


def launches(path, steps=2):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(x["Kernel Name"], float(x["Metric Value"].replace(",", "")), x["Grid Size"]) for x in csv.DictReader(lines)]
    n = len(rows) // steps
    rows = rows[-n:]
    tot = sum(r[1] for r in rows)
    print("# one train step = %d launches; per-launch device time (ncu: cold cache, serialised -> compare shares)" % n)
    for i, r in enumerate(rows):
        print("%-3d %-100s %9.1f us %5.1f%%  grid %s" % (i, r[0][:100], r[1] / 1e3, 100 * r[1] / tot, r[2]))
    print("total %.1f us" % (tot / 1e3))
    agg = {}
    for r in rows:
        k = r[0].split("(")[0]
        agg[k] = agg.get(k, 0.0) + r[1]
    print("# by kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print("%-80s %9.1f us %5.1f%%" % (k[:80], v / 1e3, 100 * v / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    kn = idx["Kernel Name"]
    for j, d in enumerate(data):
        print("launch %d  %s" % (j, d[kn][:110]))
        for m, label in FULL:
            if m in idx:
                print("    %-22s %14s %s" % (label, d[idx[m]], units[idx[m]]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 2)
    else:
        full(sys.argv[2])
