#!/usr/bin/env python
"""Warp-stall summary of one kernel from an `ncu --set full --import-source on` report (SASS view).

  python tools/ncu_stalls.py gpurun_out/x.ncu-rep expand_bwd_t_kernel [top_n]
"""
import csv
import subprocess
import sys


def main(rep, kernel, top_n=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Address")
    data = [r for r in rows if r and r[0].startswith("0x") and len(r) >= len(hdr) - 2]
    ix = {h: i for i, h in enumerate(hdr)}
    n = ix["# Samples"]
    tot = sum(int(r[n]) for r in data) or 1
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("# %s: %d SASS instructions, %d warp samples" % (kernel, len(data), tot))
    agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
    for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print("%-28s %7d %5.1f%%" % (s, v, 100.0 * v / tot))
    print("# hottest instructions (samples, share, SASS, top stall reasons)")
    for r in sorted(data, key=lambda r: -int(r[n]))[:top_n]:
        st = sorted([(int(r[ix[s]] or 0), s) for s in stalls], reverse=True)[:2]
        print("%6s %5.1f%%  %-72s %s" % (r[n], 100.0 * int(r[n]) / tot, r[ix["Source"]].strip()[:72],
                                       " ".join("%s=%d" % (s[6:], v) for v, s in st if v)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
