#!/usr/bin/env python3
"""Top CUDA source lines of a one-kernel .ncu-rep by warp-stall samples, with the long-scoreboard share.
usage: ncu_source_top.py <rep> [N]"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, data = "?", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":
        g = lambda name: float(r[hdr.index(name)].replace(",", "") or 0) if name in hdr else 0.0
        data.append((g("Warp Stall Sampling (All Samples)"), g("stall_long_sb"), g("stall_math"), g("Instructions Executed"),
                     fname, r[0], r[1].strip()))
tot = sum(d[0] for d in data) or 1
toti = sum(d[3] for d in data) or 1
print("# %s: CUDA lines by warp-stall samples (all %d); columns: %% samples, %% of them long_scoreboard, %% math throttle, %% of instructions" % (rep, tot))
for s, lsb, m, ins, f, ln, src in sorted(data, key=lambda d: -d[0])[:top]:
    print("%5.1f%%  lsb %4.0f%%  math %4.0f%%  inst %4.1f%%  %s:%s  %s" % (100 * s / tot, 100 * lsb / max(s, 1), 100 * m / max(s, 1), 100 * ins / toti, f, ln, src[:100]))
