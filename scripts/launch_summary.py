#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui].strip(), 1)  # ncu prints ns/us/ms depending on the value
    name = re.sub(r"\(.*", "", r[ki]).replace("qpzk::", "").strip()
    tot[name] += v * scale
    cnt[name] += 1
allt = sum(tot.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# gpu__time_duration.sum per kernel (ns), cold-cache/serialised under ncu: compare SHARES, not absolutes")
print("%-44s %8s %14s %7s" % ("kernel", "launches", "total ns", "share"))
for k, v in tot.most_common():
    print("%-44s %8d %14.0f %6.1f%%" % (k, cnt[k], v, 100 * v / allt))
