#!/usr/bin/env python3
"""Static SASS statistics for one kernel of a cubin/.so: instruction mix per loop region.

Usage: sass_stats.py <lib.so> <kernel-substring>
Splits the kernel at backward branches so that loop bodies can be weighted by trip count by hand.
"""
import collections
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    s = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", s)[1:]:
        name = f.split("\n")[0]
        if pat not in name:
            continue
        ins = []
        for l in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        print("==", name, len(ins), "instructions")
        loops = []
        for a, t in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) <= a:
                loops.append((int(m.group(1), 16), a))
        bounds = sorted(set([0] + [x for lp in loops for x in (lp[0], lp[1] + 16)] + [ins[-1][0] + 16]))
        for lo, hi in zip(bounds, bounds[1:]):
            seg = [t for a, t in ins if lo <= a < hi]
            if not seg:
                continue
            tag = "LOOP" if any(l == (lo, hi - 16) for l in loops) else "    "
            mix = collections.Counter()
            for t in seg:
                op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
                base = op.split(".")[0]
                if base == "IMAD":
                    if "WIDE" in op: base = "IMAD.WIDE"
                    elif "MOV" in op: base = "IMAD.MOV"
                    elif "IADD" in op: base = "IMAD.IADD"
                    elif ".X" in op: base = "IMAD.X"
                    elif "SHL" in op: base = "IMAD.SHL"
                    elif "HI" in op: base = "IMAD.HI"
                mix[base] += 1
            print("%s 0x%05x-0x%05x %5d  %s" % (tag, lo, hi, len(seg), dict(mix.most_common(14))))


main()
