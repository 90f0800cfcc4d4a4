#!/usr/bin/env python3
"""Dynamic instruction mix of one kernel launch from `ncu --page source --csv` output (first kernel block in the file).
usage: ncu -i rep --page source --csv --kernel-name regex:X --launch-count 1 > f.csv ; ncu_srcmix.py f.csv [top]"""
import csv
import sys
from collections import Counter


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > iex:
            data.append((int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[ismp])))
    return data


if __name__ == "__main__":
    data = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    tot = sum(d[2] for d in data)
    mx = sorted(d[2] for d in data)[-20]
    c, cs = Counter(), Counter()
    for a, s, ex, sm in data:
        t = s.split()
        op = (t[1] if s.startswith("@") else t[0]).split(".")[0]
        c[op] += ex
        cs[op] += sm
    print("total warp instructions", tot, " static", len(data), " hot-loop trip count ~", mx)
    for op, v in c.most_common(top):
        print(f"{op:12s} {v / mx:7.1f} per hot trip   {100.0 * v / tot:5.1f}%   samples {cs[op]}")
