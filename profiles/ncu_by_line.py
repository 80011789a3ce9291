#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` output by CUDA source line.
usage: ncu_by_line.py file.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
cur = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ie, ns = hdr.index("Instructions Executed"), hdr.index("# Samples")
        s0, s1 = hdr.index("stall_barrier"), hdr.index("stall_wait")
        continue
    if hdr is None:
        continue
    try:
        ln = int(r[0])
        n, s = int(r[ie]), int(r[ns])
    except ValueError:
        continue
    key = (cur, ln)
    src[key] = r[1].strip()[:84]
    agg[key][0] += n
    agg[key][1] += s
    for i in range(s0, s1 + 1):
        try:
            v = int(r[i])
        except ValueError:
            continue
        if v:
            agg[key][2][hdr[i][6:]] += v
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
allst = collections.Counter()
for v in agg.values():
    allst.update(v[2])
print(f"warp instructions {tot}, samples {tots}")
print("stall mix:", ", ".join(f"{k}:{100 * c / max(tots, 1):.1f}%" for k, c in allst.most_common(8)))
for key, v in sorted(agg.items(), key=lambda kv: -(kv[1][0] / tot + kv[1][1] / max(tots, 1)))[:top]:
    st = ", ".join(f"{k}:{c}" for k, c in v[2].most_common(3))
    print(f"{key[0][:20]:20s}:{key[1]:4d} inst {100 * v[0] / tot:5.1f}% smp {100 * v[1] / max(tots, 1):5.1f}% [{st}] {src[key]}")
