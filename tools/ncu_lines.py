#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line:
usage: ncu_lines.py dump.csv [launch_index] [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
launches = []
cur = None
for r in rows:
    if r and r[0] == "Function Name":
        cur = {"name": r[1], "lines": []}
        launches.append(cur)
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r and r[0].isdigit() and r[2] == "-":
        cur["lines"].append(r)
L = launches[want]
hdr = L["hdr"]
i_s = hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
# columns repeat (all / not-issued): keep the first occurrence of each name
seen, sc = set(), []
for i, h in stall_cols:
    if h not in seen:
        seen.add(h); sc.append((i, h))
tot = sum(int(r[i_s] or 0) for r in L["lines"])
print(L["name"], "launches in file:", len(launches), "samples:", tot)
for r in sorted(L["lines"], key=lambda r: -int(r[i_s] or 0))[:topn]:
    st = sorted(((int(r[i] or 0), h) for i, h in sc), reverse=True)[:3]
    print("%5.1f%%  L%-4s %-78s %s" % (100.0 * int(r[i_s]) / tot, r[0], r[1].strip()[:78],
                                       " ".join("%s=%d%%" % (h[6:], 100 * v // max(int(r[i_s]), 1)) for v, h in st if v)))
