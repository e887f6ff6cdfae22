#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isamp = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for i, r in enumerate(rows[h + 1:]):
    if len(r) > isamp and r[isamp].isdigit():
        data.append((int(r[isamp]), i, r[ia]))
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "instructions", len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for s, i, src in sorted(data, reverse=True)[:n]:
    print(f"{100 * s / tot:5.1f}%  #{i:4d} {src.strip()[:120]}")
