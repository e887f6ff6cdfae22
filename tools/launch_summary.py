#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration) of bench.py --no-graph: one forward, by kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
detail = len(sys.argv) > 3
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if any(k in r["Kernel Name"] for k in ("nonzero_count", "nonzero_ratio", "unpack_count", "packed1_count"))]
idx.append(len(rows))
which = min(which, len(idx) - 2)
s, e = idx[which], idx[which + 1]
agg, tot = collections.OrderedDict(), 0.0
for r in rows[s:e]:
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] == "ns":
        v /= 1000
    tot += v
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:64]
    if detail and any(p in name for p in ("sast::", "fl::", "gl::", "sn::", "sb::")):
        print(f"{v:8.1f} {r['Grid Size']:>16} {name}")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
print(f"--- one forward: {e - s} launches, {tot:.1f} us (serialised, cold cache)")
ours = sum(t for k, (n, t) in agg.items() if any(p in k for p in ("sast::", "fl::", "gl::", "sn::", "sb::")))
print(f"--- sast:: kernels {ours:.1f} us ({100 * ours / tot:.0f} %), library/torch kernels {tot - ours:.1f} us")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{t:9.1f} us {100 * t / tot:5.1f}% {n:4d}  {k}")
