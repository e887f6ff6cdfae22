#!/usr/bin/env python
"""Condense an `ncu --set full` report (.ncu-rep) into a small CSV of the metrics DESIGN.md cites.
usage: ncu_summary.py <report.ncu-rep> <out.csv> [traffic.json] [kernels.json tag]
With a 4th argument the per-kernel figures roofline.py quotes (tensor-pipe %, DRAM bytes read / written, duration, issue
utilisation) are merged into that JSON (profiles/ncu_kernels.json), keyed by the short kernel name, tagged with `tag`."""
import csv
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "lts__t_bytes.sum",
]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
kn = hdr.index("Kernel Name")
traffic = {}
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel"] + [f"{m} [{units[i]}]" for m, i in cols])
    for r in rows[2:]:
        w.writerow([r[kn][:90]] + [r[i] for _, i in cols])
        def val(name):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic.setdefault(r[kn].split("(")[0].strip()[:60], []).append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
if len(sys.argv) > 3 and sys.argv[3] != "-":
    json.dump({k: v for k, v in traffic.items()}, open(sys.argv[3], "w"), indent=1)
if len(sys.argv) > 4:
    import os
    import re
    path, tag = sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else sys.argv[1]
    table = json.load(open(path)) if os.path.exists(path) else {}
    for r in rows[2:]:
        def val(name, scale=True):
            i = hdr.index(name)
            v = float(r[i].replace(",", ""))
            u = units[i].lower()
            return v * ({"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1) if scale else 1)
        name = re.sub(r"^void |sast::|fl::|\(.*$", "", r[kn]).replace("(int)", "").replace(" ", "")
        table[name] = {"duration_us": val("gpu__time_duration.sum"), "tensor_pipe_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                       "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                       "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active") if "smsp__issue_active.avg.pct_of_peak_sustained_active" in hdr else None,
                       "registers": val("launch__registers_per_thread"), "source": tag,
                       "note": "ncu --set full --clock-control none: serialised, L2 flushed before the launch; duration is not a bench value"}
    json.dump(table, open(path, "w"), indent=1)
print(open(sys.argv[2]).read()[:3000])
