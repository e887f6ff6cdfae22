#!/bin/bash
# quick GPU call while iterating on one kernel: the parity tests that touch it, the headline bench, a launch list
# usage: gpu_quick.sh "<pytest -k expression>"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -k "${1:-fullsize}" 2>&1 | tail -8 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*\|"gpu_launches": [0-9]*' gpurun_out/bench_quick.log | head -6
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv 2>/dev/null | head -30
