#!/bin/bash
# One gpurun call: parity tests, smoke, bench (both arms + the eager-GPU reference line), launch list.
# Outputs land in gpurun_out/.   usage: gpu_round.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
if [ "$1" != "quick" ]; then
timeout 600 python bench.py --impl reference --device cuda --steps 20 --warmup 5 > gpurun_out/bench_ref_cuda.log 2>&1
echo "ref-cuda exit $?" >> gpurun_out/bench_ref_cuda.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_bench.log
fi
grep -E "passed|failed|error" gpurun_out/pytest.log | tail -5; grep -E "^FAILED|^ERROR" gpurun_out/pytest.log | head -40
tail -3 gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench.log; tail -c 600 gpurun_out/bench_ref_cuda.log 2>/dev/null
