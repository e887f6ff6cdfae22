#!/bin/bash
# quick GPU call: selection + block tests, full-size parity tests, fused-kernel phase traces, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_select.py tests/test_gpu_block.py tests/test_gpu_fullsize.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
for a in "1 1.0" "2 1.0" "1 0.05"; do
  timeout 300 python tools/fused_trace.py $a > "gpurun_out/fused_trace_${a// /_}.txt" 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_quick.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_quick.log | head -30
cat gpurun_out/fused_trace_1_1.0.txt gpurun_out/fused_trace_2_1.0.txt gpurun_out/fused_trace_1_0.05.txt
tail -c 700 gpurun_out/bench_quick.log
