mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_block.py tests/test_gpu_fullsize.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
for a in "3 1.0" "4 1.0"; do
  timeout 300 python tools/fused_trace.py $a > "gpurun_out/group_trace_${a// /_}.txt" 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_quick.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_quick.log | head -30
cat gpurun_out/group_trace_3_1.0.txt gpurun_out/group_trace_4_1.0.txt
tail -c 700 gpurun_out/bench_quick.log
