mkdir -p gpurun_out
SAST_B200_LIB=sast_b200/libsast_b200_trace.so python tools/gemm_trace.py score 1 | tail -3
timeout 900 python -m pytest tests/test_gpu_select.py tests/test_gpu_small_kernels.py tests/test_gpu_fullsize.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider 2>&1 | tail -5 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*\|"gpu_launches": [0-9]*' gpurun_out/bench_quick.log | head -6
