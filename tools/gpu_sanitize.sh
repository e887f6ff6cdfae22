#!/bin/bash
# compute-sanitizer memcheck over parity tests of the kernels added / changed this round (small shapes only: memcheck is ~50x slower)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_neighbors.py -m gpu -q -x --timeout 900 -p no:cacheprovider \
    -k "nhwc_stem and (64-64 or 100-96 or 36-160) or fused_downsample and (10-24 or 22-36) or fused_stem and 48-80" > gpurun_out/sanitize_neighbors.log 2>&1
echo "neighbors exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_neighbors.log | head -10
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_block.py tests/test_gpu_small_kernels.py -m gpu -q --timeout 900 -p no:cacheprovider -k "explicit_selection or block_golden or score_fwd or edge_cases" > gpurun_out/sanitize_block.log 2>&1
echo "block exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_block.log | head -10
grep -c "Program hit CUDA_ERROR_INVALID_HANDLE" gpurun_out/sanitize_block.log
