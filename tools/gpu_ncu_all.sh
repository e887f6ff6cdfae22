#!/bin/bash
# ncu --set full captures of the three kernels DESIGN.md quotes: fused layer (stage 1), group layer (stage 3), bits stem.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:layer_fused_kernel" -s 4 -c 1 -o gpurun_out/ncu_fused64 -f \
    python tools/layer_bench.py 1 1.0 3 > gpurun_out/ncu_fused64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:layer_group_kernel" -s 4 -c 1 -o gpurun_out/ncu_group256 -f \
    python tools/layer_bench.py 3 1.0 3 > gpurun_out/ncu_group256.log 2>&1
SAST_B200_LIB=sast_b200/libsast_b200.so timeout 600 ncu --set full --clock-control none --import-source on -k "regex:stem_bits_kernel" -s 2 -c 1 -o gpurun_out/ncu_stem_bits -f \
    python tools/stem_trace.py 0.5 > gpurun_out/ncu_stem_bits.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -n 2 gpurun_out/ncu_fused64.log
