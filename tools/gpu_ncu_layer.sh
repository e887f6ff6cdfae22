#!/bin/bash
# ncu --set full on the fused layer kernel at one stage shape.  usage: gpu_ncu_layer.sh <stage> <keep> <tag>
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:layer_fused_kernel" -s 4 -c 1 -o "gpurun_out/$3" -f \
    python tools/layer_bench.py $1 $2 3 > "gpurun_out/$3.log" 2>&1
echo "ncu exit $?" >> "gpurun_out/$3.log"
tail -3 "gpurun_out/$3.log"
