#!/bin/bash
# ncu --set full on selected kernels of one eager forward (1 GPU).  usage: gpu_ncu_full.sh <regex> <skip> <count> <tag>
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s "$2" -c "$3" -o "gpurun_out/$4" -f \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > "gpurun_out/$4.log" 2>&1
echo "ncu exit $?" >> "gpurun_out/$4.log"
tail -3 "gpurun_out/$4.log"
