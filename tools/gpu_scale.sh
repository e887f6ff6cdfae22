#!/bin/bash
# weak-scaling bench on one 8-GPU box: N = 8 (and 2), one rank per GPU, no data-path collective
mkdir -p gpurun_out
for n in 8 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/scale_$n.err | grep '^{"metric"' > gpurun_out/bench_${n}gpu.json
  echo "N=$n exit $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_${n}gpu.json | head -4
done
