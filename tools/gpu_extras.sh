#!/bin/bash
# One gpurun call: phase traces of the three persistent kernels, the layer sparsity sweep, and the non-headline bench lines.
mkdir -p gpurun_out
{ for a in "1 1.0" "2 1.0" "1 0.05"; do timeout 200 python tools/fused_trace.py $a; done; } > gpurun_out/fused_trace.txt 2>&1
{ for a in "3 1.0" "4 1.0"; do timeout 200 python tools/fused_trace.py $a; done; } > gpurun_out/group_trace.txt 2>&1
{ timeout 200 python tools/stem_trace.py 1.0; timeout 200 python tools/stem_trace.py 0.05; } > gpurun_out/stem_trace.txt 2>&1
timeout 600 python tools/sparsity_sweep.py > gpurun_out/sparsity_sweep_layer.jsonl 2> gpurun_out/sparsity_sweep.err
: > gpurun_out/bench_input_sparsity.jsonl
for s in 0.99 0.999 0.9999; do timeout 300 python bench.py --sparsity $s --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{"metric"' >> gpurun_out/bench_input_sparsity.jsonl; done
: > gpurun_out/bench_poisson.jsonl
for s in 0.90 0.99; do timeout 300 python bench.py --input poisson --sparsity $s --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{"metric"' >> gpurun_out/bench_poisson.jsonl; done
timeout 300 python bench.py --workload gen1_b1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_gen1.json
timeout 300 python bench.py --workload 1mpx_stream --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_stream.json
timeout 300 python bench.py --pack 0 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{"metric"' > gpurun_out/bench_u8_input.json
wc -l gpurun_out/*.jsonl gpurun_out/bench_gen1.json gpurun_out/bench_stream.json gpurun_out/bench_u8_input.json
tail -3 gpurun_out/sparsity_sweep.err
