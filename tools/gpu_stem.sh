mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_neighbors.py -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -k "stem" 2>&1 | tail -30 | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv | head -30
