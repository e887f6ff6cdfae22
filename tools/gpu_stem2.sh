python tools/stem_trace.py 1.0; python tools/stem_trace.py 0.05
bash tools/gpu_stem.sh 2>&1 | grep -v "^    .*select\|layernorm\|score_reduce\|Cat\|Fill\|finalize" | tail -22
