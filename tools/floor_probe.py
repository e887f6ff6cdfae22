#!/usr/bin/env python
"""Fixed per-launch cost of the library's kernels: N back-to-back launches of tiny problems inside one CUDA
graph, time per launch.  Used to separate launch/prologue floors from per-tile work."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sast_b200 import _lib as L, ops

dev = torch.device("cuda:0")
lib = L.lib()
st_of = lambda: L.stream_ptr(dev)


def time_graph(fn, n=40, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * reps)


out = {}
for (M, N, K) in ((128, 64, 64), (128, 512, 512), (1920, 512, 512), (18944, 64, 64), (122880, 192, 64)):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = torch.randn(N, K, device=dev).to(torch.bfloat16)
    D = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out[f"gemm_bf16 {M}x{N}x{K}"] = time_graph(lambda: L.check(lib.sast_gemm_bf16(A.data_ptr(), W.data_ptr(), 0, D.data_ptr(), 1, M, N, K, st_of()), "g"))
x = torch.randn(4096, 64, device=dev)
out["layernorm 4096x64 (simt)"] = time_graph(lambda: ops.layernorm(x, None, None, 1e-5))
a = torch.randn(1024, device=dev)
out["torch add 1024 (reference floor)"] = time_graph(lambda: a.add_(1.0))
print(json.dumps(out, indent=1))
