"""Live roofline measurements for bench.py: individual kernels of the SAST block timed with CUDA
events on the launching stream (after warm-up, synchronised on both sides, L2 flushed between
launches), against the driver-measured peaks in MEASURED_PEAKS.json.

Algorithmic work per launch (DESIGN.md section "Roofline"):
  gather / scatter : (S*C read + S*C written) * 4 B            -> HBM bound
  tcgen05 GEMM     : 2*M*N*K FLOP                              -> tensor bound
"""
from __future__ import annotations

import json
import os

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
FALLBACK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "MEASURED_PEAKS.json (burst)"}
    return dict(FALLBACK, source="fallback (B200_PROFILING.md)")


def _time_kernel(fn, device, iters=20, flush_mb=256):
    """Average device time of fn() with an L2 flush (write of a buffer larger than L2) before each launch."""
    flush = torch.empty(flush_mb * 1024 * 1024, dtype=torch.uint8, device=device)
    for _ in range(3):
        fn()
    torch.cuda.synchronize(device)
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize(device)
        total += e0.elapsed_time(e1)
    return total / iters * 1e-3


def roofline_block(net, workload, args, device):
    from sast_b200 import _lib as L
    from sast_b200 import ops

    pk = peaks()
    B, (Hin, Win) = workload["batch"], workload["res"]
    H, W, C = Hin // 4, Win // 4, 64                      # stage 1: the largest map
    mult = 32 * workload["split"]
    p0, p1 = Hin // mult, Win // mult
    T, N = p0 * p1, H * W // (p0 * p1)
    P = B * H * W
    out = {}

    # ---- gather / scatter of all tokens of stage 1 (keep ratio 1.0), HBM bound ----
    x = torch.randn(B, H, W, C, device=device)
    wf = torch.ones(B * N, dtype=torch.uint8, device=device)
    tf = torch.ones(B * N * T, dtype=torch.uint8, device=device)
    sel = ops.Selection(ops.select_from_flags(wf, tf, B, H, W, p0, p1), B, H, W, p0, p1)
    rows = torch.empty(P, C, device=device)
    import ctypes as Cc
    g = L.Geom(B, H, W, C, p0, p1)
    lib = L.lib()
    st = L.stream_ptr(device)

    def gather():
        L.check(lib.sast_gather(Cc.byref(g), L.GRID, x.data_ptr(), Cc.byref(sel.struct), rows.data_ptr(), st), "gather")

    def scatter():
        L.check(lib.sast_scatter(Cc.byref(g), L.GRID, rows.data_ptr(), Cc.byref(sel.struct), x.data_ptr(), st), "scatter")

    for name, fn in (("sast_gather(grid)", gather), ("sast_scatter(grid)", scatter)):
        t = _time_kernel(fn, device)
        bytes_ = 2.0 * P * C * 4 + P * 4
        out[name] = {"bound": "hbm", "achieved": bytes_ / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": bytes_ / t / 1e9 / pk["hbm_gbs"], "traffic": None, "us": t * 1e6}

    # ---- the largest token-wise GEMM of stage 1 (GLU in: [P,64] x [320,64]^T) on tcgen05 ----
    if args.precision == "bf16":
        for (M, Nn, K, label) in ((P, 320, 64, "gemm_tc glu-in s1"), (P // 16, 1536, 256 * 2, "gemm_tc qkv s4-like")):
            A = torch.randn(M, K, device=device).to(torch.bfloat16)
            Wt = torch.randn(Nn, K, device=device).to(torch.bfloat16)
            D = torch.empty(M, Nn, device=device, dtype=torch.bfloat16)

            def gemm():
                L.check(lib.sast_gemm_bf16(A.data_ptr(), Wt.data_ptr(), 0, D.data_ptr(), 1, M, Nn, K, st), "gemm")

            t = _time_kernel(gemm, device)
            fl = 2.0 * M * Nn * K
            out[label] = {"bound": "tensor", "achieved": fl / t / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                          "frac": fl / t / 1e12 / pk["bf16_tflops"], "traffic": None, "us": t * 1e6,
                          "shape": [M, Nn, K]}
    head = dict(out["sast_gather(grid)"])
    head["kernel"] = "rows_copy_kernel<GATHER> (sast_gather, grid flavour, stage-1 1 Mpx B=8, keep 1.0)"
    head["peak_source"] = pk["source"]
    head["others"] = {k: v for k, v in out.items() if k != "sast_gather(grid)"}
    return head
