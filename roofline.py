"""Live roofline measurements for bench.py: individual kernels of the SAST block timed with CUDA
events on the launching stream (after warm-up, synchronised on both sides, L2 flushed between
launches), against the driver-measured peaks in MEASURED_PEAKS.json.

`roofline` (the JSON object bench.py prints) is the DOMINANT kernel of a 1 Mpx B=8 forward by
summed device time in the committed ncu launch list (profiles/r01_launches_*.csv):
`gemm_tc_kernel<EPI_GLU>`, the MLP-in GEMM with the GLU epilogue, at its stage-1 shape
[S=122880, 2I=320, K=64].  With K = 64 its arithmetic intensity is 91 FLOP/B, far below the
219 FLOP/B ridge of this GPU: it is HBM bound and is reported as such:
    algorithmic bytes / launch = S*K*2 (A, bf16) + S*I*2 (GLU output, bf16) + 2I*K*2 (W)
`others` carries the north-star's named targets (gather / scatter, HBM) and two more GEMM shapes
(tensor figures: 2*M*N*K FLOP / launch).  `traffic` is dram__bytes_read+write of the same kernel
from the committed `ncu --set full` capture (profiles/), per launch.
"""
from __future__ import annotations

import ctypes as Cc
import json
import os

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
FALLBACK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
# dram__bytes_read.sum + dram__bytes_write.sum per launch, from profiles/r01_ncu_full_summary.csv (bytes)
NCU_TRAFFIC = {"gemm_tc glu s1": None, "sast_gather(grid)": None}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "MEASURED_PEAKS.json (burst)"}
    return dict(FALLBACK, source="fallback (B200_PROFILING.md)")


def _traffic_table():
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return {}
    return {}


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
    traffic = _traffic_table()
    B, (Hin, Win) = workload["batch"], workload["res"]
    H, W, C, I = Hin // 4, Win // 4, 64, 160              # stage 1: the largest map
    mult = 32 * workload["split"]
    p0, p1 = Hin // mult, Win // mult
    T, N = p0 * p1, H * W // (p0 * p1)
    P = B * H * W
    lib = L.lib()
    st = L.stream_ptr(device)
    out = {}

    def hbm(label, fn, bytes_):
        t = _time_kernel(fn, device)
        out[label] = {"bound": "hbm", "achieved": bytes_ / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                      "frac": bytes_ / t / 1e9 / pk["hbm_gbs"], "traffic": traffic.get(label), "us": t * 1e6,
                      "algorithmic_bytes": bytes_}

    # ---- dominant kernel: GLU GEMM at the stage-1 shape ----
    if args.precision == "bf16":
        A = torch.randn(P, C, device=device).to(torch.bfloat16)
        Wt = (torch.randn(2 * I, C, device=device) / 8).to(torch.bfloat16)
        bias = torch.randn(2 * I, device=device)
        D = torch.empty(P, I, device=device, dtype=torch.bfloat16)

        def glu():
            L.check(lib.sast_gemm_bf16_glu(A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), D.data_ptr(), P, 2 * I, C, st), "glu")

        hbm("gemm_tc glu s1", glu, float(P * C * 2 + P * I * 2 + 2 * I * C * 2))
        out["gemm_tc glu s1"]["shape"] = [P, 2 * I, C]
        out["gemm_tc glu s1"]["tflops"] = 2.0 * P * 2 * I * C / (out["gemm_tc glu s1"]["us"] * 1e-6) / 1e12

    # ---- gather / scatter of all tokens of stage 1 (keep ratio 1.0) ----
    x = torch.randn(B, H, W, C, device=device)
    wf = torch.ones(B * N, dtype=torch.uint8, device=device)
    tf = torch.ones(B * N * T, dtype=torch.uint8, device=device)
    sel = ops.Selection(ops.select_from_flags(wf, tf, B, H, W, p0, p1, L.GRID), B, H, W, p0, p1)
    rows = torch.empty(P, C, device=device)
    g = L.Geom(B, H, W, C, p0, p1)

    def gather():
        L.check(lib.sast_gather(Cc.byref(g), L.GRID, x.data_ptr(), Cc.byref(sel.struct), rows.data_ptr(), st), "gather")

    def scatter():
        L.check(lib.sast_scatter(Cc.byref(g), L.GRID, rows.data_ptr(), Cc.byref(sel.struct), x.data_ptr(), st), "scatter")

    hbm("sast_gather(grid)", gather, 2.0 * P * C * 4 + P * 4)
    hbm("sast_scatter(grid)", scatter, 2.0 * P * C * 4 + P * 4)

    # ---- two more GEMM shapes, tensor figures ----
    if args.precision == "bf16":
        for (M, Nn, K, label) in ((P, 192, 64, "gemm_tc qkv s1"), (P // 16, 1536, 512, "gemm_tc qkv s3-like")):
            A2 = torch.randn(M, K, device=device).to(torch.bfloat16)
            W2 = torch.randn(Nn, K, device=device).to(torch.bfloat16)
            D2 = torch.empty(M, Nn, device=device, dtype=torch.bfloat16)

            def gemm():
                L.check(lib.sast_gemm_bf16(A2.data_ptr(), W2.data_ptr(), 0, D2.data_ptr(), 1, M, Nn, K, st), "gemm")

            t = _time_kernel(gemm, device)
            fl = 2.0 * M * Nn * K
            out[label] = {"bound": "tensor", "achieved": fl / t / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                          "frac": fl / t / 1e12 / pk["bf16_tflops"], "traffic": None, "us": t * 1e6, "shape": [M, Nn, K]}

    key = "gemm_tc glu s1" if "gemm_tc glu s1" in out else "sast_gather(grid)"
    head = dict(out[key])
    head["kernel"] = ("gemm_tc_kernel<EPI_GLU> (MLP-in GEMM + GLU epilogue, stage-1 1 Mpx B=8, keep 1.0)" if key.startswith("gemm")
                      else "rows_copy_kernel<GATHER> (sast_gather, grid flavour, stage-1 1 Mpx B=8, keep 1.0)")
    head["peak_source"] = pk["source"]
    head["others"] = {k: v for k, v in out.items() if k != key}
    return head
