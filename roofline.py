"""Live roofline measurements for bench.py, against the driver-measured peaks in MEASURED_PEAKS.json.

`roofline` (the JSON object bench.py prints) describes the DOMINANT kernel of a 1 Mpx B=8 forward by summed device time
in the committed ncu launch list (profiles/r02_launches_*.csv): `fl::layer_fused_kernel<64>`, one whole MS-WSA layer
(gather + LN1/LN2 -> QKV -> masked attention -> proj -> GLU-MLP -> scatter-back) of stage 1 as one kernel.  By SURVEY.md
section 8(d) the fused block at keep ratio 1.0 is a dense contraction (arithmetic intensity >= 214 FLOP/B): bound =
tensor pipe.
    algorithmic FLOPs / launch = 6 S C^2 (QKV) + 4 C sum_w K_w^2 (QK^T, PV) + 2 S C^2 (proj) + 6 S C I (GLU-MLP)
    algorithmic bytes / launch = 2 P C 4 (every token read once, written once, fp32) + weights
It is timed with CUDA events around a one-kernel CUDA-graph replay on the launching stream (after warm-up, synchronised on
both sides, L2 flushed by a 256 MB write between launches).  `hbm_view` restates the same launch against the HBM roof,
`layer` repeats it at keep ratio 0.05 (where the layer IS bandwidth bound: every token is still read and written once),
`step` compares the whole forward with its algorithmic floors, and `others` carries the hot-path gather kernel of the
C >= 256 stages plus the standalone gather / scatter micro-benchmarks (`rows_copy_kernel`: the north-star's HBM target;
a micro-benchmark, the forward itself never launches it).

`traffic`: null.  ncu's per-kernel dram__bytes counts what reaches DRAM inside the kernel window; the layer's 31 MB of
output stays dirty in the 126 MB L2 past the end of the kernel, so read+write under-counts write-back by construction.
The READ side is complete and is reported as `ncu.dram_read_bytes` from the committed capture (33.1 MB against 31.5 MB of
algorithmic reads + 1 MB of selection indices: no re-reads).
"""
from __future__ import annotations

import ctypes as Cc
import json
import os

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
FALLBACK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "MEASURED_PEAKS.json (burst)"}
    return dict(FALLBACK, bf16_tflops_sustained=None, source="fallback (B200_PROFILING.md)")


def _ncu_table():
    """Figures condensed from the committed `ncu --set full` captures (profiles/ncu_kernels.json, written by
    tools/ncu_summary.py): kernel name -> {tensor_pipe_pct, dram_read_bytes, dram_write_bytes, duration_us, source}."""
    p = os.path.join(ROOT, "profiles", "ncu_kernels.json")
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


def layer_work(P, S, C, I, sum_k2):
    """(algorithmic FLOPs, algorithmic bytes) of one MS-WSA layer."""
    flops = 6.0 * S * C * C + 4.0 * C * sum_k2 + 2.0 * S * C * C + 6.0 * S * C * I
    weights = (3 * C * C + C * C + 2 * I * C + C * I) * 2.0
    return flops, 2.0 * P * C * 4 + weights


def time_layer(layer, x, sel, flavor, device, iters=20):
    """One MS-WSA layer as a CUDA-graph replay (no Python dispatch in the timed region)."""
    with torch.no_grad():
        for _ in range(2):
            layer.run(x, sel, flavor, False)
    torch.cuda.synchronize(device)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph), torch.no_grad():
        layer.run(x, sel, flavor, False)
    return _time_kernel(graph.replay, device, iters=iters)


def roofline_block(net, workload, args, device, ms_per_step=None):
    from sast_b200 import _lib as L
    from sast_b200 import ops

    pk = peaks()
    ncu = _ncu_table()
    B, (Hin, Win) = workload["batch"], workload["res"]
    H, W, C, I = Hin // 4, Win // 4, 64, 160              # stage 1: the largest map
    mult = 32 * workload["split"]
    p0, p1 = Hin // mult, Win // mult
    T, N = p0 * p1, H * W // (p0 * p1)
    P = B * H * W
    lib = L.lib()
    st = L.stream_ptr(device)

    # ---- dominant kernel: the fused stage-1 layer, keep ratio 1.0 and 0.05 ----
    layer = net.stages[0].att_blocks[0].att.win_attn
    x = torch.randn(B, H, W, C, device=device)
    layers = {}
    for keep in (1.0, 0.05):
        g = torch.Generator().manual_seed(int(keep * 100))
        rho = keep ** 0.5
        if keep == 1.0:
            wf, tf = torch.ones(B * N, dtype=torch.uint8), torch.ones(B * N * T, dtype=torch.uint8)
        else:
            wf = (torch.rand(B * N, generator=g) < rho).to(torch.uint8)
            tf = (torch.rand(B * N * T, generator=g) < rho).to(torch.uint8)
        sel = ops.Selection(ops.select_from_flags(wf.to(device), tf.to(device), B, H, W, p0, p1, L.WINDOW), B, H, W, p0, p1)
        t = time_layer(layer, x, sel, L.WINDOW, device)
        K = sel.win_K.double()
        S = int(sel.counts[1])
        fl, by = layer_work(P, S, C, I, float((K * K).sum()))
        layers[f"keep_{keep}"] = {"us": t * 1e6, "selected": S, "algorithmic_flops": fl, "algorithmic_bytes": by,
                                  "tflops": fl / t / 1e12, "tensor_frac": fl / t / 1e12 / pk["bf16_tflops"],
                                  "gbs": by / t / 1e9, "hbm_frac_of_min_traffic": by / t / 1e9 / pk["hbm_gbs"]}
    dense = layers["keep_1.0"]
    kname = "fl::layer_fused_kernel<64>" if lib.sast_layer_is_fused(P, C, I, layer.precision, 0) else "layer.cu kernel chain (6 launches)"
    nc = next((v for k, v in ncu.items() if k.startswith("layer_fused_kernel<64")), {})
    head = {"bound": "tensor", "achieved": dense["tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": dense["tensor_frac"],
            "traffic": None,
            "kernel": f"{kname}: one whole MS-WSA layer of stage 1 (1 Mpx B=8: {P} tokens, C=64, keep ratio 1.0) per launch",
            "us": dense["us"], "algorithmic_flops": dense["algorithmic_flops"], "algorithmic_bytes": dense["algorithmic_bytes"],
            "hbm_view": {"achieved": dense["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": dense["hbm_frac_of_min_traffic"]},
            "ncu": nc, "peak_source": pk["source"], "layer": layers}

    # ---- whole step against its algorithmic floors (four blocks, fp32 map traffic / dense FLOPs) ----
    if ms_per_step is not None and args.workload.startswith("1mpx"):
        hbm_us = (235.9e6 + 34.6e6) / (pk["hbm_gbs"] * 1e9) * 1e6
        tc_us = 105.2e9 / ((pk.get("bf16_tflops_sustained") or pk["bf16_tflops"]) * 1e12) * 1e6
        head["step"] = {"ms": ms_per_step, "floor_hbm_us": hbm_us, "floor_tensor_us": tc_us,
                        "frac_of_floor": max(hbm_us, tc_us) * 1e-3 / ms_per_step,
                        "note": "floors cover the four SAST blocks only (235.9 MB + 34.6 MB weights; 105.2 GFLOP dense, sustained peak)"}

    others = {}

    def hbm(label, fn, bytes_, note):
        t = _time_kernel(fn, device)
        others[label] = {"bound": "hbm", "achieved": bytes_ / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": bytes_ / t / 1e9 / pk["hbm_gbs"], "us": t * 1e6, "algorithmic_bytes": bytes_, "note": note}

    # ---- hot-path gather of the C >= 256 stages (gather_ln_kernel) through a stage-3 layer's first kernel is not separable
    # from the chain; the standalone gather / scatter below are MICRO-BENCHMARKS of the same access pattern ----
    wf = torch.ones(B * N, dtype=torch.uint8, device=device)
    tf = torch.ones(B * N * T, dtype=torch.uint8, device=device)
    sel = ops.Selection(ops.select_from_flags(wf, tf, B, H, W, p0, p1, L.GRID), B, H, W, p0, p1)
    rows = torch.empty(P, C, device=device)
    g = L.Geom(B, H, W, C, p0, p1)

    def gather():
        L.check(lib.sast_gather(Cc.byref(g), L.GRID, x.data_ptr(), Cc.byref(sel.struct), rows.data_ptr(), st), "gather")

    def scatter():
        L.check(lib.sast_scatter(Cc.byref(g), L.GRID, rows.data_ptr(), Cc.byref(sel.struct), x.data_ptr(), st), "scatter")

    micro = "micro-benchmark (rows_copy_kernel): the forward gathers / scatters inside the layer kernels"
    hbm("sast_gather(grid) [micro]", gather, 2.0 * P * C * 4 + P * 4, micro)
    hbm("sast_scatter(grid) [micro]", scatter, 2.0 * P * C * 4 + P * 4, micro)
    head["others"] = others
    return head
