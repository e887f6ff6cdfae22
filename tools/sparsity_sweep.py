#!/usr/bin/env python
"""BASELINE.json config 3: sparsity sweep of one MS-WSA layer (sast_layer_fwd through the C ABI).

Synthetic Bernoulli selections (fixed seed) with overall keep ratio 5..100 % (window keep = token keep =
sqrt(keep)) at the stage-1 (C=64) and stage-3 (C=256) shapes of the 1 Mpx B=8 workload.  Every token is
still read and written once (unselected tokens become norm1(x)), so the time floor at keep -> 0 is the
HBM time of 2*P*C*4 bytes; the selected share adds the compacted GEMM / attention work.
Prints one JSON line per point."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sast_b200  # noqa: E402
from sast_b200 import _lib as L  # noqa: E402
from sast_b200 import ops  # noqa: E402
from sast_b200.config import attention_config  # noqa: E402
from roofline import _time_kernel, peaks  # noqa: E402

dev = torch.device("cuda:0")
pk = peaks()
for (C, H, W, tag) in ((64, 96, 160, "stage1"), (256, 24, 40, "stage3")):
    B, p0, p1 = 8, 6, 10
    T, N = p0 * p1, H * W // (p0 * p1)
    P = B * H * W
    blk = sast_b200.SAST_block(C, attention_config((p0, p1)), first_block=True).to(dev).eval()
    layer = blk.win_attn
    x = torch.randn(B, H, W, C, device=dev)
    for precision, pname in ((L.BF16, "bf16"), (L.FP32, "fp32")):
        layer.precision = precision
        layer._pack_key = None
        for keep in (0.05, 0.10, 0.25, 0.50, 0.75, 1.00):
            g = torch.Generator(device="cpu").manual_seed(int(keep * 100))
            rho = keep ** 0.5
            wf = (torch.rand(B * N, generator=g) < rho).to(torch.uint8)
            tf = (torch.rand(B * N * T, generator=g) < rho).to(torch.uint8)
            if keep == 1.0:
                wf[:] = 1
                tf[:] = 1
            sel = ops.Selection(ops.select_from_flags(wf.to(dev), tf.to(dev), B, H, W, p0, p1, L.WINDOW), B, H, W, p0, p1)
            S = int(sel.counts[1])

            # one layer call captured in a CUDA graph: the replay is timed, not the Python dispatch
            with torch.no_grad():
                for _ in range(2):
                    layer.run(x, sel, L.WINDOW, False)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph), torch.no_grad():
                layer.run(x, sel, L.WINDOW, False)
            t = _time_kernel(graph.replay, dev, iters=10)
            min_bytes = 2.0 * P * C * 4
            print(json.dumps({"shape": tag, "C": C, "tokens": P, "precision": pname, "keep_target": keep,
                              "keep_actual": S / P, "selected": S, "us_per_layer": t * 1e6,
                              "tokens_per_s": P / t, "min_bytes": min_bytes,
                              "hbm_frac_of_min_traffic": min_bytes / t / 1e9 / pk["hbm_gbs"]}), flush=True)
