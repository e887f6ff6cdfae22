#!/usr/bin/env python
"""One MS-WSA layer (sast_layer_fwd through the C ABI) at a stage shape of the 1 Mpx B=8 workload: times it with CUDA
events (L2 flushed between launches) and is the command line ncu profiles.
Usage: layer_bench.py [stage 1..4] [keep 0..1] [iters] [precision bf16|bf16_chain|fp32]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sast_b200  # noqa: E402
from sast_b200 import _lib as L, ops  # noqa: E402
from sast_b200.config import backbone_config  # noqa: E402
from roofline import _time_kernel, peaks  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
keep = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
prec = {"bf16": L.BF16, "bf16_chain": L.BF16_CHAIN, "fp32": L.FP32}[sys.argv[4] if len(sys.argv) > 4 else "bf16"]
dev = torch.device("cuda:0")
B, C = 8, 64 << (stage - 1)
H, W = 96 >> (stage - 1), 160 >> (stage - 1)
p0, p1 = 6, 10
NW = B * (H // p0) * (W // p1)
P = B * H * W
net = sast_b200.build_recurrent_backbone(backbone_config((384, 640))).to(dev).eval()
layer = net.stages[stage - 1].att_blocks[0].att.win_attn
layer.precision = prec
x = torch.randn(B, H, W, C, device=dev)
g = torch.Generator().manual_seed(3)
rho = keep ** 0.5
wf = (torch.rand(NW, generator=g) < rho).to(torch.uint8) if keep < 1 else torch.ones(NW, dtype=torch.uint8)
tf = (torch.rand(NW * p0 * p1, generator=g) < rho).to(torch.uint8) if keep < 1 else torch.ones(NW * p0 * p1, dtype=torch.uint8)
sel = ops.Selection(ops.select_from_flags(wf.to(dev), tf.to(dev), B, H, W, p0, p1, L.WINDOW), B, H, W, p0, p1)
with torch.no_grad():
    t = _time_kernel(lambda: layer.run(x, sel, L.WINDOW, False), dev, iters=iters)
S = int(sel.counts[1])
min_bytes = 2.0 * P * C * 4
print(json.dumps({"stage": stage, "C": C, "tokens": P, "keep": S / P, "tiles": int(sel.counts[3]), "us_per_layer": t * 1e6,
                  "min_bytes": min_bytes, "hbm_frac_of_min_traffic": min_bytes / t / 1e9 / peaks()["hbm_gbs"]}))
