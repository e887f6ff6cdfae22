#!/usr/bin/env python
"""Phase timeline of stem_bits_kernel CTAs (debug aid): runs the 1 Mpx B=8 stem from 1-bit packed input with
sast_debug_trace(which=6) armed (trace build) and prints the clock deltas of each CTA's FOURTH tile."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("SAST_B200_LIB", os.path.join(ROOT, "sast_b200", "libsast_b200_trace.so"))
sys.path.insert(0, ROOT)
import sast_b200  # noqa: E402
from sast_b200 import _lib as L, ops  # noqa: E402

dev = torch.device("cuda:0")
density = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
x = (torch.rand(8, 20, 384, 640) < density).to(torch.uint8).to(dev)
pk = sast_b200.pack_events(x, 1)
w = ops.pack_stem_weight_bits(torch.randn(64, 20, 7, 7, device=dev) * 0.05)
g, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
for _ in range(3):
    ops.stem_bits_fwd(pk.data, 1, 640, w, g, b, 1e-5)
torch.cuda.synchronize()
buf = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
L.lib().sast_debug_trace(buf.data_ptr(), 6)
ops.stem_bits_fwd(pk.data, 1, 640, w, g, b, 1e-5)
torch.cuda.synchronize()
L.lib().sast_debug_trace(None, 0)
t = buf.view(-1, 32).cpu()
t = t[t[:, 22] != 0]
print(f"density {density}: traced CTAs {len(t)}")


def show(name, a, b):
    d = (t[:, b] - t[:, a]).float()
    print(f"  {name:46s} median {d.median():8.0f}  p90 {d.quantile(0.9):8.0f} clk")


show("producer: 7 kernel rows", 0, 5)
show("producer: barrier with the stager", 5, 6)
show("producer: whole tile", 0, 6)
show("mma: wait for the accumulator", 8, 9)
show("mma: 7 kernel rows x 10 K steps", 9, 10)
for ky in range(4):
    show(f"mma: ky {ky}: wait for the producers", 24 + 2 * ky, 25 + 2 * ky)
    if ky < 3:
        show(f"mma: ky {ky}: issue 10 + commit", 25 + 2 * ky, 26 + 2 * ky)
show("stager: stage the next tile", 16, 17)
show("epilogue: wait for the tile", 12, 13)
show("epilogue: LayerNorm + stores", 13, 14)
show("kernel: entry -> set-up done", 20, 21)
show("kernel: total", 20, 22)
