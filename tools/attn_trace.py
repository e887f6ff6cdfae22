#!/usr/bin/env python
"""Phase timeline of attention_tc_kernel CTAs (debug aid): runs one MS-WSA layer at a stage shape of the
1 Mpx B=8 workload with sast_debug_trace armed and prints, per phase, the median / p90 clock deltas of
the CTAs' first head, plus how long each SM was busy.  Usage: attn_trace.py [stage 1..4]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("SAST_B200_LIB", os.path.join(ROOT, "sast_b200", "libsast_b200_trace.so"))
if not os.path.exists(os.environ["SAST_B200_LIB"]):          # the instrumented twin is not part of the default build
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "sast_b200", "csrc"), "trace"])
sys.path.insert(0, ROOT)
import sast_b200  # noqa: E402
from sast_b200 import _lib as L, ops  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
B, C = 8, 64 << (stage - 1)
H, W = 96 >> (stage - 1), 160 >> (stage - 1)
p0, p1 = 6, 10
NW = B * (H // p0) * (W // p1)
from sast_b200.config import backbone_config  # noqa: E402
net = sast_b200.build_recurrent_backbone(backbone_config((384, 640))).to(dev).eval()
layer = net.stages[stage - 1].att_blocks[0].att.win_attn
print('precision', layer.precision)

x = torch.randn(B, H, W, C, device=dev)
wf = torch.ones(NW, dtype=torch.uint8, device=dev)
tf = torch.ones(NW * p0 * p1, dtype=torch.uint8, device=dev)
sel = ops.Selection(ops.select_from_flags(wf, tf, B, H, W, p0, p1, L.WINDOW), B, H, W, p0, p1)
with torch.no_grad():
    for _ in range(3):
        layer.run(x, sel, L.WINDOW, False)
    torch.cuda.synchronize()
    heads = C // 32
    buf = torch.zeros(NW * heads * 16, dtype=torch.int64, device=dev)
    L.lib().sast_debug_trace(buf.data_ptr(), 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush.zero_()
    layer.run(x, sel, L.WINDOW, False)
    torch.cuda.synchronize()
    L.lib().sast_debug_trace(None, 0)
t = buf.view(-1, 16).cpu()
t = t[t[:, 10] != 0]
print(f"stage {stage}: C={C} windows={NW} traced CTAs={len(t)}")
names = ["entry->loop (setup, tmem alloc, index loads)", "TMA issue->landed", "QK^T mma->S ready", "softmax pass 1 (max)",
         "sync 1", "softmax pass 2 (exp, P store)", "fence+sync 2", "PV mma->O ready", "O ld + store", "final sync"]
d = (t[:, 1:11] - t[:, 0:10]).float()
for i, n in enumerate(names):
    col = d[:, i]
    print(f"  {n:48s} median {col.median():8.0f}  p90 {col.quantile(0.9):8.0f} clk")
for a, b, n in ((7, 14, "sync2 -> tid0 in branch"), (14, 13, "tc_fence_after"), (13, 11, "PV mma issue loop"), (11, 12, "commit"), (12, 8, "commit -> bar_o seen")):
    col = (t[:, b] - t[:, a]).float()
    print(f"  [{n:46s}] median {col.median():8.0f}  p90 {col.quantile(0.9):8.0f} clk")
tot = (t[:, 10] - t[:, 0]).float()
print(f"  first head total: median {tot.median():.0f} clk, p90 {tot.quantile(0.9):.0f}")
smid = t[:, 15]
span = []
for s in smid.unique():
    m = smid == s
    span.append((t[m, 10].max() - t[m, 0].min()).item())
    if len(span) == 1:
        order = t[m][t[m][:, 0].argsort()]
        print("  one SM's CTAs (entry, end) relative clk:", [(int(a - order[0, 0]), int(b - order[0, 0])) for a, b in zip(order[:, 0], order[:, 10])][:12])
print(f"  per-SM span (first entry -> last first-head end): median {sorted(span)[len(span) // 2]} clk; CTAs per SM ~{len(t) / len(span):.1f}")
