#!/usr/bin/env python
"""Phase timeline of layer_fused_kernel CTAs (debug aid): runs one MS-WSA layer at a stage shape of the 1 Mpx B=8
workload with sast_debug_trace(which=4) armed and prints, per phase of each CTA's SECOND tile, the median / p90 clock
deltas, plus the per-CTA totals.  Usage: fused_trace.py [stage 1|2] [keep 0..1]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("SAST_B200_LIB", os.path.join(ROOT, "sast_b200", "libsast_b200_trace.so"))
if not os.path.exists(os.environ["SAST_B200_LIB"]):
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "sast_b200", "csrc"), "trace"])
sys.path.insert(0, ROOT)
import sast_b200  # noqa: E402
from sast_b200 import _lib as L, ops  # noqa: E402
from sast_b200.config import backbone_config  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
keep = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda:0")
B, C = 8, 64 << (stage - 1)
H, W = 96 >> (stage - 1), 160 >> (stage - 1)
p0, p1 = 6, 10
NW = B * (H // p0) * (W // p1)
net = sast_b200.build_recurrent_backbone(backbone_config((384, 640))).to(dev).eval()
layer = net.stages[stage - 1].att_blocks[0].att.win_attn
x = torch.randn(B, H, W, C, device=dev)
g = torch.Generator().manual_seed(3)
rho = keep ** 0.5
wf = (torch.rand(NW, generator=g) < rho).to(torch.uint8) if keep < 1 else torch.ones(NW, dtype=torch.uint8)
tf = (torch.rand(NW * p0 * p1, generator=g) < rho).to(torch.uint8) if keep < 1 else torch.ones(NW * p0 * p1, dtype=torch.uint8)
sel = ops.Selection(ops.select_from_flags(wf.to(dev), tf.to(dev), B, H, W, p0, p1, L.WINDOW), B, H, W, p0, p1)
with torch.no_grad():
    for _ in range(3):
        layer.run(x, sel, L.WINDOW, False)
    torch.cuda.synchronize()
    buf = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
    L.lib().sast_debug_trace(buf.data_ptr(), 4 if stage <= 2 else 5)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush.zero_()
    layer.run(x, sel, L.WINDOW, False)
    torch.cuda.synchronize()
    L.lib().sast_debug_trace(None, 0)
if stage >= 3:          # group kernel (C = 256 / 512): stamps of each CTA's first tile
    t = buf.view(-1, 32).cpu()
    t = t[t[:, 22] != 0]
    cnt = sel.counts.tolist()
    print(f"stage {stage}: C={C} keep={keep} S={cnt[1]} tiles={cnt[3]} traced CTAs={len(t)}")
    names = ["LN of own rows -> scratch", "group barrier 0", "QKV stream + mma", "QKV epilogue + sync", "attention (S, softmax, PV, O -> scratch)",
             "group barrier 1", "proj stream + mma", "proj epilogue", "group barrier 2", "GLU stream (2 passes) + mma", "GLU epilogue -> scratch",
             "group barrier 3", "out stream + mma", "out epilogue + stores + sync"]
    d = (t[:, 1:15] - t[:, 0:14]).float()
    for i, n in enumerate(names):
        print(f"  {n:44s} median {d[:, i].median():8.0f}  p90 {d[:, i].quantile(0.9):8.0f} clk")
    print(f"  first tile total: median {(t[:, 14] - t[:, 0]).float().median():.0f} clk")
    for a, b, n in ((20, 21, "entry -> set-up done"), (21, 22, "all tiles"), (22, 23, "unselected pass"), (20, 23, "kernel total")):
        col = (t[:, b] - t[:, a]).float()
        print(f"  [{n:38s}] median {col.median():9.0f}  max {col.max():9.0f} clk")
    sys.exit(0)
t = buf.view(-1, 32).cpu()
t = t[t[:, 17] != 0]
t2 = t[t[:, 14] != 0]
cnt = sel.counts.tolist()
print(f"stage {stage}: C={C} keep={keep} S={cnt[1]} tiles={cnt[3]} traced CTAs={len(t)} (tiles per CTA {t[:, 20].float().mean():.2f})")
names = ["gather + LN1/LN2 + sync", "QKV mma -> done (+ shortcut regs)", "sync + QKV epilogue + sync", "S mma -> done",
         "softmax pass 1 + sync", "softmax pass 2 (P -> TMEM) + sync", "PV mma -> done", "O epilogue + sync", "proj mma -> done",
         "proj epilogue + sync", "GLU mma -> done", "GLU epilogue + sync", "out mma -> done", "out epilogue + sync"]
d = (t2[:, 1:15] - t2[:, 0:14]).float()
if len(t2) == 0:
    d = torch.zeros(1, 14)
for i, n in enumerate(names):
    col = d[:, i]
    print(f"  {n:40s} median {col.median():8.0f}  p90 {col.quantile(0.9):8.0f} clk")
for a, b, n in ((13, 21, "out: tmem ld + math + st.shared"), (21, 22, "out: fence + sync"), (22, 23, "out: bulk store issue"), (23, 14, "out: final sync")):
    if len(t2) and (t2[:, 21] != 0).all():
        col = (t2[:, b] - t2[:, a]).float()
        print(f"    {n:38s} median {col.median():8.0f}  p90 {col.quantile(0.9):8.0f} clk")
tot = (t2[:, 14] - t2[:, 0]).float() if len(t2) else torch.zeros(1)
print(f"  second tile total: median {tot.median():.0f} clk, p90 {tot.quantile(0.9):.0f}")
for a, b, n in ((15, 16, "entry -> set-up done (incl. PDL wait)"), (16, 17, "all tiles"), (17, 18, "unselected pass"), (15, 18, "kernel total")):
    col = (t[:, b] - t[:, a]).float()
    print(f"  [{n:38s}] median {col.median():9.0f}  max {col.max():9.0f} clk")
