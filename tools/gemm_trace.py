#!/usr/bin/env python
"""Phase timeline of gemm_tc_kernel CTAs (debug aid): runs one standalone GEMM with sast_debug_trace armed and
prints, per tile slot, the median clock deltas of the MMA warp / epilogue groups / producer.
Usage: gemm_trace.py [glu|store] [M N K]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("SAST_B200_LIB", os.path.join(ROOT, "sast_b200", "libsast_b200_trace.so"))
if not os.path.exists(os.environ["SAST_B200_LIB"]):          # the instrumented twin is not part of the default build
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "sast_b200", "csrc"), "trace"])
sys.path.insert(0, ROOT)
from sast_b200 import _lib as L  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "glu"
if kind == "score":          # scoring kernel at the stage-1 shape of the 1 Mpx B=8 workload (same stamp layout)
    import sast_b200
    from sast_b200 import ops
    from sast_b200.config import backbone_config
    dev = torch.device("cuda:0")
    stage = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    net = sast_b200.build_recurrent_backbone(backbone_config((384, 640))).to(dev).eval()
    blk = net.stages[stage - 1].att_blocks[0].att
    B, C = 8, 64 << (stage - 1)
    H, W = 96 >> (stage - 1), 160 >> (stage - 1)
    x = torch.randn(B, H, W, C, device=dev)
    pos = torch.randn(H, W, C, device=dev)
    r = torch.rand(B, 20, device=dev)
    with torch.no_grad():
        hi, lo = ops.split_tf32(blk.to_scores.weight)
        run = lambda: ops.score_fwd(x, pos, r, blk.to_controls.weight, blk.to_scores.weight, blk.to_scores.bias, 2e-4, hi, lo)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        buf = torch.zeros(148 * 128, dtype=torch.int64, device=dev)
        torch.empty(256 << 20, dtype=torch.uint8, device=dev).zero_()
        L.lib().sast_debug_trace(buf.data_ptr(), 3)
        run()
        torch.cuda.synchronize()
        L.lib().sast_debug_trace(None, 0)
    t = buf.view(148, 128).cpu()
    t = t[t[:, 0] != 0]
    print(f"score_tc stage {stage} [{B * H * W} x {C} x {C}]; CTAs traced {len(t)}")

else:
    M, N, K = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (122880, 320 if kind == "glu" else 192, 64)
    dev = torch.device("cuda:0")
    lib = L.lib()
    st = L.stream_ptr(dev)
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / 8).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    D = torch.empty(M, N // 2 if kind == "glu" else N, device=dev, dtype=torch.bfloat16)


    def run():
        if kind == "glu":
            L.check(lib.sast_gemm_bf16_glu(A.data_ptr(), W.data_ptr(), bias.data_ptr(), D.data_ptr(), M, N, K, st), "glu")
        else:
            L.check(lib.sast_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), D.data_ptr(), 1, M, N, K, st), "gemm")


    for _ in range(3):
        run()
    torch.cuda.synchronize()
    buf = torch.zeros(148 * 128, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush.zero_()
    lib.sast_debug_trace(buf.data_ptr(), 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    lib.sast_debug_trace(None, 0)
    t = buf.view(148, 128).cpu()
    t = t[t[:, 0] != 0]
    print(f"{kind} GEMM [{M},{N},{K}]: {e0.elapsed_time(e1) * 1e3:.1f} us with stamps; CTAs traced {len(t)}")




def med(col):
    col = col[col > -(1 << 40)]
    return f"{col.float().median():7.0f}"


rel = t - t[:, :1]
print(f"  set-up done at {med(rel[:, 1])} clk, kernel end at {med(rel[:, 2])} clk after CTA entry")
print("  tile |  producer 1st slot | MMA: acc free, 1st k-block landed, last commit | epilogue: acc full, drained (duration)")
for ti in range(8):
    m = t[:, 8 + 4 * ti] != 0
    if m.sum() == 0:
        break
    r = rel[m]
    print(f"  {ti:4d} | {med(r[:, 100 + ti])}            | {med(r[:, 8 + 4 * ti])} {med(r[:, 9 + 4 * ti])} {med(r[:, 10 + 4 * ti])}"
          f"                    | {med(r[:, 48 + 4 * ti])} {med(r[:, 49 + 4 * ti])} ({med(r[:, 49 + 4 * ti] - r[:, 48 + 4 * ti])})   [{int(m.sum())} CTAs]")
if kind == "score" and (t[:, 113] != 0).any():      # loader phases of the third tile, both k-blocks
    m = t[:, 113] != 0
    r = rel[m]
    for kb in range(2):
        o = 110 + 4 * kb
        if (t[m][:, o + 3] == 0).all():
            continue
        print(f"  loader tile 2 k-block {kb}: loads issued -> slot free {med(r[:, o + 1] - r[:, o])}, convert + stores {med(r[:, o + 2] - r[:, o + 1])}, "
              f"fence + arrive {med(r[:, o + 3] - r[:, o + 2])}" + (f", to next k-block top {med(r[:, o + 4] - r[:, o + 3])}" if kb == 0 else ""))
