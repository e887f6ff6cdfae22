"""non_zero_ratio (bit-exact), scoring/STP weighting and the tcgen05 GEMM through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import sast_oracle as O
from oracle.golden_common import event_histogram, make_params
from sast_b200 import _lib as L
from sast_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_nonzero_ratio_golden(golden):
    g = golden("small_fns")
    for nm in ("u8", "i32", "f32"):
        x = g.t(f"nzr_{nm}_x")
        assert torch.equal(ops.nonzero_ratio(x.to(DEV)).cpu(), g.t(f"nzr_{nm}_r")), nm


@pytest.mark.parametrize("shape,dtype", [((8, 20, 384, 640), torch.uint8), ((2, 20, 256, 320), torch.int32),
                                         ((1, 20, 250, 300), torch.float32), ((2, 3, 100, 36), torch.uint8)])
def test_nonzero_ratio_full_size(shape, dtype):
    """Bit-exact at the 1 Mpx B=8 size, ragged sizes and negative values (max-pool semantics)."""
    B, C, H, W = shape
    x = event_histogram(B, C, H, W, 0.03, seed=H).to(dtype)
    if dtype == torch.float32:
        x = x - 2.0 * (x == 3)          # some negative cells: max-pool of {-2, 0} is 0 -> not counted
    assert torch.equal(ops.nonzero_ratio(x.to(DEV)).cpu(), O.non_zero_ratio(x))


@pytest.mark.parametrize("B,H,W,C", [(2, 12, 20, 64), (1, 16, 30, 128), (2, 6, 10, 256), (1, 4, 5, 512), (3, 7, 9, 32), (2, 12, 20, 96)])
def test_score_fwd(B, H, W, C):
    """a4: weighted map within fp32 round-off of the reference formula; per-token score (what
    selection thresholds) within 1e-5 relative."""
    shapes = {"to_scores.weight": (C, C), "to_scores.bias": (C,), "to_controls.weight": (C, 20)}
    p = make_params(shapes, seed=C + H)
    gen = torch.Generator().manual_seed(C)
    x = torch.randn(B, H, W, C, generator=gen)
    r = torch.rand(B, 20, generator=gen) * 0.05
    pos = O.position_table(H, W, C)
    amp = 2e-3
    # oracle on the un-partitioned map: partition (1,1) keeps map order
    xw_ref, scores = O.scoring(x, pos, r, p, (1, 1), amp)
    xw_ref = xw_ref.reshape(B, H, W, C)
    tok_ref = scores.abs().sum(-1).reshape(B, H, W)
    xw, tok = ops.score_fwd(x.to(DEV), pos.to(DEV), r.to(DEV), p["to_controls.weight"].to(DEV),
                            p["to_scores.weight"].to(DEV), p["to_scores.bias"].to(DEV), amp)
    assert (xw.cpu() - xw_ref).abs().max() < 2e-5
    assert ((tok.cpu() - tok_ref).abs() / tok_ref.abs().clamp_min(1e-12)).max() < 2e-5
    # 3xTF32 tensor-core kernel (what the bf16 precision mode runs): same bars
    w_hi, w_lo = ops.split_tf32(p["to_scores.weight"].to(DEV))
    assert (w_hi + w_lo - p["to_scores.weight"].to(DEV)).abs().max() < 1e-6
    xw3, tok3 = ops.score_fwd(x.to(DEV), pos.to(DEV), r.to(DEV), p["to_controls.weight"].to(DEV),
                              p["to_scores.weight"].to(DEV), p["to_scores.bias"].to(DEV), amp, w_hi, w_lo)
    assert (xw3.cpu() - xw_ref).abs().max() < 2e-5
    assert ((tok3.cpu() - tok_ref).abs() / tok_ref.abs().clamp_min(1e-12)).max() < 2e-5
    # batched pos (reference-style repeated tensor) gives the same result
    xw2, tok2 = ops.score_fwd(x.to(DEV), pos[None].repeat(B, 1, 1, 1).to(DEV), r.to(DEV), p["to_controls.weight"].to(DEV),
                              p["to_scores.weight"].to(DEV), p["to_scores.bias"].to(DEV), amp)
    assert torch.equal(xw2, xw) and torch.equal(tok2, tok)
    assert torch.equal(ops.add_pos(x.to(DEV), pos.to(DEV)).cpu(), x + pos)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 192, 64), (1000, 128, 128), (257, 320, 160), (128, 512, 1344),
                                   (64, 32, 32), (4096, 1536, 512), (130, 96, 96)])
def test_gemm_tcgen05(M, N, K):
    """tcgen05 GEMM vs a torch fp32 reference on the same bf16-rounded operands (fp32
    accumulate: only summation order differs)."""
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=gen).to(torch.bfloat16)
    Wt = (torch.randn(N, K, generator=gen) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=gen)
    ref = A.float() @ Wt.float().t() + bias
    D = ops.gemm_bf16(A.to(DEV), Wt.to(DEV), bias.to(DEV))
    assert (D.cpu() - ref).abs().max() < 1e-3 * max(1.0, K ** 0.5 / 8)
    Db = ops.gemm_bf16(A.to(DEV), Wt.to(DEV), None, out_bf16=True)
    ref_b = (A.float() @ Wt.float().t())
    assert (Db.float().cpu() - ref_b).abs().max() < 4e-2


@pytest.mark.parametrize("tensor_core", [False, True], ids=["fp32-fma", "tf32x3-tcgen05"])
def test_score_to_selection_flip_rate(tensor_core):
    """Tier B end to end at the 1 Mpx stage-1 map size: scoring GEMM -> per-token score -> softmax ->
    threshold on the GPU against the reference pipeline on the CPU.  A flip needs a probability within
    rounding distance of the threshold: <= 1e-4 of the tokens."""
    B, H, W, C, part = 2, 96, 160, 64, (6, 10)
    T, N = 60, 96 * 160 // 60
    shapes = {"to_scores.weight": (C, C), "to_scores.bias": (C,), "to_controls.weight": (C, 20)}
    p = make_params(shapes, seed=5)
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(B, H, W, C, generator=gen) * torch.linspace(0.3, 1.7, W).view(1, 1, W, 1)
    r = torch.rand(B, 20, generator=gen) * 0.02
    pos = O.position_table(H, W, C)
    amp = 2e-3
    _, scores = O.scoring(x, pos, r, p, part, amp)                      # [B,N,T,C] window-partitioned
    lists = O.select_layer(scores, T, 1e-3)
    ref = torch.zeros(B * N * T, dtype=torch.bool)
    ref[lists[0][lists[3] // T] * T + lists[3] % T] = True
    hi_lo = ops.split_tf32(p["to_scores.weight"].to(DEV)) if tensor_core else (None, None)
    _, tok = ops.score_fwd(x.to(DEV), pos.to(DEV), r.to(DEV), p["to_controls.weight"].to(DEV),
                           p["to_scores.weight"].to(DEV), p["to_scores.bias"].to(DEV), amp, *hi_lo)
    thr_w, thr_t = ops.thresholds(N, T, 1e-3)
    sel = ops.Selection(ops.select(tok, 6, 10, L.WINDOW, thr_w, thr_t), B, H, W, 6, 10)
    got = (sel.tok_row >= 0).cpu()
    flips = int((got != ref).sum())
    assert 0.2 < ref.float().mean() < 0.9           # a genuinely partial selection
    assert flips <= 1e-4 * ref.numel() + 1, f"{flips} flips of {ref.numel()}"


@pytest.mark.parametrize("M,I,K", [(300, 160, 64), (1000, 320, 128), (130, 672, 256)])
def test_gemm_glu_epilogue(M, I, K):
    """GLU epilogue of the tcgen05 GEMM (ops.py:135-137: value * erf-GELU(gate)) on interleaved weight rows."""
    gen = torch.Generator().manual_seed(M + I)
    A = torch.randn(M, K, generator=gen).to(torch.bfloat16)
    W = (torch.randn(2 * I, K, generator=gen) / K ** 0.5).to(torch.bfloat16)       # [value rows | gate rows]
    b = torch.randn(2 * I, generator=gen)
    y = A.float() @ W.float().t() + b
    ref = y[:, :I] * torch.nn.functional.gelu(y[:, I:])
    Wi = torch.stack((W[:I], W[I:]), dim=1).reshape(2 * I, K)
    bi = torch.stack((b[:I], b[I:]), dim=1).reshape(-1)
    got = ops.gemm_bf16_glu(A.to(DEV), Wi.to(DEV), bi.to(DEV)).float().cpu()
    assert ((got - ref).abs() / (ref.abs() + 1.0)).max() < 1e-2      # bf16 output: 2^-8 relative


@pytest.mark.parametrize("bits,B,H,W", [(1, 8, 384, 640), (4, 8, 384, 640), (1, 2, 66, 72), (4, 3, 240, 304)])
def test_unpack_nonzero_ratio(bits, B, H, W):
    """Bit-packed input path: unpacked histogram and r bit-identical to the uint8 path (and hence to the reference)."""
    import sast_b200
    from sast_b200 import ops
    g = torch.Generator().manual_seed(bits + H)
    if bits == 1:
        x = (torch.rand(B, 20, H, W, generator=g) > 0.97).to(torch.uint8)
    else:
        x = torch.poisson(torch.full((B, 20, H, W), 0.05), generator=g).clamp_(max=10).to(torch.uint8)
    x[0, 3] = 0                                            # an empty bin plane
    pk = sast_b200.pack_events(x, bits).to("cuda:0")
    xu, r = ops.unpack_nonzero_ratio(pk.data, pk.bits, pk.width)
    assert torch.equal(xu.cpu(), x)
    assert torch.equal(r, ops.nonzero_ratio(x.to("cuda:0")))
