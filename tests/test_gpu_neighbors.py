"""Glue kernels of the dense callers around the SAST block (stem input / padding, LayerNorm,
LSTM gates) and the two torch modules built on them, against the oracle."""
import pytest
import torch
import torch.nn.functional as F

import sast_b200
from oracle import sast_oracle as O
from oracle.golden_common import event_histogram, make_params
from sast_b200 import _lib as L, ops
from sast_b200.config import Config

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int32, torch.float32, torch.float16])
def test_pad_input(dtype):
    x = event_histogram(2, 20, 36, 52, 0.2, seed=1).to(dtype)
    for pad in (0, 1, 3):
        ref = F.pad(x.float(), (pad, pad, pad, pad), mode="replicate").permute(0, 2, 3, 1) if pad else x.float().permute(0, 2, 3, 1)
        assert torch.equal(ops.pad_input(x.to(DEV), pad).cpu(), ref.contiguous())


def test_pad_nhwc_strided():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 9, 11, 64, generator=g)
    ref = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 1).contiguous()
    assert torch.equal(ops.pad_nhwc(x.to(DEV), 1).cpu(), ref)
    big = torch.randn(2, 9, 11, 128, generator=g).to(DEV)
    view = big[..., 64:]                                           # channel-sliced view: strides != dense
    ref = F.pad(view.cpu().permute(0, 3, 1, 2), (1, 1, 1, 1), mode="replicate").permute(0, 2, 3, 1).contiguous()
    assert torch.equal(ops.pad_nhwc(view, 1).cpu(), ref)


@pytest.mark.parametrize("C", [32, 64, 96, 128, 256, 512])
def test_layernorm(C):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(3, 7, 5, C, generator=g) * 3 + 1
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.layer_norm(x, (C,), w, b, 1e-5)
    assert (ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5).cpu() - ref).abs().max() < 2e-5
    assert (ops.layernorm(x.to(DEV), None, None, 1e-5).cpu() - F.layer_norm(x, (C,), None, None, 1e-5)).abs().max() < 2e-5


@pytest.mark.parametrize("precision", [0, 1], ids=["fp32-cudnn+gates", "tf32-tcgen05-fused"])
@pytest.mark.parametrize("C", [32, 64, 256, 512])
def test_lstm_module(C, precision):
    """DWSConvLSTM2d (ref: models/layers/rnn.py:36-69) with and without a carried state.  The fused kernel
    multiplies in TF32 (what torch's cuDNN path does by default on this GPU): 10-bit-significand bar."""
    torch.backends.cudnn.allow_tf32 = False
    tol = 2e-5 if precision == 0 else 4e-3
    try:
        lstm = sast_b200.DWSConvLSTM2d(C, dws_conv=False, dws_conv_only_hidden=True).eval()
        lstm.precision = precision
        p = make_params({"conv1x1.weight": (4 * C, 2 * C, 1, 1), "conv1x1.bias": (4 * C,)}, seed=C)
        lstm.load_state_dict(p)
        lstm = lstm.to(DEV)
        g = torch.Generator().manual_seed(1)
        x0, x1 = torch.randn(2, C, 12, 20, generator=g), torch.randn(2, C, 12, 20, generator=g)
        h_ref, c_ref = O.conv_lstm(x0, None, p)
        h2_ref, c2_ref = O.conv_lstm(x1, (h_ref, c_ref), p)
        with torch.no_grad():
            h, c = lstm(x0.to(DEV), None)
            h2, c2 = lstm(x1.to(DEV), (h, c))
            # states handed back in plain NCHW-contiguous form (as the reference's RNNStates would store them)
            h3, c3 = lstm(x1.to(DEV), (h.contiguous(), c.contiguous()))
        for got, ref in ((h, h_ref), (c, c_ref), (h2, h2_ref), (c2, c2_ref), (h3, h2_ref), (c3, c2_ref)):
            assert got.shape == ref.shape
            assert (got.cpu() - ref).abs().max() < tol
    finally:
        torch.backends.cudnn.allow_tf32 = True


@pytest.mark.parametrize("cin,cout,factor,dtype", [(20, 64, 4, torch.uint8), (20, 32, 4, torch.int32), (64, 128, 2, torch.float32)])
def test_downsample_module(cin, cout, factor, dtype):
    """ConvDownsampling_Cf2Cl (ref: ops.py:54-91): strided conv with replicate padding + LayerNorm, NCHW -> NHWC."""
    torch.backends.cudnn.allow_tf32 = False
    try:
        k = (factor - 1) * 2 + 1
        mod = sast_b200.ConvDownsampling_Cf2Cl(cin, cout, factor, Config(type="patch", overlap=True, norm_affine=True)).eval()
        p = make_params({"conv.weight": (cout, cin, k, k), "norm.weight": (cout,), "norm.bias": (cout,)}, seed=cin)
        mod.load_state_dict(p)
        mod = mod.to(DEV)
        if dtype == torch.float32:
            x = torch.randn(2, cin, 24, 40, generator=torch.Generator().manual_seed(2))
        else:
            x = event_histogram(2, cin, 48, 80, 0.1, seed=3).to(dtype)
        ref = O.conv_downsample(x.float(), p, factor)
        with torch.no_grad():
            got = mod(x.to(DEV))
            # NCHW-logical view over channels-last memory (what the previous stage's LSTM hands over)
            got2 = mod(x.float().to(DEV).contiguous(memory_format=torch.channels_last)) if dtype == torch.float32 else got
        assert got.shape == ref.shape and got.is_contiguous()
        assert (got.cpu() - ref).abs().max() < 5e-5
        assert (got2.cpu() - ref).abs().max() < 5e-5
    finally:
        torch.backends.cudnn.allow_tf32 = True


@pytest.mark.parametrize("B,H,W,cout,density", [(2, 48, 80, 64, 0.1), (1, 384, 640, 64, 0.3), (3, 36, 52, 32, 0.9), (2, 64, 64, 128, 0.02)])
def test_fused_stem(B, H, W, cout, density):
    """sast_stem_fwd (implicit-GEMM 7x7/4 conv with replicate padding + LayerNorm, uint8 in) against the oracle's
    fp32 conv: the bf16 hi/lo weight split keeps it fp32-grade (no TF32-sized error)."""
    mod = sast_b200.ConvDownsampling_Cf2Cl(20, cout, 4, Config(type="patch", overlap=True, norm_affine=True)).eval()
    p = make_params({"conv.weight": (cout, 20, 7, 7), "norm.weight": (cout,), "norm.bias": (cout,)}, seed=cout + H)
    mod.load_state_dict(p)
    mod = mod.to(DEV)
    x = event_histogram(B, 20, H, W, density, seed=H)
    x[:, :, 0, :] = 7          # make the replicated borders matter
    x[:, :, :, 0] = 9
    x[:, :, -1, :] = 3
    x[:, :, :, -1] = 5
    ref = O.conv_downsample(x.float(), p, 4)
    with torch.no_grad():
        got = mod(x.to(DEV))
        mod.fused_stem = False
        torch.backends.cudnn.allow_tf32 = False
        try:
            got_cudnn = mod(x.to(DEV))
        finally:
            torch.backends.cudnn.allow_tf32 = True
    assert got.shape == ref.shape
    assert (got.cpu() - ref).abs().max() < 1e-4, (got.cpu() - ref).abs().max()
    assert (got_cudnn.cpu() - ref).abs().max() < 1e-4


@pytest.mark.parametrize("bits,B,H,W,density", [(1, 2, 384, 640, 0.03), (4, 1, 256, 320, 0.3), (8, 3, 100, 96, 0.5),
                                                (1, 1, 64, 64, 0.01), (8, 2, 36, 160, 0.9), (1, 3, 100, 96, 0.5),
                                                (1, 1, 256, 320, 0.97)])
def test_nhwc_stem(bits, B, H, W, density):
    """The TMA-fed stem: sast_events_nhwc (histogram -> fp16 NHWC with the replicate padding materialised, + r) is
    EXACT against plain torch ops; sast_stem_nhwc_fwd (im2col by TMA, fp16 weights resident) against the oracle's fp32
    conv + LayerNorm within the fp16 weight rounding (2^-12 relative per weight; the TF32 cuDNN route it replaces
    rounds at 2^-11).  Geometries include ragged tile edges (Ho % 8 != 0, Wo % 16 != 0) and every input format."""
    cout = 64
    p = make_params({"conv.weight": (cout, 20, 7, 7), "norm.weight": (cout,), "norm.bias": (cout,)}, seed=bits + H)
    if bits == 1:
        x = (torch.rand(B, 20, H, W, generator=torch.Generator().manual_seed(H)) < density).to(torch.uint8)
        x[:, ::2, :, 0] = 1        # make the replicated borders matter
        x[:, 1::3, 0, :] = 1
        x[:, ::5, -1, :] = 1
    else:
        x = event_histogram(B, 20, H, W, density, seed=H).clamp_(max=15 if bits == 4 else 255)
        x[:, :, 0, :] = 7          # make the replicated borders matter
        x[:, :, :, 0] = 9
        x[:, :, -1, :] = 3
        x[:, :, :, -1] = 5
    x[0, 3] = 0
    xd = x.to(DEV)
    src = xd if bits == 8 else sast_b200.pack_events(xd, bits).data
    xh, r = ops.events_nhwc(src, bits, W, True)
    ref_h = F.pad(x.float(), (3, 3, 3, 3), mode="replicate").permute(0, 2, 3, 1)
    assert xh.shape == (B, H + 8, W + 8, 20) and xh.dtype == torch.float16
    # rows / columns up to H+2 / W+2 are what the stride-4 windows reach (4 (Ho-1) + 6 = H + 2); the rest is zero filler
    assert torch.equal(xh[:, :H + 3, :W + 3].float().cpu(), ref_h[:, :H + 3, :W + 3])
    assert float(xh[:, H + 3:].float().abs().max()) == 0 and float(xh[:, :, W + 3:].float().abs().max()) == 0
    assert torch.equal(r, ops.nonzero_ratio(xd))
    assert ops.stem_nhwc_supported(20, H, W, cout)
    got = ops.stem_nhwc_fwd(xh, H, W, ops.pack_stem_weight_nhwc(p["conv.weight"].to(DEV)), p["norm.weight"].to(DEV),
                            p["norm.bias"].to(DEV), 1e-5)
    ref = O.conv_downsample(x.float(), p, 4)
    assert got.shape == ref.shape
    d = (got.cpu() - ref).abs()
    assert d.max() < 4e-3 and d.mean() < 4e-4, (d.max(), d.mean())
    if bits == 1:
        # the stem that expands the packed bits itself: same fp16 products as the NHWC route, other summation order
        assert ops.stem_bits_supported(1, 20, H, W, cout)
        got_b = ops.stem_bits_fwd(src, 1, W, ops.pack_stem_weight_bits(p["conv.weight"].to(DEV)), p["norm.weight"].to(DEV),
                                  p["norm.bias"].to(DEV), 1e-5)
        assert (got_b - got).abs().max() < 2e-4, (got_b - got).abs().max()
        assert torch.equal(ops.packed_nonzero_ratio(src, 1, W), r)
    # the same through the module: a backbone stem in the 16-bit mode takes this path, the fp32 mode the split-weight kernel
    mod = sast_b200.ConvDownsampling_Cf2Cl(20, cout, 4, Config(type="patch", overlap=True, norm_affine=True)).eval()
    mod.load_state_dict(p)
    mod = mod.to(DEV)
    with torch.no_grad():
        assert mod.nhwc_stem_ok(20, H, W)
        assert torch.equal(mod(ops.EventsNHWC(xh, H, W)), got)
        torch.backends.cudnn.allow_tf32 = False
        try:
            assert not mod.nhwc_stem_ok(20, H, W)
        finally:
            torch.backends.cudnn.allow_tf32 = True


@pytest.mark.parametrize("cin,cout,B,H,W", [(64, 128, 8, 96, 160), (128, 256, 8, 48, 80), (64, 128, 1, 64, 80), (128, 256, 2, 22, 36),
                                            (64, 128, 3, 10, 24)])
def test_fused_downsample(cin, cout, B, H, W):
    """sast_pad_nhwc_bf16 + sast_downsample_fwd (stages 2-3: k3 s2 conv with replicate padding + LayerNorm as ONE tcgen05
    kernel, bf16 operands, im2col by an overlapping-stride TMA tensor map) against the oracle's fp32 conv + LayerNorm within
    bf16 operand rounding (2^-9 relative per product, fp32 accumulation), and against the cuDNN TF32 route it replaces.
    Geometries: both channel pairs at the 1 Mpx B=8 shapes, Gen1, ragged tiles (Ho % 16, Wo % 8 != 0)."""
    mod = sast_b200.ConvDownsampling_Cf2Cl(cin, cout, 2, Config(type="patch", overlap=True, norm_affine=True)).eval()
    p = make_params({"conv.weight": (cout, cin, 3, 3), "norm.weight": (cout,), "norm.bias": (cout,)}, seed=cin + H)
    mod.load_state_dict(p)
    mod = mod.to(DEV)
    x = torch.randn(B, cin, H, W, generator=torch.Generator().manual_seed(H))
    x[:, :, 0, :] += 2.0          # make the replicated borders matter
    x[:, :, :, -1] -= 2.0
    ref = O.conv_downsample(x, p, 2)
    assert ops.downsample_supported(cin, H, W, cout)
    with torch.no_grad():
        xd = x.to(DEV).contiguous(memory_format=torch.channels_last)     # what the previous stage's LSTM hands over
        before = L.lib().sast_launch_count()
        got = mod(xd)
        assert L.lib().sast_launch_count() - before == 2                  # bf16 pad + the fused kernel (cuDNN route: pad + LayerNorm)
        mod.fused_downsample = False
        got_cudnn = mod(xd)
        mod.fused_downsample, mod.precision = True, L.FP32                # the fp32 validation mode never takes the bf16 kernel
        before = L.lib().sast_launch_count()
        got_fp32mode = mod(xd)
        assert L.lib().sast_launch_count() - before == 2 and torch.equal(got_fp32mode, got_cudnn)
    assert got.shape == ref.shape and got.is_contiguous()
    d = (got.cpu() - ref).abs()
    assert d.max() < 4e-2 and d.mean() < 4e-3, (d.max(), d.mean())
    d2 = (got - got_cudnn).abs()
    assert d2.max() < 4e-2 and d2.mean() < 4e-3, (d2.max(), d2.mean())
    assert (got_cudnn.cpu() - ref).abs().mean() < 1e-3
