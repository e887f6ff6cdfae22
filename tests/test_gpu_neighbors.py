"""Glue kernels of the dense callers around the SAST block (stem input / padding, LayerNorm,
LSTM gates) and the two torch modules built on them, against the oracle."""
import pytest
import torch
import torch.nn.functional as F

import sast_b200
from oracle import sast_oracle as O
from oracle.golden_common import event_histogram, make_params
from sast_b200 import ops
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
