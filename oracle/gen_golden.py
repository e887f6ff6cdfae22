"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python oracle/gen_golden.py

The reference's modules are imported as they are, through the ``omegaconf`` stand-in
in ``oracle/_shim``.  Weights and inputs come from numpy seeds
(``oracle/golden_common.py``) so the fixtures hold only shapes, seeds and the
reference's outputs.  Test infrastructure only.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from omegaconf import DictConfig  # noqa: E402  (the stand-in)
from models.layers.SAST import SAST as ref_sast  # noqa: E402
from models.layers.SAST import ops as ref_ops  # noqa: E402
from models.detection.recurrent_backbone import build_recurrent_backbone  # noqa: E402
from models.detection.recurrent_backbone import sast_rnn as ref_rnn  # noqa: E402

from oracle.golden_common import (canonical_keys, event_histogram, make_det_params, make_params,  # noqa: E402
                                  with_aliases)

OUT = os.path.join(ROOT, "tests", "golden")


def attention_cfg(part, **kw):
    d = dict(partition_size=part, dim_head=32, attention_bias=True, mlp_activation="gelu",
             mlp_bias=True, mlp_ratio=4, drop_mlp=0, drop_path=0, ls_init_value=1e-5,
             enable_CB=False, AMP=2e-4, BOUNCE=1e-3)
    d.update(kw)
    return DictConfig(d)


def backbone_cfg(embed_dim, in_res_hw, part, num_blocks=(1, 1, 1, 1), enable_masking=False, **att):
    return DictConfig(dict(
        name="SASTRNN", input_channels=20, enable_masking=enable_masking, partition_split_32=2,
        embed_dim=embed_dim, dim_multiplier=[1, 2, 4, 8], num_blocks=list(num_blocks),
        T_max_chrono_init=[4, 8, 16, 32], stem=dict(patch_size=4), in_res_hw=list(in_res_hw),
        stage=dict(downsample=dict(type="patch", overlap=True, norm_affine=True),
                   attention=dict(attention_cfg(part, **att)),
                   lstm=dict(dws_conv=False, dws_conv_only_hidden=True, dws_conv_kernel_size=3,
                             drop_cell_update=0))))


def save(name, **arrays):
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print(f"wrote {name}.npz  ({os.path.getsize(os.path.join(OUT, name + '.npz')) / 1024:.0f} KiB)")


# --------------------------------------------------------------------------- #
def gen_select():
    """Known-answer vectors for get_score_index_2d21d / get_score_index_with_padding
    (ref: SAST.py:258-281), including values exactly at the fp32-cast threshold."""
    rng = np.random.RandomState(11)
    arrays = {}
    cases = []
    for ci, (B, N, T, sharp) in enumerate([(2, 4, 60, 3.0), (1, 6, 80, 2.0), (3, 16, 60, 1.0),
                                           (8, 256, 60, 0.5), (4, 1, 80, 1.0), (2, 64, 15, 4.0),
                                           (1, 1, 80, 0.01), (5, 4, 128, 2.0)]):
        b = 1e-3
        pw = torch.softmax(torch.from_numpy(rng.standard_normal((B, N)).astype(np.float32) * sharp * 0.05), -1)
        thr_w = np.float32((1 / N) / (1 + b))
        if N > 1:  # plant exact-threshold and just-below values
            pw[0, 0] = float(thr_w)
            pw[-1, -1] = float(np.nextafter(thr_w, np.float32(0)))
        iw = ref_sast.get_score_index_2d21d(pw.view(B, N), 1 / N, b)
        M = len(iw)
        pt = torch.softmax(torch.from_numpy(rng.standard_normal((M, T)).astype(np.float32) * sharp * 0.02), -1)
        thr_t = np.float32((1 / T) / (1 + b))
        pt[0, 1] = float(thr_t)
        pt[0, 2] = float(np.nextafter(thr_t, np.float32(0)))
        pt[-1, T - 1] = float(np.nextafter(thr_t, np.float32(1)))
        it, asy, K = ref_sast.get_score_index_with_padding(pt, 1 / T, b)
        pad = it[torch.isin(it, asy, assume_unique=True, invert=True)]
        arrays.update({f"c{ci}_pw": pw, f"c{ci}_pt": pt, f"c{ci}_iw": iw, f"c{ci}_it": it,
                       f"c{ci}_asy": asy, f"c{ci}_K": K, f"c{ci}_pad": pad})
        cases.append(dict(B=B, N=N, T=T, bounce=b))
    save("select_kat", meta=json.dumps(cases), **arrays)


def gen_small_fns():
    rng = np.random.RandomState(3)
    arrays = {}
    # non_zero_ratio (ref: sast_rnn.py:45-60) on three input dtypes, one with odd sizes
    x8 = event_histogram(2, 20, 64, 96, 0.05, seed=21)
    x32 = event_histogram(3, 20, 72, 100, 0.3, seed=22, dtype=np.int32)
    xf = torch.from_numpy(rng.standard_normal((1, 20, 32, 64)).astype(np.float32))
    xf[xf.abs() < 1.2] = 0
    for nm, x in (("u8", x8), ("i32", x32), ("f32", xf)):
        arrays[f"nzr_{nm}_x"] = x
        arrays[f"nzr_{nm}_r"] = ref_rnn.non_zero_ratio(x)
    # position embedding (ref: sast_rnn.py:180-219): table for (H,W)=(12,20),C=64 and a slice
    pe = ref_rnn.PositionEmbeddingSine(32, normalize=True, input_size=(1, 12, 20))
    arrays["pos_12_20_64"] = pe(torch.zeros(2, 12, 20, 64))[0]
    arrays["pos_12_20_64_slice"] = pe(torch.zeros(1, 6, 10, 64))[0]
    pe = ref_rnn.PositionEmbeddingSine(64, normalize=True, input_size=(1, 8, 10))
    arrays["pos_8_10_128"] = pe(torch.zeros(1, 8, 10, 128))[0]
    # partition index maps (ref: ops.py:189-220): where does pixel id land
    ids = torch.arange(2 * 12 * 20, dtype=torch.float32).view(2, 12, 20, 1)
    arrays["win_ids_6x10"] = ref_ops.window_partition(ids, (6, 10)).reshape(-1).to(torch.int32)
    arrays["grid_ids_6x10"] = ref_ops.grid_partition(ids, (6, 10)).reshape(-1).to(torch.int32)
    ids = torch.arange(1 * 16 * 30, dtype=torch.float32).view(1, 16, 30, 1)
    arrays["win_ids_8x10"] = ref_ops.window_partition(ids, (8, 10)).reshape(-1).to(torch.int32)
    arrays["grid_ids_8x10"] = ref_ops.grid_partition(ids, (8, 10)).reshape(-1).to(torch.int32)
    save("small_fns", **arrays)


class _Pos(torch.nn.Module):
    """Reference-style pos_emb callable built from the reference's own class."""

    def __init__(self, C, H, W):
        super().__init__()
        self.pe = ref_rnn.PositionEmbeddingSine(C // 2, normalize=True, input_size=(1, H, W))

    def forward(self, x):
        return self.pe(x)


def run_block(name, C, part, B, H, W, amp, seed, r_scale, enable_CB=False, second_block=False):
    """SAST_block.forward (ref: SAST.py:98-164) on seeded inputs."""
    cfg = attention_cfg(part, AMP=amp, enable_CB=enable_CB)
    blk = ref_sast.SAST_block(C, cfg, first_block=True).eval()
    sd = blk.state_dict()
    shapes = canonical_keys(sd)
    params = make_params(shapes, seed=seed)
    blk.load_state_dict(with_aliases(params, sd.keys()), strict=True)
    rng = np.random.RandomState(seed + 1000)
    x = torch.from_numpy(rng.standard_normal((B, H, W, C)).astype(np.float32))
    # spatially varying magnitude so windows differ
    x = x * torch.linspace(0.3, 1.7, W).view(1, 1, W, 1) * torch.linspace(1.5, 0.5, H).view(1, H, 1, 1)
    r = torch.from_numpy((rng.rand(B, 20) * r_scale).astype(np.float32))
    pos = _Pos(C, H, W)
    with torch.no_grad():
        y, cnt, lists = blk(x, pos, r, None)
    arrays = dict(x=x, r=r, y=y, count=np.int64(cnt))
    for li, lst in enumerate(lists):
        for nm, t in zip(("iw", "it", "pad", "asy", "K"), lst):
            arrays[f"l{li}_{nm}"] = t
    meta = dict(C=C, part=list(part), B=B, H=H, W=W, AMP=amp, BOUNCE=1e-3, enable_CB=enable_CB,
                seed=seed, shapes={k: list(v) for k, v in shapes.items()})
    if second_block:
        blk2 = ref_sast.SAST_block(C, cfg, first_block=False).eval()
        sd2 = blk2.state_dict()
        shapes2 = canonical_keys(sd2)
        params2 = make_params(shapes2, seed=seed + 7)
        blk2.load_state_dict(with_aliases(params2, sd2.keys()), strict=True)
        with torch.no_grad():
            y2, cnt2, _ = blk2(y, pos, r, lists)
        arrays.update(y2=y2, count2=np.int64(cnt2))
        meta.update(shapes2={k: list(v) for k, v in shapes2.items()}, seed2=seed + 7)
    save(name, meta=json.dumps(meta), **arrays)
    print("   ", name, "count", cnt, "M", [len(l[0]) for l in lists], "S", [len(l[3]) for l in lists])


def gen_backbone():
    """RNNDetector.forward, two recurrent steps with state carry (ref: sast_rnn.py:144-162)."""
    embed, res, part = 32, (192, 320), (3, 5)
    for name, nb, masking, cb in (("backbone_e32", (1, 1, 1, 1), False, False),
                                  ("backbone_e32_nb2_mask_cb", (2, 1, 1, 1), True, True)):
        cfg = backbone_cfg(embed, res, part, num_blocks=nb, enable_masking=masking, AMP=2e-4, enable_CB=cb)
        net = build_recurrent_backbone(cfg).eval()
        sd = net.state_dict()
        shapes = canonical_keys(sd)
        params = make_params(shapes, seed=77)
        net.load_state_dict(with_aliases(params, sd.keys()), strict=True)
        B = 2
        x0 = event_histogram(B, 20, res[0], res[1], 0.02, seed=5)
        x1 = event_histogram(B, 20, res[0], res[1], 0.004, seed=6)
        tm = None
        if masking:
            tm = torch.from_numpy(np.random.RandomState(9).rand(B, res[0] // 4, res[1] // 4) < 0.1)
        with torch.no_grad():
            f0, s0, p0 = net(x0, None, tm)
            f1, s1, p1 = net(x1, s0, tm)
        arrays = dict(P0=np.array(p0, dtype=np.int64), P1=np.array(p1, dtype=np.int64))
        if tm is not None:
            arrays["token_mask"] = tm
        for st in (1, 2, 3, 4):
            h = f1[st]
            arrays[f"h1_s{st}"] = h[:, :, ::2, ::2] if st == 1 else h
            arrays[f"sum0_s{st}"] = np.array([f0[st].double().sum().item(), f0[st].double().abs().sum().item()])
            arrays[f"sum1_s{st}"] = np.array([h.double().sum().item(), h.double().abs().sum().item()])
            arrays[f"csum1_s{st}"] = np.array([s1[st - 1][1].double().sum().item(),
                                               s1[st - 1][1].double().abs().sum().item()])
        meta = dict(embed_dim=embed, in_res_hw=list(res), partition_size=list(part), num_blocks=list(nb),
                    enable_masking=masking, enable_CB=cb, AMP=2e-4, BOUNCE=1e-3, B=B, seed=77,
                    x_seeds=[5, 6], x_density=[0.02, 0.004], mask_seed=9,
                    shapes={k: list(v) for k, v in shapes.items()},
                    all_keys=sorted(sd.keys()))
        save(name, meta=json.dumps(meta), **arrays)
        print("   ", name, "P", p0, p1)


def pick_conf(out, ncls):
    """Confidence threshold that lets ~15 % of the anchors through (random weights have no meaningful score scale),
    placed in the middle of a gap between two neighbouring scores so that rounding noise cannot move an anchor across it."""
    score = (out[..., 4] * out[..., 5:5 + ncls].max(-1)[0]).flatten().sort()[0]
    i = int(0.85 * score.numel())
    gaps = score[i + 1:i + 40] - score[i:i + 39]
    j = int(gaps.argmax())
    return float((score[i + j] + score[i + j + 1]) / 2)


def gen_yolox():
    """YOLOX PAFPN + head + postprocess on fixed features (ref: yolo_pafpn.py:109-139, yolo_head.py:165-289,
    boxes.py:32-76), and the whole detector on two recurrent steps (ref: detector.py:34-58)."""
    from models.detection.yolox_extension.models.build import build_yolox_fpn, build_yolox_head
    from models.detection.yolox.utils.boxes import postprocess

    # ---- neck + head on given features (Gen1 geometry: strides 8/16/32 of 256x320), with and without depthwise convs ----
    for name, depthwise, depth in (("yolox_head_gen1", False, 0.67), ("yolox_head_dw", True, 0.33)):
        dims, strides, ncls, B = (64, 128, 256), (8, 16, 32), 2, 2
        fpn = build_yolox_fpn(DictConfig(dict(name="PAFPN", depth=depth, in_stages=[2, 3, 4], depthwise=depthwise, act="silu")),
                              in_channels=dims).eval()
        head = build_yolox_head(DictConfig(dict(name="YoloX", depthwise=depthwise, act="silu", num_classes=ncls)),
                                in_channels=dims, strides=strides).eval()
        shapes_f = {k: tuple(v.shape) for k, v in fpn.state_dict().items()}
        shapes_h = {k: tuple(v.shape) for k, v in head.state_dict().items()}
        fpn.load_state_dict(make_det_params(shapes_f, seed=21), strict=True)
        head.load_state_dict(make_det_params(shapes_h, seed=22), strict=True)
        rng = np.random.RandomState(23)
        feats = {st: torch.from_numpy(rng.standard_normal((B, c, 256 // s, 320 // s)).astype(np.float32))
                 for st, c, s in zip((2, 3, 4), dims, strides)}
        with torch.no_grad():
            fo = fpn(feats)
            out, losses = head(fo)
            assert losses is None
            conf = pick_conf(out, ncls)
            dets = postprocess(out.clone(), ncls, conf_thre=conf, nms_thre=0.45)
        arrays = dict(out=out, fpn0=fo[0][:, ::4], fpn2=fo[2][:, ::8])
        for i, d in enumerate(dets):
            arrays[f"det{i}"] = d if d is not None else torch.zeros(0, 7)
        meta = dict(dims=list(dims), strides=list(strides), num_classes=ncls, B=B, depthwise=depthwise, depth=depth,
                    seeds=[21, 22, 23], conf_thre=conf, nms_thre=0.45,
                    shapes_fpn={k: list(v) for k, v in shapes_f.items()}, shapes_head={k: list(v) for k, v in shapes_h.items()})
        save(name, meta=json.dumps(meta), **arrays)
        print("   ", name, "anchors", out.shape[1], "detections", [len(arrays[f"det{i}"]) for i in range(B)])

    # ---- whole detector: backbone (embed 32) + neck + head, two recurrent steps ----
    embed, res, part, ncls, B = 32, (192, 320), (3, 5), 3, 2
    cfg = backbone_cfg(embed, res, part, AMP=2e-4)
    net = build_recurrent_backbone(cfg).eval()
    dims, strides = net.get_stage_dims((2, 3, 4)), net.get_strides((2, 3, 4))
    fpn = build_yolox_fpn(DictConfig(dict(name="PAFPN", depth=0.67, in_stages=[2, 3, 4], depthwise=False, act="silu")),
                          in_channels=dims).eval()
    head = build_yolox_head(DictConfig(dict(name="YoloX", depthwise=False, act="silu", num_classes=ncls)),
                            in_channels=dims, strides=strides).eval()
    sd = net.state_dict()
    shapes = canonical_keys(sd)
    net.load_state_dict(with_aliases(make_params(shapes, seed=77), sd.keys()), strict=True)
    shapes_f = {k: tuple(v.shape) for k, v in fpn.state_dict().items()}
    shapes_h = {k: tuple(v.shape) for k, v in head.state_dict().items()}
    fpn.load_state_dict(make_det_params(shapes_f, seed=31), strict=True)
    head.load_state_dict(make_det_params(shapes_h, seed=32), strict=True)
    x0 = event_histogram(B, 20, res[0], res[1], 0.02, seed=5)
    x1 = event_histogram(B, 20, res[0], res[1], 0.004, seed=6)
    with torch.no_grad():
        f0, s0, p0 = net(x0, None, None)
        f1, s1, p1 = net(x1, s0, None)
        out, _ = head(fpn(f1))
        conf = pick_conf(out, ncls)
        dets = postprocess(out.clone(), ncls, conf_thre=conf, nms_thre=0.45)
    arrays = dict(out=out, P1=np.array(p1, dtype=np.int64))
    for i, d in enumerate(dets):
        arrays[f"det{i}"] = d if d is not None else torch.zeros(0, 7)
    meta = dict(embed_dim=embed, in_res_hw=list(res), partition_size=list(part), num_classes=ncls, B=B, AMP=2e-4,
                seeds=dict(backbone=77, fpn=31, head=32), x_seeds=[5, 6], x_density=[0.02, 0.004], conf_thre=conf, nms_thre=0.45,
                shapes={k: list(v) for k, v in shapes.items()}, all_keys=sorted(sd.keys()),
                shapes_fpn={k: list(v) for k, v in shapes_f.items()}, shapes_head={k: list(v) for k, v in shapes_h.items()})
    save("detector_e32", meta=json.dumps(meta), **arrays)
    print("    detector_e32 anchors", out.shape[1], "detections", [len(arrays[f"det{i}"]) for i in range(B)], "P", p1)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    os.makedirs(OUT, exist_ok=True)
    gen_select()
    gen_small_fns()
    run_block("block_c64_w6x10", 64, (6, 10), 2, 12, 20, amp=2e-3, seed=1, r_scale=0.02, second_block=True)
    run_block("block_c128_w8x10_b1", 128, (8, 10), 1, 16, 30, amp=1e-2, seed=2, r_scale=0.02)
    run_block("block_c64_cb", 64, (6, 10), 3, 12, 30, amp=2e-3, seed=3, r_scale=0.02, enable_CB=True)
    run_block("block_c64_dense", 64, (6, 10), 2, 12, 20, amp=2e-4, seed=4, r_scale=1.0)
    run_block("block_c256_w3x5", 256, (3, 5), 2, 6, 10, amp=2e-3, seed=5, r_scale=0.02)
    gen_backbone()
    gen_yolox()


if __name__ == "__main__":
    main()
