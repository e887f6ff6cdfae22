"""YOLOX PAFPN + decoupled head + post-processing, inference path.

The detection neck and head that consume the recurrent backbone's stage outputs.  Same module tree
and state-dict keys as the reference (models/detection/yolox_extension/models/yolo_pafpn.py:19-139,
models/detection/yolox/models/network_blocks.py:28-141, yolo_head.py:21-289, utils/boxes.py:32-76,
yolox_extension/models/detector.py:19-72), so a checkpoint of the reference detector loads with
``strict=True``; the arithmetic is dense convolutions, which stay cuDNN calls.  What is done for the
GPU here is layout and folding, not kernels:

* the backbone hands over NCHW-logical tensors in channels-last memory; every conv of neck and head is
  converted to channels-last once (``prepare_inference``), so no layout copies happen between them;
* ``prepare_inference`` also folds every BatchNorm into its conv (eval-mode statistics), turning
  conv + BN + SiLU into conv + SiLU;
* the head writes reg / obj / cls of the three levels straight into one ``[B, A, 5 + classes]`` buffer and
  decodes it in place with cached grids;
* ``postprocess`` keeps everything on the device (score threshold, class-aware NMS through offset boxes).

Training (SimOTA assignment and the losses of yolo_head.py:291-606) is out of scope: the head raises if
called with labels.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .backbone import build_recurrent_backbone

Tensor = torch.Tensor


def _activation(name: str) -> nn.Module:
    if name == "silu":
        return nn.SiLU(inplace=True)
    if name == "relu":
        return nn.ReLU(inplace=True)
    if name == "lrelu":
        return nn.LeakyReLU(0.1, inplace=True)
    raise AttributeError(f"Unsupported act type: {name}")


class BaseConv(nn.Module):
    """conv (same padding, no bias) -> BatchNorm -> activation   (ref: network_blocks.py:28-55)."""

    def __init__(self, in_channels: int, out_channels: int, ksize: int, stride: int, groups: int = 1, bias: bool = False,
                 act: str = "silu"):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, ksize, stride, (ksize - 1) // 2, groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels)
        self.act = _activation(act)
        self._folded = False

    def forward(self, x: Tensor) -> Tensor:
        if self._folded and not self.training:
            c = self.conv
            return self.act(F.conv2d(x, self._fw, self._fb, c.stride, c.padding, c.dilation, c.groups))
        return self.act(self.bn(self.conv(x)))

    @torch.no_grad()
    def fold_bn(self) -> None:
        """Eval-mode BN folded into the conv: w' = w * g / sqrt(var + eps), b' = beta + (b - mean) * g / sqrt(var + eps).
        The folded weight / bias live in NON-persistent buffers used only by the inference forward: parameters and state
        dict stay exactly the reference's (a later strict load_state_dict works), and both ``train()`` and a state-dict load
        drop the fold (call ``prepare_inference`` / ``fold_bn`` again afterwards)."""
        bn, conv = self.bn, self.conv
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        bias0 = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
        self.register_buffer("_fw", (conv.weight * scale.view(-1, 1, 1, 1)).contiguous(memory_format=torch.channels_last), persistent=False)
        self.register_buffer("_fb", (bn.bias + (bias0 - bn.running_mean) * scale).contiguous(), persistent=False)
        self._folded = True

    def _unfold(self) -> None:
        if self._folded:
            self._folded = False
            self._fw = self._fb = None

    def train(self, mode: bool = True):
        if mode:
            self._unfold()
        return super().train(mode)

    def _load_from_state_dict(self, *args, **kwargs):
        self._unfold()                              # new statistics / weights: the fold is stale
        return super()._load_from_state_dict(*args, **kwargs)


class DWConv(nn.Module):
    """depthwise k x k conv + pointwise conv   (ref: network_blocks.py:58-77)."""

    def __init__(self, in_channels: int, out_channels: int, ksize: int, stride: int = 1, act: str = "silu"):
        super().__init__()
        self.dconv = BaseConv(in_channels, in_channels, ksize, stride, groups=in_channels, act=act)
        self.pconv = BaseConv(in_channels, out_channels, 1, 1, act=act)

    def forward(self, x: Tensor) -> Tensor:
        return self.pconv(self.dconv(x))


class Bottleneck(nn.Module):
    """1x1 -> 3x3 with an optional identity shortcut   (ref: network_blocks.py:80-103)."""

    def __init__(self, in_channels: int, out_channels: int, shortcut: bool = True, expansion: float = 0.5,
                 depthwise: bool = False, act: str = "silu"):
        super().__init__()
        hidden = int(out_channels * expansion)
        self.conv1 = BaseConv(in_channels, hidden, 1, 1, act=act)
        self.conv2 = (DWConv if depthwise else BaseConv)(hidden, out_channels, 3, 1, act=act)
        self.use_add = shortcut and in_channels == out_channels

    def forward(self, x: Tensor) -> Tensor:
        y = self.conv2(self.conv1(x))
        return y + x if self.use_add else y


class CSPLayer(nn.Module):
    """CSP bottleneck with three 1x1 convs around n Bottlenecks   (ref: network_blocks.py:106-141)."""

    def __init__(self, in_channels: int, out_channels: int, n: int = 1, shortcut: bool = True, expansion: float = 0.5,
                 depthwise: bool = False, act: str = "silu"):
        super().__init__()
        hidden = int(out_channels * expansion)
        self.conv1 = BaseConv(in_channels, hidden, 1, 1, act=act)
        self.conv2 = BaseConv(in_channels, hidden, 1, 1, act=act)
        self.conv3 = BaseConv(2 * hidden, out_channels, 1, 1, act=act)
        self.m = nn.Sequential(*[Bottleneck(hidden, hidden, shortcut, 1.0, depthwise, act=act) for _ in range(n)])

    def forward(self, x: Tensor) -> Tensor:
        return self.conv3(torch.cat((self.m(self.conv1(x)), self.conv2(x)), dim=1))


class YOLOPAFPN(nn.Module):
    """Top-down + bottom-up path aggregation over three backbone stages   (ref: yolo_pafpn.py:19-139)."""

    def __init__(self, depth: float = 1.0, in_stages: Sequence[int] = (2, 3, 4), in_channels: Sequence[int] = (256, 512, 1024),
                 depthwise: bool = False, act: str = "silu", compile_cfg: Optional[Dict] = None):
        super().__init__()
        assert len(in_stages) == len(in_channels) == 3, "three feature maps"
        if compile_cfg is not None and compile_cfg.get("enable", False):
            raise NotImplementedError("torch.compile of the neck is not part of this build")
        self.in_features = tuple(in_stages)
        self.in_channels = tuple(in_channels)
        c0, c1, c2 = self.in_channels
        n = round(3 * depth)
        Conv = DWConv if depthwise else BaseConv
        self.lateral_conv0 = BaseConv(c2, c1, 1, 1, act=act)
        self.C3_p4 = CSPLayer(2 * c1, c1, n, False, depthwise=depthwise, act=act)
        self.reduce_conv1 = BaseConv(c1, c0, 1, 1, act=act)
        self.C3_p3 = CSPLayer(2 * c0, c0, n, False, depthwise=depthwise, act=act)
        self.bu_conv2 = Conv(c0, c0, 3, 2, act=act)
        self.C3_n3 = CSPLayer(2 * c0, c1, n, False, depthwise=depthwise, act=act)
        self.bu_conv1 = Conv(c1, c1, 3, 2, act=act)
        self.C3_n4 = CSPLayer(2 * c1, c2, n, False, depthwise=depthwise, act=act)

    @staticmethod
    def upsample(x: Tensor) -> Tensor:
        return F.interpolate(x, scale_factor=2, mode="nearest-exact")

    def forward(self, features: Dict[int, Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
        x2, x1, x0 = (features[f] for f in self.in_features)          # strides 8, 16, 32
        top = self.lateral_conv0(x0)
        mid = self.C3_p4(torch.cat([self.upsample(top), x1], 1))
        mid_r = self.reduce_conv1(mid)
        out2 = self.C3_p3(torch.cat([self.upsample(mid_r), x2], 1))
        out1 = self.C3_n3(torch.cat([self.bu_conv2(out2), mid_r], 1))
        out0 = self.C3_n4(torch.cat([self.bu_conv1(out1), top], 1))
        return out2, out1, out0


class YOLOXHead(nn.Module):
    """Decoupled head: per level a 1x1 stem, two 3x3 convs per branch, 1x1 predictors; outputs
    ``[B, A, 5 + classes]`` = (cx, cy, w, h, objectness, class scores), decoded to input pixels
    (ref: yolo_head.py:21-153 constructor, :165-243 forward, :268-289 decode)."""

    def __init__(self, num_classes: int = 80, strides: Sequence[int] = (8, 16, 32), in_channels: Sequence[int] = (256, 512, 1024),
                 act: str = "silu", depthwise: bool = False, compile_cfg: Optional[Dict] = None):
        super().__init__()
        if compile_cfg is not None and compile_cfg.get("enable", False):
            raise NotImplementedError("torch.compile of the head is not part of this build")
        self.num_classes = num_classes
        self.decode_in_inference = True
        self.strides = tuple(strides)
        hidden = int(256 * (in_channels[-1] / 1024))                  # width follows the widest input: in[-1] / 4
        Conv = DWConv if depthwise else BaseConv
        self.cls_convs, self.reg_convs = nn.ModuleList(), nn.ModuleList()
        self.cls_preds, self.reg_preds, self.obj_preds = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.stems = nn.ModuleList()
        for c in in_channels:
            self.stems.append(BaseConv(c, hidden, 1, 1, act=act))
            self.cls_convs.append(nn.Sequential(Conv(hidden, hidden, 3, 1, act=act), Conv(hidden, hidden, 3, 1, act=act)))
            self.reg_convs.append(nn.Sequential(Conv(hidden, hidden, 3, 1, act=act), Conv(hidden, hidden, 3, 1, act=act)))
            self.cls_preds.append(nn.Conv2d(hidden, num_classes, 1))
            self.reg_preds.append(nn.Conv2d(hidden, 4, 1))
            self.obj_preds.append(nn.Conv2d(hidden, 1, 1))
        prior = -math.log((1 - 0.01) / 0.01)                           # focal-loss prior on the two sigmoid outputs
        for conv in list(self.cls_preds) + list(self.obj_preds):
            nn.init.constant_(conv.bias, prior)
        self.hw: List[Tuple[int, int]] = []
        self._grid_key = None
        self._grid: Optional[Tensor] = None
        self._stride: Optional[Tensor] = None

    def _grids(self, hw: List[Tuple[int, int]], device, dtype) -> Tuple[Tensor, Tensor]:
        key = (tuple(hw), device, dtype)
        if self._grid_key != key:
            grids, strides = [], []
            for (h, w), s in zip(hw, self.strides):
                yv, xv = torch.meshgrid(torch.arange(h, device=device, dtype=dtype), torch.arange(w, device=device, dtype=dtype),
                                        indexing="ij")
                grids.append(torch.stack((xv, yv), 2).view(1, -1, 2))
                strides.append(torch.full((1, h * w, 1), s, device=device, dtype=dtype))
            self._grid, self._stride, self._grid_key = torch.cat(grids, 1), torch.cat(strides, 1), key
        return self._grid, self._stride

    def forward(self, xin: Sequence[Tensor], labels=None):
        if labels is not None or self.training:
            raise NotImplementedError("sast_b200.yolox implements the inference path; the YOLOX losses are out of scope")
        B = xin[0].shape[0]
        self.hw = [tuple(x.shape[-2:]) for x in xin]
        A = sum(h * w for h, w in self.hw)
        out = xin[0].new_empty(B, A, 5 + self.num_classes)
        a0 = 0
        for k, x in enumerate(xin):
            x = self.stems[k](x)
            cls_feat = self.cls_convs[k](x)
            reg_feat = self.reg_convs[k](x)
            n = x.shape[-2] * x.shape[-1]
            lvl = out[:, a0:a0 + n]
            lvl[..., 0:4] = self.reg_preds[k](reg_feat).flatten(2).transpose(1, 2)
            lvl[..., 4:5] = self.obj_preds[k](reg_feat).sigmoid().flatten(2).transpose(1, 2)
            lvl[..., 5:] = self.cls_preds[k](cls_feat).sigmoid().flatten(2).transpose(1, 2)
            a0 += n
        if self.decode_in_inference:
            grid, stride = self._grids(self.hw, out.device, out.dtype)
            out[..., 0:2] = (out[..., 0:2] + grid) * stride
            out[..., 2:4] = torch.exp(out[..., 2:4]) * stride
        return out, None


def _nms(boxes: Tensor, scores: Tensor, thr: float) -> Tensor:
    """Greedy NMS, indices in descending score order.  torchvision's kernel when it is importable."""
    try:
        import torchvision
        return torchvision.ops.nms(boxes, scores, thr)
    except Exception:  # pragma: no cover - image without torchvision
        order = scores.argsort(descending=True)
        b = boxes[order]
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        lt = torch.max(b[:, None, :2], b[None, :, :2])
        rb = torch.min(b[:, None, 2:], b[None, :, 2:])
        inter = (rb - lt).clamp(min=0).prod(2)
        iou = inter / (area[:, None] + area[None, :] - inter)
        keep = []
        alive = torch.ones(len(b), dtype=torch.bool, device=b.device)
        for i in range(len(b)):
            if alive[i]:
                keep.append(i)
                alive &= ~(iou[i] > thr)
        return order[torch.tensor(keep, dtype=torch.long, device=b.device)]


def postprocess(prediction: Tensor, num_classes: int, conf_thre: float = 0.7, nms_thre: float = 0.45,
                class_agnostic: bool = False) -> List[Optional[Tensor]]:
    """[B, A, 5 + classes] (cx, cy, w, h, obj, cls...) -> per image ``[n, 7]`` rows
    (x1, y1, x2, y2, obj_conf, class_conf, class_id) in descending score order, or None
    (ref: utils/boxes.py:32-76; unlike the reference the input tensor is not modified in place)."""
    xy, wh = prediction[..., 0:2], prediction[..., 2:4]
    boxes = torch.cat((xy - wh / 2, xy + wh / 2), dim=-1)
    class_conf, class_id = prediction[..., 5:5 + num_classes].max(dim=-1)
    score = prediction[..., 4] * class_conf
    out: List[Optional[Tensor]] = [None] * prediction.shape[0]
    for i in range(prediction.shape[0]):
        keep = score[i] >= conf_thre
        if not bool(keep.any()):
            continue
        b, s, c = boxes[i][keep], score[i][keep], class_id[i][keep]
        if class_agnostic:
            idx = _nms(b, s, nms_thre)
        else:                                   # class-aware: boxes of different classes are moved apart, one NMS call
            span = b.max() + 1 if b.numel() else b.new_tensor(1.0)
            idx = _nms(b + (c.to(b.dtype) * span).view(-1, 1), s, nms_thre)
        det = torch.cat((b, prediction[i][keep][:, 4:5], class_conf[i][keep].unsqueeze(1), c.to(b.dtype).unsqueeze(1)), dim=1)
        out[i] = det[idx]
    return out


def build_yolox_fpn(fpn_cfg, in_channels: Sequence[int]) -> YOLOPAFPN:
    """(ref: yolox_extension/models/build.py:20-28)"""
    cfg = dict(fpn_cfg)
    name = cfg.pop("name")
    if name not in {"PAFPN", "pafpn"}:
        raise NotImplementedError(name)
    compile_cfg = cfg.pop("compile", None)
    return YOLOPAFPN(in_channels=tuple(in_channels), compile_cfg=compile_cfg, **cfg)


def build_yolox_head(head_cfg, in_channels: Sequence[int], strides: Sequence[int]) -> YOLOXHead:
    """(ref: yolox_extension/models/build.py:9-17)"""
    cfg = dict(head_cfg)
    cfg.pop("name", None)
    cfg.pop("version", None)
    compile_cfg = cfg.pop("compile", None)
    return YOLOXHead(in_channels=tuple(in_channels), strides=tuple(strides), compile_cfg=compile_cfg, **cfg)


class YoloXDetector(nn.Module):
    """Recurrent SAST backbone + PAFPN + YOLOX head   (ref: yolox_extension/models/detector.py:19-72)."""

    def __init__(self, model_cfg):
        super().__init__()
        self.backbone = build_recurrent_backbone(model_cfg.backbone)
        in_channels = self.backbone.get_stage_dims(tuple(model_cfg.fpn.in_stages))
        self.fpn = build_yolox_fpn(model_cfg.fpn, in_channels=in_channels)
        strides = self.backbone.get_strides(tuple(model_cfg.fpn.in_stages))
        self.yolox_head = build_yolox_head(model_cfg.head, in_channels=in_channels, strides=strides)

    def forward_backbone(self, x: Tensor, previous_states=None, token_mask: Optional[Tensor] = None):
        return self.backbone(x, previous_states, token_mask)

    def forward_detect(self, backbone_features: Dict[int, Tensor], targets: Optional[Tensor] = None):
        if targets is not None:
            raise NotImplementedError("training of the detection head is out of scope")
        return self.yolox_head(self.fpn(backbone_features))

    def forward(self, x: Tensor, previous_states=None, retrieve_detections: bool = True, targets: Optional[Tensor] = None):
        features, states, p = self.forward_backbone(x, previous_states)
        if not retrieve_detections:
            assert targets is None
            return None, None, states
        outputs, losses = self.forward_detect(features, targets)
        return outputs, losses, states, p

    @torch.no_grad()
    def prepare_inference(self) -> "YoloXDetector":
        """eval(), BatchNorm folded into the convs of neck and head, channels-last weights (the backbone's
        stage outputs already are channels-last memory, so no layout copy happens on the way)."""
        self.eval()
        for m in list(self.fpn.modules()) + list(self.yolox_head.modules()):
            if isinstance(m, BaseConv):
                m.fold_bn()
        self.fpn.to(memory_format=torch.channels_last)
        self.yolox_head.to(memory_format=torch.channels_last)
        return self


def detector_config(in_res_hw, num_classes: int, embed_dim: int = 64, partition_split_32: int = 2, fpn_depth: float = 0.67):
    """``model`` of config/model/sast_yolox/default.yaml as config/modifier.py completes it."""
    from .config import Config, backbone_config
    return Config(dict(backbone=dict(backbone_config(in_res_hw, embed_dim=embed_dim, partition_split_32=partition_split_32)),
                       fpn=dict(name="PAFPN", depth=fpn_depth, in_stages=[2, 3, 4], depthwise=False, act="silu"),
                       head=dict(name="YoloX", depthwise=False, act="silu", num_classes=num_classes),
                       postprocess=dict(confidence_threshold=0.01, nms_threshold=0.45)))
