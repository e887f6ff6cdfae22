"""Helpers shared by the -m gpu parity tests: build product modules from golden metadata."""
import torch

import sast_b200
from sast_b200 import _lib as L
from sast_b200.backbone import PositionEmbeddingSine
from sast_b200.config import attention_config, backbone_config
from oracle.golden_common import make_params, with_aliases

DEV = "cuda:0"


def build_block(meta, precision, first_block=True, shapes_key="shapes", seed_key="seed"):
    blk = sast_b200.SAST_block(meta["C"], attention_config(meta["part"], AMP=meta["AMP"], BOUNCE=meta["BOUNCE"],
                                                            enable_CB=meta["enable_CB"]), first_block=first_block)
    params = make_params(meta[shapes_key], seed=meta[seed_key])
    blk.load_state_dict(with_aliases(params, blk.state_dict().keys()), strict=True)
    blk = blk.to(DEV).eval()
    for m in (blk.win_attn, blk.grid_attn):
        m.precision = precision
    return blk, params


def pos_module(meta):
    return PositionEmbeddingSine(meta["C"] // 2, normalize=True, input_size=(1, meta["H"], meta["W"]))


def build_backbone(meta, precision):
    net = sast_b200.build_recurrent_backbone(backbone_config(
        meta["in_res_hw"], embed_dim=meta["embed_dim"], num_blocks=meta["num_blocks"],
        enable_masking=meta["enable_masking"], AMP=meta["AMP"], BOUNCE=meta["BOUNCE"], enable_CB=meta["enable_CB"]))
    params = make_params(meta["shapes"], seed=meta["seed"])
    net.load_state_dict(with_aliases(params, net.state_dict().keys()), strict=True)
    net = net.to(DEV).eval()
    set_precision(net, precision)
    return net, params


def set_precision(net, precision):
    for m in net.modules():
        if isinstance(m, (sast_b200.MS_WSA, sast_b200.DWSConvLSTM2d)):
            m.precision = precision


def sel_mask(iw, asy, NW, T):
    """[NW,T] bool mask from reference-style (index_window, asy_index)."""
    m = torch.zeros(NW * T, dtype=torch.bool)
    if len(asy):
        m[iw[asy // T] * T + asy % T] = True
    return m.view(NW, T)
