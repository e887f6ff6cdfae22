"""CPU oracle for the SAST hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (no nn.Module) restatement, on torch-CPU fp32, of the algorithm in
the reference's ``models/layers/SAST/SAST.py``, ``models/layers/SAST/ops.py``,
``models/detection/recurrent_backbone/sast_rnn.py`` and ``models/layers/rnn.py``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product package
``sast_b200`` never does.

Pinning: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference itself, produced in the build
container by ``oracle/gen_golden.py`` (which imports /root/reference unmodified)
and committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
every function here against those fixtures.  Below the torch boundary
(``F.layer_norm``, ``softmax``, ``erf`` GELU, ``topk``) arithmetic is torch's own.

All weights come in as a flat ``dict[str, Tensor]`` using the reference's
state-dict keys, relative to the object being evaluated.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


def sub(params: Params, prefix: str) -> Params:
    """Sub-dictionary of ``params`` under ``prefix.`` with the prefix stripped."""
    pl = prefix + "."
    return {k[len(pl):]: v for k, v in params.items() if k.startswith(pl)}


# --------------------------------------------------------------------------- #
# a1  scene sparsity ratio r            (ref: sast_rnn.py:45-60)
# --------------------------------------------------------------------------- #
def non_zero_ratio(x: Tensor) -> Tensor:
    """[B,Cin,H,W] any dtype -> [B,4,Cin] fp32.

    Max-pool by 4 then three times by 2; count non-zero cells per (frame, bin)
    with int16 accumulation (wraps like the reference); multiply by the fp32-cast
    scalar B / numel(pooled)."""
    xf = x.float()
    levels = []
    cur = F.max_pool2d(xf, kernel_size=4, stride=4)
    for lvl in range(4):
        if lvl > 0:
            cur = F.max_pool2d(cur, kernel_size=2, stride=2)
        cnt = (cur != 0).sum(dim=2, dtype=torch.int16).sum(dim=-1, dtype=torch.int16)
        levels.append(x.shape[0] / cur.numel() * cnt.float())
    return torch.stack(levels, dim=1)


# --------------------------------------------------------------------------- #
# a2  sine position table               (ref: sast_rnn.py:180-219)
# --------------------------------------------------------------------------- #
def position_table(H: int, W: int, C: int, temperature: float = 10000.0) -> Tensor:
    """[H,W,C] fp32 table; first C/2 channels encode y, last C/2 encode x."""
    npf = C // 2
    scale = 2 * math.pi
    ones = torch.ones(1, H, W, dtype=torch.bool)
    y = ones.cumsum(1, dtype=torch.float32)
    xx = ones.cumsum(2, dtype=torch.float32)
    y = (y - 0.5) / (y[:, -1:, :] + 1e-6) * scale
    xx = (xx - 0.5) / (xx[:, :, -1:] + 1e-6) * scale
    k = torch.arange(npf, dtype=torch.float32)
    div = temperature ** (2 * (k // 2) / npf)
    py = y[..., None] / div
    px = xx[..., None] / div
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3)[0]


# --------------------------------------------------------------------------- #
# a7  partitions                        (ref: ops.py:189-220)
# --------------------------------------------------------------------------- #
def window_partition(x: Tensor, p: Tuple[int, int]) -> Tensor:
    B, H, W, C = x.shape
    assert H % p[0] == 0 and W % p[1] == 0
    t = x.reshape(B, H // p[0], p[0], W // p[1], p[1], C).permute(0, 1, 3, 2, 4, 5)
    return t.reshape(-1, p[0], p[1], C)


def window_reverse(w: Tensor, p: Tuple[int, int], hw: Tuple[int, int]) -> Tensor:
    H, W = hw
    C = w.shape[-1]
    t = w.reshape(-1, H // p[0], W // p[1], p[0], p[1], C).permute(0, 1, 3, 2, 4, 5)
    return t.reshape(-1, H, W, C)


def grid_partition(x: Tensor, g: Tuple[int, int]) -> Tensor:
    B, H, W, C = x.shape
    assert H % g[0] == 0 and W % g[1] == 0
    t = x.reshape(B, g[0], H // g[0], g[1], W // g[1], C).permute(0, 2, 4, 1, 3, 5)
    return t.reshape(-1, g[0], g[1], C)


def grid_reverse(w: Tensor, g: Tuple[int, int], hw: Tuple[int, int]) -> Tensor:
    H, W = hw
    C = w.shape[-1]
    t = w.reshape(-1, H // g[0], W // g[1], g[0], g[1], C).permute(0, 3, 1, 4, 2, 5)
    return t.reshape(-1, H, W, C)


# --------------------------------------------------------------------------- #
# a5/a6  selection                      (ref: SAST.py:258-281)
# --------------------------------------------------------------------------- #
def select_windows_from_probs(prob: Tensor, d: float, b: float) -> Tensor:
    """prob [B,N] fp32 -> ascending flat ids b*N+n (int64) with prob >= d/(1+b).

    The Python-double threshold is cast to fp32 by torch before the compare."""
    keep = prob >= d / (1 + b)
    ij = torch.nonzero(keep)
    return ij[:, 0] * prob.shape[-1] + ij[:, 1]


def select_tokens_from_probs(prob: Tensor, d: float, b: float) -> Tuple[Tensor, Tensor, Tensor]:
    """prob [M,T] fp32 -> (index_token [M*Kmax], asy_index [S], K [M]).

    ``index_token`` = per-row top-Kmax (unsorted) + row offset; ``asy_index`` = the
    truly selected positions, ascending, in the compacted [M*T] space."""
    keep = prob >= d / (1 + b)
    K = keep.sum(dim=1)
    top = torch.topk(prob, k=int(K.max()), dim=1, largest=True, sorted=False)[1]
    base = torch.arange(0, prob.shape[0] * prob.shape[1], prob.shape[1]).view(-1, 1)
    ij = torch.nonzero(keep)
    return (top + base).reshape(-1), ij[:, 0] * prob.shape[-1] + ij[:, 1], K


def window_probs(scores: Tensor, T: int) -> Tensor:
    """scores [B,N,T,C] -> softmax_N( sum_{t,c}|s| / T )   (ref: SAST.py:84-89)."""
    return (torch.norm(scores, dim=[2, 3], p=1) / T).softmax(-1)


def token_probs(scores: Tensor, index_window: Tensor) -> Tensor:
    """scores [B,N,T,C] -> softmax_T( sum_c|s| ) of the selected windows  (ref: SAST.py:91-96)."""
    B, N = scores.shape[:2]
    return torch.norm(scores, dim=[3], p=1).view(B * N, -1)[index_window].softmax(-1)


def select_layer(scores: Tensor, T: int, bounce: float) -> List[Tensor]:
    """One layer's selection -> [index_window, index_token, padding_index, asy_index, K]."""
    B, N = scores.shape[:2]
    iw = select_windows_from_probs(window_probs(scores, T).view(B, N), 1 / N, bounce)
    it, asy, K = select_tokens_from_probs(token_probs(scores, iw), 1 / T, bounce)
    pad = it[torch.isin(it, asy, assume_unique=True, invert=True)]
    return [iw, it, pad, asy, K]


# --------------------------------------------------------------------------- #
# a8-a13  one MS-WSA layer, sparse form  (ref: SAST.py:199-255)
# --------------------------------------------------------------------------- #
def _ln(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def glu_mlp(x: Tensor, p: Params) -> Tensor:
    """GLU(C->2I, value*gelu_erf(gate)) -> Linear(I->C)   (ref: ops.py:135-137,165-175)."""
    val, gate = torch.tensor_split(F.linear(x, p["net.0.proj.weight"], p.get("net.0.proj.bias")), 2, dim=-1)
    return F.linear(val * F.gelu(gate), p["net.2.weight"], p.get("net.2.bias"))


def ms_wsa(x: Tensor, index_window: Tensor, index_token: Tensor, padding_index: Tensor,
           asy_index: Tensor, M: int, B: int, enable_CB: bool, p: Params,
           dim_head: int = 32, eps: float = 1e-5) -> Tensor:
    """x [B*N,T,C] (partitioned) -> same shape, following the reference's gather ->
    padded attention -> un-pad -> MLP -> scatter sequence step by step."""
    shape = x.shape
    C = shape[-1]
    heads = C // dim_head
    x = _ln(x.reshape(shape[0], -1, C), p["norm1.weight"], p["norm1.bias"], eps)
    if len(index_token) == 0:
        return x.reshape(shape)
    full = x.clone()                                   # every token, norm1'd
    win = x[index_window].reshape(-1, C).clone()       # tokens of the selected windows
    n2 = _ln(win[asy_index], p["norm2.weight"], p["norm2.bias"], eps)
    work = win.clone()
    work[asy_index] = n2
    tile = work[index_token].reshape(M, -1, C)         # [M,Kmax,C], padded
    Kmax = tile.shape[1]

    qkv = F.linear(tile, p["qkv.weight"], p.get("qkv.bias"))
    q, k, v = qkv.reshape(M, Kmax, heads, 3 * dim_head).transpose(1, 2).chunk(3, dim=3)
    att = (q @ k.transpose(-2, -1)) * dim_head ** -0.5  # [M,h,Kq,Kk]
    # key columns that are padding get -1e4 (ref: SAST.py:223-226)
    is_pad = torch.zeros(win.shape[0], dtype=torch.bool)
    is_pad[padding_index] = True
    col_pad = is_pad[index_token].reshape(M, 1, 1, Kmax)
    att = torch.where(col_pad, torch.full_like(att, -1e4), att)
    o = (att.softmax(dim=-1) @ v).transpose(1, 2).reshape(M, Kmax, C)
    o = F.linear(o, p["proj.weight"], p.get("proj.bias"))

    scratch = win.clone()
    scratch[index_token] = o.reshape(-1, C)
    o_sel = scratch[asy_index]
    y = n2 + o_sel * p["ls1.gamma"] if "ls1.gamma" in p else n2 + o_sel
    m = glu_mlp(y, sub(p, "mlp"))
    if enable_CB:                                      # (ref: SAST.py:240-246)
        tw = torch.zeros_like(win)
        tw[asy_index] = m
        tf = torch.zeros_like(full)
        tf[index_window] = tw.reshape(M, -1, C)
        tf = tf.reshape(B, -1, C)
        tf = (0.5 * tf + 0.5 * tf.mean(dim=1, keepdim=True)).reshape(full.shape)
        m = tf[index_window].reshape(-1, C)[asy_index]
    out = y + m * p["ls2.gamma"] if "ls2.gamma" in p else y + m

    win_out = win.clone()                              # unselected + padding keep norm1(x)
    win_out[asy_index] = out
    full[index_window] = win_out.reshape(M, -1, C)
    return full.reshape(shape)


def ms_wsa_dense(x: Tensor, sel: Tensor, p: Params, B: int, enable_CB: bool = False,
                 dim_head: int = 32, eps: float = 1e-5) -> Tensor:
    """Dense-equivalent form: x [B*N,T,C], sel [B*N,T] bool.
    Independent cross-check of :func:`ms_wsa` -- no index lists, keys masked with -inf."""
    Wn, T, C = x.shape
    heads = C // dim_head
    n1 = _ln(x, p["norm1.weight"], p["norm1.bias"], eps)
    n2 = _ln(n1, p["norm2.weight"], p["norm2.bias"], eps)
    qkv = F.linear(n2, p["qkv.weight"], p.get("qkv.bias"))
    q, k, v = qkv.reshape(Wn, T, heads, 3 * dim_head).transpose(1, 2).chunk(3, dim=3)
    att = (q @ k.transpose(-2, -1)) * dim_head ** -0.5
    att = att.masked_fill(~sel[:, None, None, :], float("-inf"))
    att = torch.nan_to_num(att.softmax(dim=-1), nan=0.0)
    o = F.linear((att @ v).transpose(1, 2).reshape(Wn, T, C), p["proj.weight"], p.get("proj.bias"))
    y = n2 + o * p.get("ls1.gamma", torch.ones(C))
    m = glu_mlp(y, sub(p, "mlp"))
    if enable_CB:
        ms = torch.where(sel[..., None], m, torch.zeros_like(m)).reshape(B, -1, C)
        m = (0.5 * ms + 0.5 * ms.mean(dim=1, keepdim=True)).reshape(Wn, T, C)
    out = y + m * p.get("ls2.gamma", torch.ones(C))
    return torch.where(sel[..., None], out, n1)


# --------------------------------------------------------------------------- #
# a4 + a14  the SAST block               (ref: SAST.py:98-160)
# --------------------------------------------------------------------------- #
def scoring(x: Tensor, pos: Tensor, r: Tensor, p: Params, part: Tuple[int, int],
            amp: float) -> Tuple[Tensor, Tensor]:
    """x [B,H,W,C], pos broadcastable to x, r [B,20] ->
    (STP-weighted x [B*N,T,C], scaled scores [B,N,T,C])   (ref: SAST.py:105-119)."""
    B, H, W, C = x.shape
    T = part[0] * part[1]
    N = H * W // T
    xw = window_partition(x + pos, part).reshape(B, N, T, C)
    ctrl = F.linear(r + 1e-6, torch.exp(p["to_controls.weight"]))[:, None, None, :]
    s = F.relu(F.linear(xw, p["to_scores.weight"], p["to_scores.bias"]))
    xw = (ctrl.sigmoid() * s.sigmoid() * xw).reshape(B * N, T, C)
    inv = amp / ctrl
    inv[inv == torch.inf] = 0
    return xw, inv * s


def sast_block(x: Tensor, pos: Tensor, r: Tensor, p: Params, part: Tuple[int, int],
               first_block: bool = True, index_list: Optional[List] = None,
               amp: float = 2e-4, bounce: float = 1e-3, enable_CB: bool = False,
               dim_head: int = 32, eps: float = 1e-5):
    """x [B,H,W,C] -> (x [B,H,W,C], index_count, [list1, list2])."""
    B, H, W, C = x.shape
    T = part[0] * part[1]
    N = H * W // T
    if first_block:
        xw, scores = scoring(x, pos, r, p, part, amp)
        l1 = select_layer(scores, T, bounce)
    else:
        xw = window_partition(x + pos, part).reshape(B * N, T, C)
        l1, l2 = index_list
    iw, it, pad, asy, K = l1
    if len(it):
        xw = ms_wsa(xw, iw, it, pad, asy, len(iw), B, enable_CB, sub(p, "win_attn"), dim_head, eps)
    x = window_reverse(xw, part, (H, W))
    count = len(asy) // B

    if first_block:
        s_map = window_reverse(scores.reshape(B * N, T, C), part, (H, W))
        l2 = select_layer(grid_partition(s_map, part).reshape(B, N, T, C), T, bounce)
    iw, it, pad, asy, K = l2
    xg = grid_partition(x, part).reshape(B * N, T, C)
    if len(it):
        xg = ms_wsa(xg, iw, it, pad, asy, len(iw), B, enable_CB, sub(p, "grid_attn"), dim_head, eps)
    x = grid_reverse(xg, part, (H, W))
    count += len(asy) // B
    return x, count, [l1, l2]


# --------------------------------------------------------------------------- #
# callers on either side of the block (SURVEY.md section 8f "next" rows)
# --------------------------------------------------------------------------- #
def conv_downsample(x: Tensor, p: Params, factor: int) -> Tensor:
    """NCHW -> NHWC: overlapping strided conv (replicate pad, no bias) + LayerNorm
    (ref: ops.py:54-91)."""
    k = (factor - 1) * 2 + 1
    pad = k // 2
    xp = F.pad(x, (pad, pad, pad, pad), mode="replicate")
    y = F.conv2d(xp, p["conv.weight"], None, stride=factor).permute(0, 2, 3, 1).contiguous()
    return _ln(y, p["norm.weight"], p["norm.bias"], 1e-5)


def conv_lstm(x: Tensor, state: Optional[Tuple[Tensor, Tensor]], p: Params) -> Tuple[Tensor, Tensor]:
    """1x1-conv LSTM cell, NCHW (ref: models/layers/rnn.py:36-69, dws_conv False)."""
    C = x.shape[1]
    h0, c0 = state if state is not None else (torch.zeros_like(x), torch.zeros_like(x))
    mix = F.conv2d(torch.cat((x, h0), dim=1), p["conv1x1.weight"], p["conv1x1.bias"])
    f, i, o = torch.sigmoid(mix[:, :3 * C]).tensor_split(3, dim=1)
    g = torch.tanh(mix[:, 3 * C:])
    c1 = f * c0 + i * g
    return o * torch.tanh(c1), c1


def backbone_forward(x: Tensor, prev_states: Optional[Sequence], params: Params, cfg: dict,
                     token_mask: Optional[Tensor] = None):
    """RNNDetector.forward (ref: sast_rnn.py:144-162, 265-287).

    cfg keys: embed_dim, dim_multiplier, num_blocks, patch_size, in_res_hw,
    partition_size, and optional AMP / BOUNCE / enable_CB / dim_head / norm_eps."""
    n_stage = len(cfg["num_blocks"])
    prev_states = list(prev_states) if prev_states is not None else [None] * n_stage
    r = non_zero_ratio(x)
    x = x.float()
    part = tuple(cfg["partition_size"])
    res = list(cfg["in_res_hw"])
    feats, states, counts = {}, [], []
    for s in range(n_stage):
        factor = cfg["patch_size"] if s == 0 else 2
        res = [res[0] // factor, res[1] // factor]
        sp = sub(params, f"stages.{s}")
        C = cfg["embed_dim"] * cfg["dim_multiplier"][s]
        y = conv_downsample(x, sub(sp, "downsample_cf2cl"), factor)
        if s == 0 and token_mask is not None:
            y[token_mask] = sp["mask_token"]
        Hs, Ws = y.shape[1:3]
        pos = position_table(res[0], res[1], C)[:Hs, :Ws]
        P, il, rr = 0, None, r[:, s]
        for bi in range(cfg["num_blocks"][s]):
            y, cnt, il = sast_block(y, pos, rr, sub(sp, f"att_blocks.{bi}.att"), part,
                                    first_block=(bi == 0), index_list=il,
                                    amp=cfg.get("AMP", 2e-4), bounce=cfg.get("BOUNCE", 1e-3),
                                    enable_CB=cfg.get("enable_CB", False),
                                    dim_head=cfg.get("dim_head", 32), eps=cfg.get("norm_eps", 1e-5))
            P += cnt
        h, c = conv_lstm(y.permute(0, 3, 1, 2).contiguous(), prev_states[s], sub(sp, "lstm"))
        x = h
        feats[s + 1] = h
        states.append((h, c))
        counts.append(P)
    return feats, states, counts
