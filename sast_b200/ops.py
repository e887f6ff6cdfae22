"""torch custom ops (`torch.ops.sast.*`) over the C ABI of libsast_b200.so.

Every op takes/returns fp32 NHWC CUDA tensors, launches on the current CUDA stream, never
synchronises with the host and has no CPU implementation: calling one with CPU tensors raises.
Selections live on the device inside an int32 "pool" tensor (see ``sast_selection`` in
include/sast_b200.h); :class:`Selection` binds a pool to the C struct and can lazily
materialise the reference's index lists when a caller really asks for them."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

Tensor = torch.Tensor


def _geom(B, H, W, Cc, p0, p1) -> L.Geom:
    return L.Geom(int(B), int(H), int(W), int(Cc), int(p0), int(p1))


def _f32c(t: Tensor, name: str) -> Tensor:
    L.require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def thresholds(N: int, T: int, bounce: float) -> Tuple[float, float]:
    """fp32-cast thresholds exactly as torch evaluates ``x >= d / (1 + b)`` (SAST.py:262,272)."""
    return float(np.float32((1 / N) / (1 + bounce))), float(np.float32((1 / T) / (1 + bounce)))


# --------------------------------------------------------------------------------------------
# a1  non_zero_ratio
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("sast::nonzero_ratio", mutates_args=())
def nonzero_ratio(x: Tensor) -> Tensor:
    L.require_cuda(x, "x")
    if x.dtype == torch.uint8:
        dt = L.U8
    elif x.dtype == torch.int32:
        dt = L.I32
    elif x.dtype == torch.float32:
        dt = L.F32
    else:  # other integer / float types: one cast, same values
        x = x.to(torch.float32 if x.is_floating_point() else torch.int32)
        dt = L.F32 if x.is_floating_point() else L.I32
    x = x.contiguous()
    B, Cin, H, W = x.shape
    r = torch.empty(4, B, Cin, device=x.device, dtype=torch.float32)          # level-major: r[:, i] below is contiguous
    scratch = torch.zeros(B * Cin * 4, device=x.device, dtype=torch.int32)
    L.run(x.device, "sast_nonzero_ratio", x.data_ptr(), dt, B, Cin, H, W, r.data_ptr(), scratch.data_ptr())
    return r.permute(1, 0, 2)                                                   # the reference's [B, 4, Cin]


@nonzero_ratio.register_fake
def _(x):
    return x.new_empty(4, x.shape[0], x.shape[1], dtype=torch.float32).permute(1, 0, 2)


class PackedEvents:
    """Event histogram packed ``bits`` (1 or 4) bits per bin along x, little endian: ``data`` is uint8
    [B, Cin, H, W * bits // 8].  Binary histograms (benchmark.py:58-60) fit 1 bit, dataset histograms (clipped at
    count_cutoff = 10, data/utils/representations.py) 4 bits, so the host->device copy shrinks 8x / 2x without loss;
    ``RNNDetector.forward`` accepts it in place of the [B, Cin, H, W] tensor and unpacks on the device in the pass
    that computes the scene sparsity ratio.  Tensor-like enough for the runners (clone / copy_ / to / pin_memory)."""

    def __init__(self, data: Tensor, bits: int, width: int):
        assert bits in (1, 4) and data.dtype == torch.uint8 and data.dim() == 4
        assert width % 8 == 0 and data.shape[-1] == width * bits // 8
        self.data, self.bits, self.width = data, bits, width

    @property
    def shape(self):
        return tuple(self.data.shape[:3]) + (self.width,)

    @property
    def device(self):
        return self.data.device

    @property
    def is_cuda(self):
        return self.data.is_cuda

    def numel(self):
        """bytes actually held (what a host->device copy moves)"""
        return self.data.numel()

    def data_ptr(self):
        return self.data.data_ptr()

    def clone(self):
        return PackedEvents(self.data.clone(), self.bits, self.width)

    def to(self, *a, **k):
        return PackedEvents(self.data.to(*a, **k), self.bits, self.width)

    def pin_memory(self):
        return PackedEvents(self.data.pin_memory(), self.bits, self.width)

    def copy_(self, other, non_blocking=False):
        assert isinstance(other, PackedEvents) and (other.bits, other.width) == (self.bits, self.width)
        self.data.copy_(other.data, non_blocking=non_blocking)
        return self

    def unpack_reference(self) -> Tensor:
        """uint8 [B, Cin, H, W] by plain torch ops (tests; works on CPU)."""
        d = self.data.to(torch.int32)
        if self.bits == 1:
            out = torch.stack([(d >> k) & 1 for k in range(8)], dim=-1)
        else:
            out = torch.stack([d & 15, (d >> 4) & 15], dim=-1)
        return out.reshape(*self.data.shape[:3], self.width).to(torch.uint8)


def pack_events(x: Tensor, bits: int) -> PackedEvents:
    """[B, Cin, H, W] integer histogram (values < 2**bits, W % 8 == 0) -> :class:`PackedEvents` on x's device."""
    assert bits in (1, 4) and x.dim() == 4 and x.shape[-1] % 8 == 0
    assert int(x.max()) < (1 << bits) and int(x.min()) >= 0, f"values do not fit {bits} bit(s)"
    n = 8 // bits
    d = x.to(torch.int32).reshape(*x.shape[:3], x.shape[-1] // n, n)
    w = torch.zeros(d.shape[:-1], dtype=torch.int32, device=x.device)
    for k in range(n):
        w |= d[..., k] << (k * bits)
    return PackedEvents(w.to(torch.uint8).contiguous(), bits, x.shape[-1])


@torch.library.custom_op("sast::unpack_nonzero_ratio", mutates_args=())
def unpack_nonzero_ratio(data: Tensor, bits: int, width: int) -> Tuple[Tensor, Tensor]:
    """packed uint8 [B,Cin,H,W*bits/8] -> (x uint8 [B,Cin,H,W], r [B,4,Cin]); see sast_unpack_nonzero_ratio."""
    L.require_cuda(data, "data")
    data = data.contiguous()
    B, Cin, H, _ = data.shape
    x = torch.empty(B, Cin, H, width, device=data.device, dtype=torch.uint8)
    r = torch.empty(4, B, Cin, device=data.device, dtype=torch.float32)
    scratch = torch.zeros(B * Cin * 4, device=data.device, dtype=torch.int32)
    L.run(data.device, "sast_unpack_nonzero_ratio", data.data_ptr(), int(bits), B, Cin, H, int(width), x.data_ptr(),
          r.data_ptr(), scratch.data_ptr())
    return x, r.permute(1, 0, 2)


@unpack_nonzero_ratio.register_fake
def _(data, bits, width):
    B, Cin, H, _ = data.shape
    return (data.new_empty(B, Cin, H, width), data.new_empty(4, B, Cin, dtype=torch.float32).permute(1, 0, 2))


# --------------------------------------------------------------------------------------------
# a4  scoring + STP weighting
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("sast::score_fwd", mutates_args=())
def score_fwd(x: Tensor, pos: Tensor, r: Tensor, ctrl_w: Tensor, score_w: Tensor, score_b: Tensor,
              amp: float, score_w_hi: Optional[Tensor] = None, score_w_lo: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """score_w_hi / score_w_lo (see :func:`split_tf32`) select the 3xTF32 tensor-core kernel; without them the
    fp32 CUDA-core kernel runs."""
    x = _f32c(x, "x")
    B, H, W, Cc = x.shape
    pos = _f32c(pos, "pos")
    if pos.dim() == 4 and pos.shape[0] == 1:
        pos = pos[0]
    pstride = 0 if pos.dim() == 3 else H * W * Cc
    assert pos.shape[-3:] == (H, W, Cc), f"pos shape {tuple(pos.shape)} does not match x {tuple(x.shape)}"
    r = _f32c(r, "r")
    xw = torch.empty_like(x)
    tok = torch.empty(B, H, W, device=x.device, dtype=torch.float32)
    scratch = torch.empty(2 * B * Cc + (Cc // 32) * B * H * W, device=x.device, dtype=torch.float32)
    # converted copies (non-fp32 / non-contiguous parameters) must stay alive until after the launch
    ctrl_w, score_w, score_b = _f32c(ctrl_w, "ctrl_w"), _f32c(score_w, "score_w"), _f32c(score_b, "score_b")
    a = L.ScoreArgs(_geom(B, H, W, Cc, 1, 1), x.data_ptr(), pos.data_ptr(), pstride, r.data_ptr(), r.shape[1],
                    ctrl_w.data_ptr(), score_w.data_ptr(), score_b.data_ptr(), float(amp), xw.data_ptr(), tok.data_ptr(),
                    scratch.data_ptr(), L.ptr(score_w_hi), L.ptr(score_w_lo))
    L.run(x.device, "sast_score_fwd", C.byref(a))
    return xw, tok


def split_tf32(w: Tensor) -> Tuple[Tensor, Tensor]:
    """w = hi + lo (+ ~2^-22 |w|) with both halves representable in TF32 (round to nearest, ties away,
    like cvt.rna.tf32.f32): the weight operands of the 3xTF32 scoring GEMM."""
    def rna(t):
        bits = t.contiguous().view(torch.int32)
        return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    w = w.detach().float().contiguous()
    hi = rna(w)
    return hi, rna(w - hi)


@score_fwd.register_fake
def _(x, pos, r, ctrl_w, score_w, score_b, amp, score_w_hi=None, score_w_lo=None):
    return torch.empty_like(x, dtype=torch.float32), x.new_empty(x.shape[:3], dtype=torch.float32)


@torch.library.custom_op("sast::score_bwd", mutates_args=())
def score_bwd(d_xw: Tensor, x: Tensor, pos: Tensor, r: Tensor, ctrl_w: Tensor, score_w: Tensor, score_b: Tensor
              ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Gradient of the STP weighting xw = sig(ctrl) sig(relu(x0 Ws^T + bs)) x0 (SAST.py:105-114): (dx, dWs, dbs, dWc).
    Hand-written kernels (sast_score_bwd), fp32, recompute based; tok_score carries no gradient (selection)."""
    x, d_xw = _f32c(x, "x"), _f32c(d_xw, "d_xw")
    B, H, W, Cc = x.shape
    pos = _f32c(pos, "pos")
    if pos.dim() == 4 and pos.shape[0] == 1:
        pos = pos[0]
    pstride = 0 if pos.dim() == 3 else H * W * Cc
    r, ctrl_w, score_w, score_b = _f32c(r, "r"), _f32c(ctrl_w, "ctrl_w"), _f32c(score_w, "score_w"), _f32c(score_b, "score_b")
    dx = torch.empty_like(x)
    d_sw, d_sb, d_cw = torch.zeros_like(score_w), torch.zeros_like(score_b), torch.empty_like(ctrl_w)
    lib = L.lib()
    nbytes = lib.sast_score_bwd_workspace_bytes(B * H * W, Cc, B)
    ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    a = L.ScoreArgs(_geom(B, H, W, Cc, 1, 1), x.data_ptr(), pos.data_ptr(), pstride, r.data_ptr(), r.shape[1],
                    ctrl_w.data_ptr(), score_w.data_ptr(), score_b.data_ptr(), 0.0, 0, 0, 0, 0, 0)
    with torch.cuda.device(x.device):
        rc = lib.sast_score_bwd(C.byref(a), d_xw.data_ptr(), dx.data_ptr(), d_sw.data_ptr(), d_sb.data_ptr(), d_cw.data_ptr(),
                                ws.data_ptr(), nbytes, L.stream_ptr(x.device))
    L.check(rc, "sast_score_bwd")
    return dx, d_sw, d_sb, d_cw


@score_bwd.register_fake
def _(d_xw, x, pos, r, ctrl_w, score_w, score_b):
    return torch.empty_like(x), torch.empty_like(score_w), torch.empty_like(score_b), torch.empty_like(ctrl_w)


def _score_setup(ctx, inputs, output):
    x, pos, r, ctrl_w, score_w, score_b = inputs[:6]
    ctx.save_for_backward(x, pos, r, ctrl_w, score_w, score_b)
    ctx.set_materialize_grads(True)


def _score_backward(ctx, d_xw, d_tok):
    x, pos, r, ctrl_w, score_w, score_b = ctx.saved_tensors
    dx, d_sw, d_sb, d_cw = score_bwd(d_xw, x, pos, r, ctrl_w, score_w, score_b)
    return dx, None, None, d_cw, d_sw, d_sb, None, None, None


torch.library.register_autograd("sast::score_fwd", _score_backward, setup_context=_score_setup)


@torch.library.custom_op("sast::add_pos", mutates_args=())
def add_pos(x: Tensor, pos: Tensor) -> Tensor:
    x = _f32c(x, "x")
    B, H, W, Cc = x.shape
    pos = _f32c(pos, "pos")
    if pos.dim() == 4 and pos.shape[0] == 1:
        pos = pos[0]
    pstride = 0 if pos.dim() == 3 else H * W * Cc
    out = torch.empty_like(x)
    a = L.ScoreArgs(_geom(B, H, W, Cc, 1, 1), x.data_ptr(), pos.data_ptr(), pstride, 0, 0, 0, 0, 0, 0.0,
                    out.data_ptr(), 0, 0, 0, 0)
    L.run(x.device, "sast_score_fwd", C.byref(a))
    return out


@add_pos.register_fake
def _(x, pos):
    return torch.empty_like(x, dtype=torch.float32)


torch.library.register_autograd("sast::add_pos", lambda ctx, g: (g, None))


# --------------------------------------------------------------------------------------------
# a5/a6  selection
# --------------------------------------------------------------------------------------------
def _pool_words(B, NW, P) -> int:
    return (L.lib().sast_selection_bytes(B, NW, P) + 3) // 4


def _pool_alloc(B, NW, P, device) -> Tensor:
    return torch.empty(_pool_words(B, NW, P), device=device, dtype=torch.int32)


def _bind(pool: Tensor, B, NW, P) -> L.Selection:
    s = L.Selection()
    L.check(L.lib().sast_selection_bind(pool.data_ptr(), B, NW, P, C.byref(s)), "sast_selection_bind")
    return s


def _select_impl(mode: int, B, H, W, p0, p1, flavor, thr_win, thr_tok, device, tok_score=None, win_prob=None,
                 tok_prob=None, win_flag=None, tok_flag=None, want_probs=False):
    T = p0 * p1
    N = (H * W) // T
    NW, P = B * N, B * H * W
    pool = _pool_alloc(B, NW, P, device)
    sel = _bind(pool, B, NW, P)
    wp = torch.empty(B, N, device=device, dtype=torch.float32) if want_probs else None
    tp = torch.zeros(NW, T, device=device, dtype=torch.float32) if want_probs else None
    a = L.SelectArgs(_geom(B, H, W, 32, p0, p1), int(flavor), mode, L.ptr(tok_score), L.ptr(win_prob),
                     L.ptr(tok_prob), L.ptr(win_flag), L.ptr(tok_flag), float(thr_win), float(thr_tok),
                     L.ptr(wp), L.ptr(tp), sel)
    L.run(device, "sast_select", C.byref(a))
    return pool, wp, tp


@torch.library.custom_op("sast::select", mutates_args=())
def select(tok_score: Tensor, p0: int, p1: int, flavor: int, thr_win: float, thr_tok: float) -> Tensor:
    tok_score = _f32c(tok_score, "tok_score")
    B, H, W = tok_score.shape
    return _select_impl(L.SEL_SCORES, B, H, W, p0, p1, flavor, thr_win, thr_tok, tok_score.device,
                        tok_score=tok_score)[0]


@select.register_fake
def _(tok_score, p0, p1, flavor, thr_win, thr_tok):
    B, H, W = tok_score.shape
    return tok_score.new_empty(_pool_words(B, B * (H * W // (p0 * p1)), B * H * W), dtype=torch.int32)


@torch.library.custom_op("sast::select_pair", mutates_args=())
def select_pair(tok_score: Tensor, p0: int, p1: int, thr_win: float, thr_tok: float) -> Tuple[Tensor, Tensor]:
    """Window-flavour and grid-flavour selections of one score map in the same launches
    (what a first SAST block needs, SAST.py:120-123 and :138-147)."""
    tok_score = _f32c(tok_score, "tok_score")
    B, H, W = tok_score.shape
    T = p0 * p1
    NW, P = B * (H * W // T), B * H * W
    dev = tok_score.device
    pool_a, pool_b = _pool_alloc(B, NW, P, dev), _pool_alloc(B, NW, P, dev)
    sel_a, sel_b = _bind(pool_a, B, NW, P), _bind(pool_b, B, NW, P)
    a = L.SelectArgs(_geom(B, H, W, 32, p0, p1), L.WINDOW, L.SEL_SCORES, tok_score.data_ptr(), 0, 0, 0, 0,
                     float(thr_win), float(thr_tok), 0, 0, sel_a)
    L.run(dev, "sast_select2", C.byref(a), L.GRID, C.byref(sel_b))
    return pool_a, pool_b


@select_pair.register_fake
def _(tok_score, p0, p1, thr_win, thr_tok):
    B, H, W = tok_score.shape
    n = _pool_words(B, B * (H * W // (p0 * p1)), B * H * W)
    return tok_score.new_empty(n, dtype=torch.int32), tok_score.new_empty(n, dtype=torch.int32)


def select_with_probs(tok_score: Tensor, p0, p1, flavor, thr_win, thr_tok):
    """Like :func:`select` but also returns the softmax probabilities the kernel thresholded."""
    tok_score = _f32c(tok_score, "tok_score")
    B, H, W = tok_score.shape
    return _select_impl(L.SEL_SCORES, B, H, W, p0, p1, flavor, thr_win, thr_tok, tok_score.device,
                        tok_score=tok_score, want_probs=True)


def select_from_probs(win_prob: Tensor, tok_prob: Tensor, H: int, W: int, p0: int, p1: int, thr_win: float,
                      thr_tok: float) -> Tensor:
    """Tier-A twin of get_score_index_2d21d / get_score_index_with_padding: thresholds on given
    fp32 probabilities.  win_prob [B,N]; tok_prob [B*N,T] (rows of dropped windows are ignored)."""
    win_prob, tok_prob = _f32c(win_prob, "win_prob"), _f32c(tok_prob, "tok_prob")
    B = win_prob.shape[0]
    return _select_impl(L.SEL_PROBS, B, H, W, p0, p1, L.FLAT, thr_win, thr_tok, win_prob.device,
                        win_prob=win_prob, tok_prob=tok_prob)[0]


def select_from_flags(win_flag: Tensor, tok_flag: Tensor, B: int, H: int, W: int, p0: int, p1: int,
                      flavor: int = L.FLAT) -> Tensor:
    """Selection from explicit keep flags (partitioned order).  `flavor` fixes how compacted rows map
    back to NHWC pixels (row_pix): WINDOW / GRID for a map, FLAT for an already partitioned tensor."""
    L.require_cuda(win_flag, "win_flag")
    win_flag = win_flag.to(torch.uint8).contiguous()
    tok_flag = tok_flag.to(torch.uint8).contiguous()
    return _select_impl(L.SEL_FLAGS, B, H, W, p0, p1, flavor, 0.0, 0.0, win_flag.device, win_flag=win_flag,
                        tok_flag=tok_flag)[0]


class Selection:
    """One layer's device-resident selection (replaces the reference's index list
    [index_window, index_token, padding_index, asy_index, K], SAST.py:123).

    Behaves like that 5-element list when indexed / iterated -- the int64 tensors are built on
    first use (this synchronises with the device; the fast path never does it)."""

    def __init__(self, pool: Tensor, B: int, H: int, W: int, p0: int, p1: int,
                 tok_prob: Optional[Tensor] = None, given: Optional[Sequence[Tensor]] = None):
        self.pool, self.B, self.H, self.W, self.p0, self.p1 = pool, B, H, W, p0, p1
        self.T = p0 * p1
        self.N = (H * W) // self.T
        self.NW, self.P = B * self.N, B * H * W
        self.struct = _bind(pool, B, self.NW, self.P)
        self.tok_prob = tok_prob
        self._lists = list(given) if given is not None else None

    # -- device views (no sync) -------------------------------------------------------------
    def _view(self, field: str, n: int) -> Tensor:
        off = (getattr(self.struct, field) - self.pool.data_ptr()) // 4
        return self.pool[off:off + n]

    @property
    def counts(self) -> Tensor:
        return self._view("counts", 8)

    @property
    def win_K(self) -> Tensor:
        return self._view("win_K", self.NW)

    @property
    def tok_row(self) -> Tensor:
        return self._view("tok_row", self.P)

    @property
    def row_tok(self) -> Tensor:
        return self._view("row_tok", self.P)

    @property
    def sel_win(self) -> Tensor:
        return self._view("sel_win", self.NW)

    # -- reference-style lists (syncs) ----------------------------------------------------------
    def lists(self) -> List[Tensor]:
        if self._lists is None:
            M, S, Kmax = self.counts[:3].tolist()
            iw = self.sel_win[:M].long()
            K = self.win_K[iw].long()
            q = self.row_tok[:S].long()                       # token ids w*T+t, ascending
            rank = self._view("win_rank", self.NW).long()
            asy = rank[q // self.T] * self.T + q % self.T   # compacted [M*T] space
            # index_token = selected U padding per window, Kmax entries, padding = highest-probability
            # unselected tokens (what torch.topk(sorted=False) returns as a set); order is arbitrary.
            keep = torch.zeros(M * self.T, dtype=torch.bool, device=self.pool.device)
            keep[asy] = True
            if self.tok_prob is not None:
                pr = self.tok_prob.view(self.NW, self.T)[iw]
            else:
                pr = torch.zeros(M, self.T, device=self.pool.device)
            key = pr + keep.view(M, self.T).float() * 2.0     # selected first, then by probability
            top = torch.topk(key, k=max(Kmax, 0), dim=1, sorted=False)[1] if M > 0 else key.new_zeros(0, 0).long()
            it = (top + torch.arange(M, device=top.device).view(-1, 1) * self.T).reshape(-1)
            pad = it[~keep[it]]
            self._lists = [iw, it, pad, asy, K]
        return self._lists

    def __iter__(self):
        return iter(self.lists())

    def __getitem__(self, i):
        return self.lists()[i]

    def __len__(self):
        return 5

    def num_selected(self) -> Tensor:
        """S as a 0-dim device tensor (no sync)."""
        return self.counts[1]


def selection_from_lists(index_window: Tensor, asy_index: Tensor, B: int, H: int, W: int, p0: int, p1: int,
                         given: Optional[Sequence[Tensor]] = None, flavor: int = L.FLAT) -> Selection:
    """Build a device selection from reference-style index tensors (MS_WSA.forward called with
    explicit indices, SAST.py:199-201)."""
    T = p0 * p1
    NW = B * (H * W // T)
    dev = index_window.device
    wf = torch.zeros(NW, dtype=torch.uint8, device=dev)
    wf[index_window.long()] = 1
    tf = torch.zeros(NW * T, dtype=torch.uint8, device=dev)
    if asy_index.numel():
        a = asy_index.long()
        tf[index_window.long()[a // T] * T + a % T] = 1
    pool = select_from_flags(wf, tf, B, H, W, p0, p1, flavor)
    return Selection(pool, B, H, W, p0, p1, given=given)


# --------------------------------------------------------------------------------------------
# a8-a13  one MS-WSA layer
# --------------------------------------------------------------------------------------------
WEIGHT_ORDER = ("ln1_w", "ln1_b", "ln2_w", "ln2_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "gamma1", "gamma2",
                "mlp1_w", "mlp1_b", "mlp2_w", "mlp2_b", "qkv_w_bf16", "proj_w_bf16", "mlp1_w_bf16", "mlp2_w_bf16")


def _layer_workspace(lib, P, Cc, mlp_inner, B, precision, enable_cb, device):
    """Scratch of the multi-kernel chain; the fused one-kernel layer (sast_layer_is_fused) needs none."""
    if lib.sast_layer_is_fused(int(P), int(Cc), int(mlp_inner), int(precision), int(bool(enable_cb))):
        return None, 0
    nbytes = lib.sast_layer_workspace_bytes(P, Cc, mlp_inner, B, precision)
    return torch.empty(nbytes, device=device, dtype=torch.uint8), nbytes


@torch.library.custom_op("sast::layer_fwd", mutates_args=())
def layer_fwd(x: Tensor, pool: Tensor, weights: List[Tensor], p0: int, p1: int, flavor: int, precision: int,
              enable_cb: bool, mlp_inner: int, ln_eps: float, dim_head: int) -> Tensor:
    """x [B,H,W,C] fp32 NHWC -> same shape (flavor WINDOW or GRID)."""
    x = _f32c(x, "x")
    L.require_cuda(pool, "pool")
    B, H, W, Cc = x.shape
    g = _geom(B, H, W, Cc, p0, p1)
    out = torch.empty_like(x)
    P = x.numel() // Cc
    lib = L.lib()
    sel = _bind(pool, g.B, (g.H * g.W // (g.p0 * g.p1)) * g.B, P)
    ws, nbytes = _layer_workspace(lib, P, Cc, mlp_inner, g.B, precision, enable_cb, x.device)
    w = L.LayerWeights()
    assert len(weights) == len(WEIGHT_ORDER)
    for name, t in zip(WEIGHT_ORDER, weights):
        setattr(w, name, t.data_ptr() if t.numel() else 0)
    w.I, w.ln_eps, w.dim_head = int(mlp_inner), float(ln_eps), int(dim_head)
    a = L.LayerArgs(g, int(flavor), int(precision), int(bool(enable_cb)), x.data_ptr(), out.data_ptr(), w, sel,
                    L.ptr(ws), nbytes)
    L.run(x.device, "sast_layer_fwd", C.byref(a))
    return out


@layer_fwd.register_fake
def _(x, pool, weights, p0, p1, flavor, precision, enable_cb, mlp_inner, ln_eps, dim_head):
    return torch.empty_like(x, dtype=torch.float32)


GRAD_ORDER = WEIGHT_ORDER[:14]        # the fp32 entries; the bf16 copies carry no gradient


@torch.library.custom_op("sast::layer_bwd", mutates_args=())
def layer_bwd(d_out: Tensor, x: Tensor, pool: Tensor, weights: List[Tensor], p0: int, p1: int, flavor: int,
              mlp_inner: int, ln_eps: float, dim_head: int) -> List[Tensor]:
    """Gradient of one MS-WSA layer: [dx] + one gradient per fp32 entry of `weights` (WEIGHT_ORDER[:14]; an empty
    tensor where the weight is absent).  Hand-written kernels (sast_layer_bwd): fp32 recompute of the layer on the
    compacted rows, then the chain rule backwards; the selection is a constant (as in the reference)."""
    x, d_out = _f32c(x, "x"), _f32c(d_out, "d_out")
    B, H, W, Cc = x.shape
    g = _geom(B, H, W, Cc, p0, p1)
    P = x.numel() // Cc
    lib = L.lib()
    sel = _bind(pool, g.B, (g.H * g.W // (g.p0 * g.p1)) * g.B, P)
    w = L.LayerWeights()
    for name, t in zip(WEIGHT_ORDER, weights):
        setattr(w, name, t.data_ptr() if t.numel() else 0)
    w.I, w.ln_eps, w.dim_head = int(mlp_inner), float(ln_eps), int(dim_head)
    grads = [torch.zeros_like(t, dtype=torch.float32) for t in weights[:14]]
    gs = L.LayerGrads()
    for name, t in zip(GRAD_ORDER, grads):
        setattr(gs, name, t.data_ptr() if t.numel() else 0)
    dx = torch.empty_like(x)
    nbytes = lib.sast_layer_bwd_workspace_bytes(P, Cc, mlp_inner)
    ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
    a = L.LayerArgs(g, int(flavor), L.FP32, 0, x.data_ptr(), 0, w, sel, ws.data_ptr(), nbytes)
    with torch.cuda.device(x.device):
        rc = lib.sast_layer_bwd(C.byref(a), d_out.data_ptr(), dx.data_ptr(), C.byref(gs), L.stream_ptr(x.device))
    L.check(rc, "sast_layer_bwd")
    return [dx] + grads


@layer_bwd.register_fake
def _(d_out, x, pool, weights, p0, p1, flavor, mlp_inner, ln_eps, dim_head):
    return [torch.empty_like(x)] + [torch.empty_like(t, dtype=torch.float32) for t in weights[:14]]


def _layer_setup(ctx, inputs, output):
    x, pool, weights, p0, p1, flavor, precision, enable_cb, mlp_inner, ln_eps, dim_head = inputs
    if enable_cb:
        raise NotImplementedError("sast::layer_fwd backward: context broadcast (enable_CB) is not differentiable here")
    ctx.save_for_backward(x, pool, *weights)
    ctx.meta = (p0, p1, flavor, mlp_inner, ln_eps, dim_head)


def _layer_backward(ctx, d_out):
    x, pool, *weights = ctx.saved_tensors
    p0, p1, flavor, mlp_inner, ln_eps, dim_head = ctx.meta
    out = layer_bwd(d_out, x, pool, list(weights), p0, p1, flavor, mlp_inner, ln_eps, dim_head)
    wg = [gq if weights[i].numel() else None for i, gq in enumerate(out[1:])] + [None] * (len(weights) - 14)
    return out[0], None, wg, None, None, None, None, None, None, None, None


torch.library.register_autograd("sast::layer_fwd", _layer_backward, setup_context=_layer_setup)


def layer_fwd_flat(x: Tensor, sel: Selection, weights, precision, enable_cb, mlp_inner, ln_eps, B: int, dim_head: int = 32) -> Tensor:
    """MS_WSA on an already partitioned [B*N,T,C] tensor (frames matter only for context broadcast)."""
    x = _f32c(x, "x")
    NWn, T, Cc = x.shape
    out = torch.empty_like(x)
    lib = L.lib()
    N = NWn // B
    g = _geom(B, N, T, Cc, 1, T)   # FLAT: a "frame" is N rows of T tokens, window n = row n
    ws, nbytes = _layer_workspace(lib, NWn * T, Cc, mlp_inner, B, precision, enable_cb, x.device)
    w = L.LayerWeights()
    for name, t in zip(WEIGHT_ORDER, weights):
        setattr(w, name, t.data_ptr() if t.numel() else 0)
    w.I, w.ln_eps, w.dim_head = int(mlp_inner), float(ln_eps), int(dim_head)
    a = L.LayerArgs(g, L.FLAT, int(precision), int(bool(enable_cb)), x.data_ptr(), out.data_ptr(), w,
                    sel.struct, L.ptr(ws), nbytes)
    L.run(x.device, "sast_layer_fwd", C.byref(a))
    return out


# --------------------------------------------------------------------------------------------
# standalone gather / scatter and the tensor-core GEMM (tests, microbenchmarks)
# --------------------------------------------------------------------------------------------
def gather_rows(x: Tensor, sel: Selection, flavor: int) -> Tensor:
    x = _f32c(x, "x")
    Cc = x.shape[-1]
    rows = torch.zeros(sel.P, Cc, device=x.device, dtype=torch.float32)
    g = _geom(sel.B, sel.H, sel.W, Cc, sel.p0, sel.p1)
    L.run(x.device, "sast_gather", C.byref(g), flavor, x.data_ptr(), C.byref(sel.struct), rows.data_ptr())
    return rows


def scatter_rows(rows: Tensor, sel: Selection, flavor: int, x: Tensor) -> Tensor:
    """In place: x[token of row r] = rows[r] for the S selected rows."""
    rows = _f32c(rows, "rows")
    Cc = x.shape[-1]
    g = _geom(sel.B, sel.H, sel.W, Cc, sel.p0, sel.p1)
    L.run(x.device, "sast_scatter", C.byref(g), flavor, rows.data_ptr(), C.byref(sel.struct), x.data_ptr())
    return x


def gemm_bf16(A: Tensor, Wt: Tensor, bias: Optional[Tensor] = None, out_bf16: bool = False) -> Tensor:
    """D = A @ Wt.T (+ bias) on tcgen05; A [M,K] bf16, Wt [N,K] bf16."""
    L.require_cuda(A, "A")
    A, Wt = A.contiguous(), Wt.contiguous()
    M, K = A.shape
    N = Wt.shape[0]
    D = torch.empty(M, N, device=A.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    L.run(A.device, "sast_gemm_bf16", A.data_ptr(), Wt.data_ptr(), L.ptr(bias), D.data_ptr(), int(out_bf16), M, N, K)
    return D


# --------------------------------------------------------------------------------------------
# glue of the dense callers around the block (stem input, downsample padding, LayerNorm, LSTM gates)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("sast::pad_input", mutates_args=())
def pad_input(x: Tensor, pad: int) -> Tensor:
    """[B,Cin,H,W] NCHW (uint8 / int32 / float32) -> fp32 NHWC [B,H+2p,W+2p,Cin], replicate padding."""
    L.require_cuda(x, "x")
    if x.dtype == torch.uint8:
        dt = L.U8
    elif x.dtype == torch.int32:
        dt = L.I32
    elif x.dtype == torch.float32:
        dt = L.F32
    else:
        x = x.to(torch.float32 if x.is_floating_point() else torch.int32)
        dt = L.F32 if x.is_floating_point() else L.I32
    x = x.contiguous()
    B, Cin, H, W = x.shape
    out = torch.empty(B, H + 2 * pad, W + 2 * pad, Cin, device=x.device, dtype=torch.float32)
    L.run(x.device, "sast_pad_input", x.data_ptr(), dt, B, Cin, H, W, pad, out.data_ptr())
    return out


@pad_input.register_fake
def _(x, pad):
    B, Cin, H, W = x.shape
    return x.new_empty(B, H + 2 * pad, W + 2 * pad, Cin, dtype=torch.float32)


@torch.library.custom_op("sast::pad_nhwc", mutates_args=())
def pad_nhwc(x: Tensor, pad: int) -> Tensor:
    """fp32 [B,H,W,C] (any strides with contiguous channels) -> dense replicate-padded [B,H+2p,W+2p,C]."""
    L.require_cuda(x, "x")
    if x.dtype != torch.float32:
        x = x.float()
    if x.stride(3) != 1 or any(s % 4 for s in x.stride()[:3]):
        x = x.contiguous()
    B, H, W, Cc = x.shape
    out = torch.empty(B, H + 2 * pad, W + 2 * pad, Cc, device=x.device, dtype=torch.float32)
    L.run(x.device, "sast_pad_nhwc", x.data_ptr(), B, H, W, Cc, pad, x.stride(0), x.stride(1), x.stride(2),
                                  out.data_ptr())
    return out


@pad_nhwc.register_fake
def _(x, pad):
    B, H, W, Cc = x.shape
    return x.new_empty(B, H + 2 * pad, W + 2 * pad, Cc, dtype=torch.float32)


@torch.library.custom_op("sast::layernorm", mutates_args=())
def layernorm(x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], eps: float) -> Tensor:
    x = _f32c(x, "x")
    Cc = x.shape[-1]
    out = torch.empty_like(x)
    weight = None if weight is None else _f32c(weight, "weight")     # keep converted copies alive past the launch
    bias = None if bias is None else _f32c(bias, "bias")
    L.run(x.device, "sast_layernorm", x.data_ptr(), L.ptr(weight), L.ptr(bias), float(eps),
                                   x.numel() // Cc, Cc, out.data_ptr())
    return out


@layernorm.register_fake
def _(x, weight, bias, eps):
    return torch.empty_like(x, dtype=torch.float32)


@torch.library.custom_op("sast::lstm_gates", mutates_args=())
def lstm_gates(mix: Tensor, bias: Optional[Tensor], c_prev: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """mix [..., 4C] channels-last (+ bias [4C]), c_prev [..., C] or None -> (h, c) [..., C]."""
    mix = _f32c(mix, "mix")
    Cc = mix.shape[-1] // 4
    shape = mix.shape[:-1] + (Cc,)
    if c_prev is not None:
        c_prev = _f32c(c_prev, "c_prev")
    h = torch.empty(shape, device=mix.device, dtype=torch.float32)
    c = torch.empty(shape, device=mix.device, dtype=torch.float32)
    bias = None if bias is None else _f32c(bias, "bias")             # keep the converted copy alive past the launch
    L.run(mix.device, "sast_lstm_gates", mix.data_ptr(), L.ptr(bias), L.ptr(c_prev),
                                    mix.numel() // (4 * Cc), Cc, h.data_ptr(),
                                    c.data_ptr())
    return h, c


@lstm_gates.register_fake
def _(mix, bias, c_prev):
    shape = mix.shape[:-1] + (mix.shape[-1] // 4,)
    return mix.new_empty(shape, dtype=torch.float32), mix.new_empty(shape, dtype=torch.float32)


@torch.library.custom_op("sast::lstm_fwd", mutates_args=())
def lstm_fwd(x: Tensor, h_prev: Optional[Tensor], c_prev: Optional[Tensor], w_packed: Tensor,
             bias_packed: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """Fused conv-LSTM cell on NHWC fp32 maps [..., C]; see sast_lstm_fwd in include/sast_b200.h."""
    x = _f32c(x, "x")
    Cc = x.shape[-1]
    if h_prev is not None:
        h_prev, c_prev = _f32c(h_prev, "h_prev"), _f32c(c_prev, "c_prev")
    h, c = torch.empty_like(x), torch.empty_like(x)
    L.run(x.device, "sast_lstm_fwd", x.data_ptr(), L.ptr(h_prev), L.ptr(c_prev), w_packed.data_ptr(), L.ptr(bias_packed),
                                  x.numel() // Cc, Cc, h.data_ptr(), c.data_ptr())
    return h, c


@lstm_fwd.register_fake
def _(x, h_prev, c_prev, w_packed, bias_packed):
    return torch.empty_like(x, dtype=torch.float32), torch.empty_like(x, dtype=torch.float32)


def gemm_bf16_glu(A: Tensor, Wt_interleaved: Tensor, bias_interleaved: Optional[Tensor] = None) -> Tensor:
    """bf16 [M, N/2] = value * gelu_erf(gate) of A @ Wt.T + bias; Wt rows interleaved value_j, gate_j."""
    L.require_cuda(A, "A")
    A, Wt = A.contiguous(), Wt_interleaved.contiguous()
    M, K = A.shape
    N = Wt.shape[0]
    D = torch.empty(M, N // 2, device=A.device, dtype=torch.bfloat16)
    L.run(A.device, "sast_gemm_bf16_glu", A.data_ptr(), Wt.data_ptr(), L.ptr(bias_interleaved), D.data_ptr(), M, N, K)
    return D


def pack_stem_weight(w: Tensor) -> Tuple[Tensor, Tensor, int]:
    """conv.weight [Cout,Cin,7,7] fp32 -> (w_hi, w_lo) fp16 [Cout, G*8] with K ordered (c, ky, kx padded to 8) and the
    Cin*7 (c,ky) groups padded to a multiple of 8; w_hi + w_lo ~= w to ~2^-22."""
    Cout, Cin, kh, kw = w.shape
    assert (kh, kw) == (7, 7)
    groups = Cin * 7
    gpad = (groups + 7) // 8 * 8
    wp = torch.zeros(Cout, gpad, 8, device=w.device, dtype=torch.float32)
    wp[:, :groups, :7] = w.detach().float().reshape(Cout, groups, 7)
    wp = wp.reshape(Cout, gpad * 8)
    hi = wp.to(torch.float16)
    lo = (wp - hi.float()).to(torch.float16)
    return hi.contiguous(), lo.contiguous(), gpad


def pack_stem_weight_nhwc(w: Tensor) -> Tensor:
    """conv.weight [Cout,Cin,7,7] fp32 -> fp16 [7*Cout, 144] for sast_stem_nhwc_fwd: row ky*Cout + n holds w[n, :, ky, :]
    ordered (kx, c), zero-padded from 7*Cin = 140 to 144."""
    Cout, Cin, kh, kw = w.shape
    assert (kh, kw) == (7, 7) and 7 * Cin <= 144
    wp = torch.zeros(7, Cout, 144, device=w.device, dtype=torch.float32)
    wp[:, :, :7 * Cin] = w.detach().float().permute(2, 0, 3, 1).reshape(7, Cout, 7 * Cin)      # [ky, n, kx, c]
    return wp.reshape(7 * Cout, 144).to(torch.float16).contiguous()


def stem_nhwc_supported(Cin: int, H: int, W: int, Cout: int) -> bool:
    return bool(L.lib().sast_stem_nhwc_supported(int(Cin), int(H), int(W), int(Cout)))


def pack_downsample_weight(w: Tensor) -> Tensor:
    """conv.weight [Cout,Cin,3,3] fp32 -> bf16 [Cout, 9*Cin] for sast_downsample_fwd: column ky*3*Cin + kx*Cin + c."""
    Cout, Cin, kh, kw = w.shape
    assert (kh, kw) == (3, 3)
    return w.detach().float().permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(torch.bfloat16).contiguous()


def downsample_supported(Cin: int, H: int, W: int, Cout: int) -> bool:
    return bool(L.lib().sast_downsample_supported(int(Cin), int(H), int(W), int(Cout)))


@torch.library.custom_op("sast::downsample_fwd", mutates_args=())
def downsample_fwd(x: Tensor, w9: Tensor, ln_w: Optional[Tensor], ln_b: Optional[Tensor], eps: float) -> Tensor:
    """fp32 NHWC [B,H,W,Cin] (any strides with contiguous channels) -> LayerNorm(conv3x3/2, replicate padding) fp32 NHWC
    [B,H/2,W/2,Cout]: sast_pad_nhwc_bf16 (padded bf16 operand map) + sast_downsample_fwd."""
    L.require_cuda(x, "x")
    assert x.dtype == torch.float32 and w9.dtype == torch.bfloat16 and w9.is_contiguous()
    if x.stride(3) != 1 or any(s % 4 for s in x.stride()[:3]):
        x = x.contiguous()
    B, H, W, Cin = x.shape
    Cout = w9.shape[0]
    xp = torch.empty(B, H + 2, W + 2, Cin, device=x.device, dtype=torch.bfloat16)
    L.run(x.device, "sast_pad_nhwc_bf16", x.data_ptr(), B, H, W, Cin, 1, x.stride(0), x.stride(1), x.stride(2), xp.data_ptr())
    out = torch.empty(B, H // 2, W // 2, Cout, device=x.device, dtype=torch.float32)
    L.run(x.device, "sast_downsample_fwd", xp.data_ptr(), B, Cin, H, W, w9.data_ptr(), Cout, L.ptr(ln_w), L.ptr(ln_b),
          float(eps), out.data_ptr())
    return out


@downsample_fwd.register_fake
def _(x, w9, ln_w, ln_b, eps):
    B, H, W, _ = x.shape
    return x.new_empty(B, H // 2, W // 2, w9.shape[0])


def pack_stem_weight_bits(w: Tensor) -> Tensor:
    """conv.weight [Cout,Cin,7,7] fp32 -> fp16 [7*Cout, Cin*8] for sast_stem_bits_fwd: row ky*Cout + n holds w[n, c, ky, kx] at
    column c*8 + kx; the 8th tap of every bin is zero."""
    Cout, Cin, kh, kw = w.shape
    assert (kh, kw) == (7, 7)
    wp = torch.zeros(7, Cout, Cin, 8, device=w.device, dtype=torch.float32)
    wp[..., :7] = w.detach().float().permute(2, 0, 1, 3)                                        # [ky, n, c, kx]
    return wp.reshape(7 * Cout, Cin * 8).to(torch.float16).contiguous()


def stem_bits_supported(bits: int, Cin: int, H: int, W: int, Cout: int) -> bool:
    return bool(L.lib().sast_stem_bits_supported(int(bits), int(Cin), int(H), int(W), int(Cout)))


@torch.library.custom_op("sast::packed_nonzero_ratio", mutates_args=())
def packed_nonzero_ratio(data: Tensor, bits: int, width: int) -> Tensor:
    """packed uint8 [B,Cin,H,W*bits/8] -> r [B,4,Cin] without unpacking; see sast_events_nhwc (xh == NULL)."""
    L.require_cuda(data, "data")
    data = data.contiguous()
    B, Cin, H, _ = data.shape
    r = torch.empty(4, B, Cin, device=data.device, dtype=torch.float32)
    scratch = None if bits == 1 else torch.zeros(B * Cin * 4, device=data.device, dtype=torch.int32)   # 1 bit: one CTA per plane, no scratch
    L.run(data.device, "sast_events_nhwc", data.data_ptr(), int(bits), B, Cin, H, int(width), 0, r.data_ptr(), L.ptr(scratch))
    return r.permute(1, 0, 2)


@packed_nonzero_ratio.register_fake
def _(data, bits, width):
    B, Cin = data.shape[:2]
    return data.new_empty(4, B, Cin, dtype=torch.float32).permute(1, 0, 2)


@torch.library.custom_op("sast::stem_bits_fwd", mutates_args=())
def stem_bits_fwd(data: Tensor, bits: int, width: int, w16: Tensor, ln_w: Optional[Tensor], ln_b: Optional[Tensor],
                  eps: float) -> Tensor:
    """1-bit packed uint8 [B,Cin,H,W/8] -> LayerNorm(conv7x7/4) fp32 NHWC [B,H/4,W/4,Cout]; see sast_stem_bits_fwd."""
    L.require_cuda(data, "data")
    assert data.dtype == torch.uint8 and w16.dtype == torch.float16 and w16.is_contiguous()
    data = data.contiguous()
    B, Cin, H, _ = data.shape
    Cout = w16.shape[0] // 7
    out = torch.empty(B, H // 4, width // 4, Cout, device=data.device, dtype=torch.float32)
    L.run(data.device, "sast_stem_bits_fwd", data.data_ptr(), int(bits), B, Cin, H, int(width), w16.data_ptr(), Cout, L.ptr(ln_w),
          L.ptr(ln_b), float(eps), out.data_ptr())
    return out


@stem_bits_fwd.register_fake
def _(data, bits, width, w16, ln_w, ln_b, eps):
    return data.new_empty(data.shape[0], data.shape[2] // 4, width // 4, w16.shape[0] // 7, dtype=torch.float32)


class EventsNHWC:
    """The stem's input after :func:`events_nhwc`: fp16 [B, H+8, W+8, Cin] with the replicate padding materialised."""

    def __init__(self, xh: Tensor, H: int, W: int):
        self.xh, self.H, self.W = xh, H, W


@torch.library.custom_op("sast::events_nhwc", mutates_args=())
def events_nhwc(data: Tensor, bits: int, width: int, want_r: bool) -> Tuple[Tensor, Tensor]:
    """histogram (packed uint8 [B,Cin,H,W*bits/8], bits 1 / 4; or uint8 [B,Cin,H,W], bits 8) -> (xh fp16
    [B,H+8,W+8,Cin], r [B,4,Cin] (empty unless want_r)); see sast_events_nhwc."""
    L.require_cuda(data, "data")
    assert data.dtype == torch.uint8
    data = data.contiguous()
    B, Cin, H, _ = data.shape
    xh = torch.empty(B, H + 8, width + 8, Cin, device=data.device, dtype=torch.float16)
    r = torch.empty(4, B, Cin, device=data.device, dtype=torch.float32)
    scratch = torch.zeros(B * Cin * 4, device=data.device, dtype=torch.int32) if (want_r and bits != 1) else None
    L.run(data.device, "sast_events_nhwc", data.data_ptr(), int(bits), B, Cin, H, int(width), xh.data_ptr(),
          r.data_ptr() if want_r else 0, L.ptr(scratch))
    return xh, r.permute(1, 0, 2)


@events_nhwc.register_fake
def _(data, bits, width, want_r):
    B, Cin, H, _ = data.shape
    return (data.new_empty(B, H + 8, width + 8, Cin, dtype=torch.float16),
            data.new_empty(4, B, Cin, dtype=torch.float32).permute(1, 0, 2))


@torch.library.custom_op("sast::stem_nhwc_fwd", mutates_args=())
def stem_nhwc_fwd(xh: Tensor, H: int, W: int, w16: Tensor, ln_w: Optional[Tensor], ln_b: Optional[Tensor], eps: float) -> Tensor:
    """fp16 [B,H+8,W+8,Cin] -> LayerNorm(conv7x7/4) fp32 NHWC [B,H/4,W/4,Cout]; see sast_stem_nhwc_fwd."""
    L.require_cuda(xh, "xh")
    assert xh.dtype == torch.float16 and xh.is_contiguous() and w16.dtype == torch.float16 and w16.is_contiguous()
    B, Hp, Wp, Cin = xh.shape
    assert (Hp, Wp) == (H + 8, W + 8)
    Cout = w16.shape[0] // 7
    out = torch.empty(B, H // 4, W // 4, Cout, device=xh.device, dtype=torch.float32)
    L.run(xh.device, "sast_stem_nhwc_fwd", xh.data_ptr(), B, Cin, int(H), int(W), w16.data_ptr(), Cout, L.ptr(ln_w), L.ptr(ln_b),
          float(eps), out.data_ptr())
    return out


@stem_nhwc_fwd.register_fake
def _(xh, H, W, w16, ln_w, ln_b, eps):
    return xh.new_empty(xh.shape[0], H // 4, W // 4, w16.shape[0] // 7, dtype=torch.float32)


@torch.library.custom_op("sast::stem_fwd", mutates_args=())
def stem_fwd(x: Tensor, w_hi: Tensor, w_lo: Tensor, n_groups_pad: int, ln_w: Optional[Tensor], ln_b: Optional[Tensor],
             eps: float) -> Tensor:
    """uint8 [B,Cin,H,W] -> LayerNorm(conv7x7/4) fp32 NHWC [B,H/4,W/4,Cout]; see sast_stem_fwd."""
    L.require_cuda(x, "x")
    assert x.dtype == torch.uint8
    x = x.contiguous()
    B, Cin, H, W = x.shape
    Cout = w_hi.shape[0]
    out = torch.empty(B, H // 4, W // 4, Cout, device=x.device, dtype=torch.float32)
    L.run(x.device, "sast_stem_fwd", x.data_ptr(), B, Cin, H, W, w_hi.data_ptr(), w_lo.data_ptr(), Cout, n_groups_pad,
                                  L.ptr(ln_w), L.ptr(ln_b), float(eps), out.data_ptr())
    return out


@stem_fwd.register_fake
def _(x, w_hi, w_lo, n_groups_pad, ln_w, ln_b, eps):
    B, Cin, H, W = x.shape
    return x.new_empty(B, H // 4, W // 4, w_hi.shape[0], dtype=torch.float32)
