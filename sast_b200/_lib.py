"""ctypes binding of ``libsast_b200.so`` (C ABI in ``include/sast_b200.h``).

The library is built in-tree by ``sast_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no fallback: if the shared object is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SAST_B200_LIB selects another build of the same ABI (the instrumented `make trace` twin); there is still no fallback
LIB_PATH = os.environ.get("SAST_B200_LIB") or os.path.join(_HERE, "libsast_b200.so")

# enums (mirror include/sast_b200.h)
WINDOW, GRID, FLAT = 0, 1, 2
FP32, BF16, BF16_CHAIN = 0, 1, 2
U8, I32, F32 = 0, 1, 2
SEL_SCORES, SEL_PROBS, SEL_FLAGS = 0, 1, 2

_ERR = {-1: "SAST_E_NULL (a required pointer is NULL)", -2: "SAST_E_SHAPE (shape constraint violated)",
        -3: "SAST_E_UNSUPPORTED", -4: "SAST_E_WORKSPACE (workspace too small)"}


class Geom(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("p0", C.c_int32), ("p1", C.c_int32)]


class Selection(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("win_K", C.c_void_p), ("win_rank", C.c_void_p),
                ("win_row0", C.c_void_p), ("sel_win", C.c_void_p), ("tok_row", C.c_void_p),
                ("row_tok", C.c_void_p), ("row_pix", C.c_void_p), ("win_logit", C.c_void_p),
                ("tok_keep", C.c_void_p), ("tiles", C.c_void_p), ("tile_list", C.c_void_p), ("row_win", C.c_void_p)]


class ScoreArgs(C.Structure):
    _fields_ = [("g", Geom), ("x", C.c_void_p), ("pos", C.c_void_p), ("pos_batch_stride", C.c_int64),
                ("r", C.c_void_p), ("n_bins", C.c_int32), ("ctrl_w", C.c_void_p), ("score_w", C.c_void_p),
                ("score_b", C.c_void_p), ("amp", C.c_float), ("xw", C.c_void_p), ("tok_score", C.c_void_p),
                ("ctrl_scratch", C.c_void_p), ("score_w_hi", C.c_void_p), ("score_w_lo", C.c_void_p)]


class SelectArgs(C.Structure):
    _fields_ = [("g", Geom), ("flavor", C.c_int32), ("mode", C.c_int32), ("tok_score", C.c_void_p),
                ("win_prob", C.c_void_p), ("tok_prob", C.c_void_p), ("win_flag", C.c_void_p),
                ("tok_flag", C.c_void_p), ("thr_win", C.c_float), ("thr_tok", C.c_float),
                ("win_prob_out", C.c_void_p), ("tok_prob_out", C.c_void_p), ("sel", Selection)]


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_w", "ln1_b", "ln2_w", "ln2_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "gamma1", "gamma2",
        "mlp1_w", "mlp1_b", "mlp2_w", "mlp2_b", "qkv_w_bf16", "proj_w_bf16", "mlp1_w_bf16", "mlp2_w_bf16")] + \
        [("I", C.c_int32), ("ln_eps", C.c_float), ("dim_head", C.c_int32)]


class LayerArgs(C.Structure):
    _fields_ = [("g", Geom), ("flavor", C.c_int32), ("precision", C.c_int32), ("enable_cb", C.c_int32),
                ("x", C.c_void_p), ("out", C.c_void_p), ("w", LayerWeights), ("sel", Selection),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class LayerGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_w", "ln1_b", "ln2_w", "ln2_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "gamma1", "gamma2",
        "mlp1_w", "mlp1_b", "mlp2_w", "mlp2_b")]


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C sast_b200/csrc` (or __graft_entry__.build()). "
            "sast_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
    sig = {
        "sast_abi_version": (C.c_int, []),
        "sast_build_info": (C.c_char_p, []),
        "sast_struct_size": (sz, [i32]),
        "sast_launch_count": (C.c_uint64, []),
        "sast_debug_trace": (None, [vp, i32]),
        "sast_selection_bytes": (sz, [i32, i32, i32]),
        "sast_selection_bind": (C.c_int, [vp, i32, i32, i32, C.POINTER(Selection)]),
        "sast_nonzero_ratio": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp, vp]),
        "sast_unpack_nonzero_ratio": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "sast_score_fwd": (C.c_int, [C.POINTER(ScoreArgs), vp]),
        "sast_select": (C.c_int, [C.POINTER(SelectArgs), vp]),
        "sast_select2": (C.c_int, [C.POINTER(SelectArgs), i32, C.POINTER(Selection), vp]),
        "sast_layer_workspace_bytes": (sz, [i64, i32, i32, i32, i32]),
        "sast_layer_fwd": (C.c_int, [C.POINTER(LayerArgs), vp]),
        "sast_layer_is_fused": (i32, [i64, i32, i32, i32, i32]),
        "sast_layer_bwd_workspace_bytes": (sz, [i64, i32, i32]),
        "sast_layer_bwd": (C.c_int, [C.POINTER(LayerArgs), vp, vp, C.POINTER(LayerGrads), vp]),
        "sast_score_bwd_workspace_bytes": (sz, [i64, i32, i32]),
        "sast_score_bwd": (C.c_int, [C.POINTER(ScoreArgs), vp, vp, vp, vp, vp, vp, sz, vp]),
        "sast_gather": (C.c_int, [C.POINTER(Geom), i32, vp, C.POINTER(Selection), vp, vp]),
        "sast_scatter": (C.c_int, [C.POINTER(Geom), i32, vp, C.POINTER(Selection), vp, vp]),
        "sast_gemm_bf16": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
        "sast_gemm_bf16_glu": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, vp]),
        "sast_pad_input": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, vp]),
        "sast_pad_nhwc": (C.c_int, [vp, i32, i32, i32, i32, i32, i64, i64, i64, vp, vp]),
        "sast_layernorm": (C.c_int, [vp, vp, vp, f32, i64, i32, vp, vp]),
        "sast_lstm_gates": (C.c_int, [vp, vp, vp, i64, i32, vp, vp, vp]),
        "sast_stem_fwd": (C.c_int, [vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp, f32, vp, vp]),
        "sast_events_nhwc": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "sast_stem_nhwc_supported": (C.c_int, [i32, i32, i32, i32]),
        "sast_stem_nhwc_fwd": (C.c_int, [vp, i32, i32, i32, i32, vp, i32, vp, vp, f32, vp, vp]),
        "sast_downsample_supported": (C.c_int, [i32, i32, i32, i32]),
        "sast_pad_nhwc_bf16": (C.c_int, [vp, i32, i32, i32, i32, i32, i64, i64, i64, vp, vp]),
        "sast_downsample_fwd": (C.c_int, [vp, i32, i32, i32, i32, vp, i32, vp, vp, f32, vp, vp]),
        "sast_stem_bits_supported": (C.c_int, [i32, i32, i32, i32, i32]),
        "sast_stem_bits_fwd": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, f32, vp, vp]),
        "sast_lstm_fwd": (C.c_int, [vp, vp, vp, vp, vp, i64, i32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.sast_abi_version() != 2:
        raise RuntimeError("libsast_b200.so ABI version mismatch")
    for which, cls in enumerate((Geom, Selection, ScoreArgs, SelectArgs, LayerWeights, LayerArgs, LayerGrads)):
        if lib.sast_struct_size(which) != C.sizeof(cls):
            raise RuntimeError(f"ctypes mirror of {cls.__name__} is {C.sizeof(cls)} bytes, library says "
                               f"{lib.sast_struct_size(which)}")
    return lib


EXPORTS = ("sast_abi_version", "sast_build_info", "sast_struct_size", "sast_launch_count", "sast_selection_bytes", "sast_selection_bind",
           "sast_nonzero_ratio", "sast_unpack_nonzero_ratio", "sast_score_fwd", "sast_select", "sast_select2", "sast_layer_workspace_bytes",
           "sast_layer_fwd", "sast_layer_is_fused", "sast_layer_bwd_workspace_bytes", "sast_layer_bwd",
           "sast_score_bwd_workspace_bytes", "sast_score_bwd", "sast_gather", "sast_scatter", "sast_gemm_bf16", "sast_gemm_bf16_glu", "sast_pad_input", "sast_pad_nhwc",
           "sast_layernorm", "sast_lstm_gates", "sast_lstm_fwd", "sast_stem_fwd", "sast_events_nhwc", "sast_stem_nhwc_supported",
           "sast_stem_nhwc_fwd", "sast_stem_bits_supported", "sast_stem_bits_fwd", "sast_downsample_supported",
           "sast_downsample_fwd", "sast_pad_nhwc_bf16", "sast_debug_trace")

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"{what}: {_ERR.get(rc, rc)}")
    raise RuntimeError(f"{what}: CUDA error {rc}")


def run(device, name: str, *args):
    """Call entry point `name` with `args` + the current stream of `device`, with that device current (the
    library launches on the current CUDA device; per-device kernel attributes are set inside each call)."""
    with torch.cuda.device(device):
        rc = getattr(lib(), name)(*args, torch.cuda.current_stream(device).cuda_stream)
    check(rc, name)


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"sast_b200: `{name}` must be a CUDA tensor (got {t.device}); there is no CPU path")
