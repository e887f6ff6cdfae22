/*
 * sast_b200.h -- C ABI of libsast_b200.so: the SAST scene-adaptive sparse-attention
 * block (SAST, CVPR'24) as hand-written CUDA for NVIDIA B200 (sm_100a).
 *
 * The reference (Peterande/SAST) is pure PyTorch: there is no FFI in it to mirror.  The
 * boundary a reference user sees is the Python module API (sast_b200.SAST_block, MS_WSA,
 * RNNDetector -- same constructor / forward signatures and state-dict keys); this header
 * is the plain-C layer underneath it, bound with ctypes.  Each entry
 * point cites the reference lines it replaces.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns
 *     every buffer; nothing is allocated, freed or synchronised inside;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *   - data-dependent counts (selected windows M, selected tokens S, Kmax) are written to
 *     device memory (sast_selection.counts), never returned to the host;
 *   - return value: 0 ok; <0 invalid argument (SAST_E_*); >0 a cudaError_t;
 *   - re-entrant: no global mutable state; concurrent calls on different streams with
 *     disjoint buffers are safe.
 *
 * Feature maps are NHWC fp32 [B,H,W,C] and are never physically partitioned: a "window"
 * flavour addresses a p0 x p1 patch, a "grid" flavour the (H/p0, W/p1)-strided lattice,
 * exactly the index maps of ops.py:189-220.
 */
#ifndef SAST_B200_H_
#define SAST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAST_ABI_VERSION 2

enum sast_error {
  SAST_OK = 0,
  SAST_E_NULL = -1,      /* a required pointer is NULL            */
  SAST_E_SHAPE = -2,     /* H % p0, W % p1, C % 32 ... violated   */
  SAST_E_UNSUPPORTED = -3,
  SAST_E_WORKSPACE = -4  /* workspace too small                   */
};

enum sast_flavor {       /* how (window w, token t) maps to a pixel of the NHWC map */
  SAST_WINDOW = 0,       /* ops.py:189-195 window_partition                         */
  SAST_GRID = 1,         /* ops.py:206-212 grid_partition                           */
  SAST_FLAT = 2          /* x is already [B*N, T, C] (MS_WSA.forward called directly) */
};

enum sast_precision {
  SAST_FP32 = 0,         /* CUDA-core fp32 FMA everywhere (validation grade)        */
  SAST_BF16 = 1,         /* tcgen05 bf16 operands / fp32 accumulate for GEMMs+attention: the fused one-kernel
                            layer where sast_layer_is_fused() says so, else the multi-kernel chain          */
  SAST_BF16_CHAIN = 2    /* same arithmetic, always the multi-kernel chain (A/B comparisons, tests)          */
};

enum sast_dtype { SAST_U8 = 0, SAST_I32 = 1, SAST_F32 = 2, SAST_I64 = 3, SAST_F16 = 4, SAST_BF16_T = 5 };

/* Geometry of one stage's feature map and its partition. T = p0*p1, N = H*W/T. */
typedef struct sast_geom {
  int32_t B, H, W, C;
  int32_t p0, p1;
} sast_geom;

/*
 * One layer's selection, device resident.  NW = B*N windows, P = B*H*W tokens.
 * "Partitioned order" of a token is q = w*T + t with w = b*N + n (the row order of the
 * reference's [B*N, T, C] tensors).  Replaces the index lists
 * [index_window, index_token, padding_index, asy_index, K] of SAST.py:123.
 */
typedef struct sast_selection {
  int32_t* counts;    /* [8]  0:M  1:S  2:Kmax  3:number of entries of tile_list  4:work tickets of sast_layer_fwd (zero between launches)  5..7 reserved */
  int32_t* win_K;     /* [NW]   selected tokens in window w (0 when the window is dropped)  */
  int32_t* win_rank;  /* [NW]   rank m of window w among selected windows, or -1            */
  int32_t* win_row0;  /* [NW+1] first compacted row of window w (exclusive prefix of win_K) */
  int32_t* sel_win;   /* [NW]   first M entries: ascending ids of selected windows (= index_window) */
  int32_t* tok_row;   /* [P]    compacted row of token q (partitioned order), or -1          */
  int32_t* row_tok;   /* [P]    first S entries: token q of compacted row r                 */
  int32_t* row_pix;   /* [P]    first S entries: NHWC pixel index (b*H*W + y*W + x) of row r */
  float*   win_logit; /* [NW]   scratch: mean token score of window w (SCORES mode)         */
  uint8_t* tok_keep;  /* [P]    scratch: keep flag per token, partitioned order             */
  int32_t* tiles;     /* [NW*2] attention tiles (consecutive windows of one frame, <= 128 rows): the j-th tile
                                of frame b is slot b*N + j: tiles[2 slot] = its first window (-1 = unused slot),
                                tiles[2 slot + 1] = one past its last window                   */
  int32_t* tile_list; /* [NW*2] the same tiles as one dense work list (any order): entry k < counts[3] is
                                {first compacted row, number of rows | split << 8}; a tile holds whole windows; split
                                = K of the first window when the tile is exactly two windows of <= 64 tokens, else 0 */
  int32_t* row_win;   /* [P]    first S entries: (first compacted row of row r's window) << 8 | (K of that window):
                                the keys a row attends to, without a dependent look-up (needs P < 2^23)        */
} sast_selection;

/* Bytes of one int32 pool able to hold a sast_selection for (NW, P, B); see sast_selection_bind. */
size_t sast_selection_bytes(int32_t B, int32_t NW, int32_t P);
/* Carve `pool` (>= sast_selection_bytes, 16-byte aligned) into a sast_selection. */
int sast_selection_bind(void* pool, int32_t B, int32_t NW, int32_t P, sast_selection* out);

/*
 * a1  scene sparsity ratio r.  ref: sast_rnn.py:45-60 non_zero_ratio.
 * x [B,Cin,H,W] of `dtype` (SAST_U8 / SAST_I32 / SAST_F32) -> r [4,B,Cin] fp32 (level-major, so that the
 * per-stage slice r[level] the scoring kernel takes is contiguous), bit-exact:
 * count of non-zero cells after max-pooling by 4,8,16,32 (int16 wrap), times fp32(B/numel).
 */
int sast_nonzero_ratio(const void* x, int32_t dtype, int32_t B, int32_t Cin, int32_t H, int32_t W,
                       float* r, int32_t* scratch /* [B*Cin*4] ints, zero on entry, left zero on exit */, void* stream);

/*
 * Bit-packed input: the event histogram as BITS (1 or 4) bits per bin, packed along x, little endian (bits = 1: bit k of
 * byte j is column 8 j + k; bits = 4: low nibble of byte j = column 2 j) -- [B,Cin,H,W*bits/8] bytes.  Event histograms are
 * binary in benchmark.py:58-60 and clipped at count_cutoff 10 in the datasets (representations.py), so 1 / 4 bits are
 * lossless; the host->device copy shrinks 8x / 2x.  One pass unpacks to uint8 [B,Cin,H,W] (x_out, the stem's input) and
 * produces r exactly as sast_nonzero_ratio does.  W % 8 == 0.
 */
int sast_unpack_nonzero_ratio(const uint8_t* packed, int32_t bits, int32_t B, int32_t Cin, int32_t H, int32_t W,
                              uint8_t* x_out, float* r, int32_t* scratch, void* stream);

/*
 * a4  scoring module + STP weighting.  ref: SAST.py:105-119, 305-328.
 *   x0 = x + pos;  ctrl = exp(Wc) (r + 1e-6);  s = relu(x0 Ws^T + bs)
 *   xw = sigmoid(ctrl) sigmoid(s) x0   (NHWC, same layout as x)
 *   tok_score[p] = sum_c | (amp/ctrl_c) s_c |     (the only thing selection needs)
 * pos: [H,W,C] table when pos_batch_stride == 0, else [B,H,W,C] with that element stride.
 * When Ws == NULL (non-first block, SAST.py:124-128) only xw = x + pos is produced.
 */
typedef struct sast_score_args {
  sast_geom g;
  const float* x;          /* [B,H,W,C] */
  const float* pos;
  int64_t pos_batch_stride;
  const float* r;          /* [B,n_bins] */
  int32_t n_bins;          /* 20 */
  const float* ctrl_w;     /* to_controls.weight [C,n_bins] (pre-exp) */
  const float* score_w;    /* to_scores.weight [C,C] */
  const float* score_b;    /* to_scores.bias [C] */
  float amp;
  float* xw;               /* out [B,H,W,C]; must not alias x */
  float* tok_score;        /* out [B,H,W] */
  float* ctrl_scratch;     /* scratch floats: 2*B*C (sigmoid(ctrl), amp/ctrl) + (C/32)*B*H*W (partial scores) */
  const float* score_w_hi; /* optional: to_scores.weight split for the 3xTF32 tensor-core path: */
  const float* score_w_lo; /*   w_hi = tf32(w), w_lo = tf32(w - w_hi), both [C,C] fp32. NULL -> fp32 FMA kernel */
} sast_score_args;
int sast_score_fwd(const sast_score_args* a, void* stream);

/*
 * a5/a6  scene-adaptive selection.  ref: SAST.py:84-96, 258-281.
 * mode SAST_SEL_SCORES: from tok_score [B,H,W] (map order): window logit = mean of its T
 *   token scores, softmax over the frame's N windows, keep if >= thr_win; per kept window
 *   softmax over its T token scores, keep if >= thr_tok.
 * mode SAST_SEL_PROBS : from post-softmax probabilities (win_prob [B,N], tok_prob [NW,T],
 *   partitioned order) -- bit-exact twin of get_score_index_2d21d /
 *   get_score_index_with_padding on identical fp32 inputs.
 * mode SAST_SEL_FLAGS : from explicit keep flags (win_flag [NW], tok_flag [NW*T] uint8).
 * thr_* are the fp32-cast thresholds float((1/N)/(1+BOUNCE)), float((1/T)/(1+BOUNCE)).
 */
enum sast_select_mode { SAST_SEL_SCORES = 0, SAST_SEL_PROBS = 1, SAST_SEL_FLAGS = 2 };
typedef struct sast_select_args {
  sast_geom g;
  int32_t flavor;          /* SAST_WINDOW or SAST_GRID (how tok_score is partitioned) */
  int32_t mode;
  const float* tok_score;  /* SCORES */
  const float* win_prob;   /* PROBS  */
  const float* tok_prob;   /* PROBS  */
  const uint8_t* win_flag; /* FLAGS  */
  const uint8_t* tok_flag; /* FLAGS  */
  float thr_win, thr_tok;
  float* win_prob_out;     /* optional [B,N]: softmax probabilities (SCORES mode)  */
  float* tok_prob_out;     /* optional [NW,T] */
  sast_selection sel;      /* out */
} sast_select_args;
int sast_select(const sast_select_args* a, void* stream);
/* Same, producing a second selection `sel_b` under partition flavour `flavor_b` from the same
 * inputs in the same launches (a first block needs the window and the grid selection of one
 * score map, SAST.py:138-147).  sel_b == NULL behaves like sast_select. */
int sast_select2(const sast_select_args* a, int32_t flavor_b, const sast_selection* sel_b, void* stream);

/*
 * a8-a13  one MS-WSA layer.  ref: SAST.py:199-255 (MS_WSA.forward).
 *   n1 = LN1(x);  unselected tokens: out = n1.
 *   selected tokens (compacted, S rows): n2 = LN2(n1); qkv = n2 Wqkv^T + b (head-major
 *   [h][q,k,v][32]); per window softmax(q k^T / sqrt(32)) v over the window's selected
 *   tokens only; y = n2 + g1 (o Wp^T + bp); out = y + g2 MLP_GLU(y); scattered back.
 *   enable_cb: context broadcast (SAST.py:240-246).
 */
typedef struct sast_layer_weights {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const float *qkv_w, *qkv_b;   /* [3C,C], [3C] (bias may be NULL)                       */
  const float *proj_w, *proj_b; /* [C,C], [C]                                             */
  const float *gamma1, *gamma2; /* LayerScale [C] (NULL = identity)                       */
  const float *mlp1_w, *mlp1_b; /* GLU proj, rows INTERLEAVED value_j, gate_j: [2I,C],[2I] */
  const float *mlp2_w, *mlp2_b; /* [C,I], [C]                                             */
  const uint16_t *qkv_w_bf16, *proj_w_bf16, *mlp1_w_bf16, *mlp2_w_bf16; /* SAST_BF16 only */
  int32_t I;                    /* GLU width */
  float ln_eps;
  int32_t dim_head;             /* 0 or 32: the default (all precisions); 8, 16, 24: SAST_FP32 only (the reference's
                                   "small" configs use 24, config/experiment/gen1/small.yaml:10); C % dim_head == 0 */
} sast_layer_weights;

typedef struct sast_layer_args {
  sast_geom g;
  int32_t flavor;
  int32_t precision;
  int32_t enable_cb;
  const float* x;          /* in  [B,H,W,C] (or [NW,T,C] for SAST_FLAT) */
  float* out;              /* out, same layout; must not alias x        */
  sast_layer_weights w;
  sast_selection sel;
  void* workspace;
  size_t workspace_bytes;  /* >= sast_layer_workspace_bytes(P, C, I, B, precision); 256-byte aligned */
} sast_layer_args;
size_t sast_layer_workspace_bytes(int64_t P, int32_t C, int32_t I, int32_t B, int32_t precision);
int sast_layer_fwd(const sast_layer_args* a, void* stream);
/* 1 if sast_layer_fwd runs this configuration (P = B*H*W tokens) as ONE fused kernel (SAST_BF16, dim_head 32, C = 64,
 * or C = 128 with P >= 8192, with the mlp_ratio-4 GLU width, no context broadcast): the whole layer per 128-row tile
 * with qkv / attention / MLP intermediates kept in shared and tensor memory.  Such calls need no workspace (workspace
 * may be NULL).  0: the multi-kernel chain. */
int32_t sast_layer_is_fused(int64_t P, int32_t C, int32_t I, int32_t precision, int32_t enable_cb);

/*
 * Backward twins for the training configuration (the reference trains through stock autograd: train.py,
 * modules/detection.py:113-221).  Recompute based: nothing is saved by the forward; fp32 throughout.
 * The selection is a constant of the backward (SAST.py's index tensors carry no gradient).
 *
 * sast_layer_bwd: a = the forward's arguments (x, selection, fp32 weights; a.out unused; a.workspace >=
 *   sast_layer_bwd_workspace_bytes); d_out [B,H,W,C] -> dx [B,H,W,C] and the parameter gradients, ACCUMULATED into the
 *   buffers of `grads` (zero them first; same shapes / row order as sast_layer_weights, mlp1 rows interleaved; an
 *   entry may be NULL where the weight is).  enable_cb is not supported (SAST_E_UNSUPPORTED); needs 2 I >= 3 C.
 * sast_score_bwd: a = the forward's arguments (xw / tok_score / scratch unused); d_xw -> dx and, for a first block,
 *   d to_scores.weight [C,C] / .bias [C] (accumulated, zero them first) and d to_controls.weight [C,n_bins] (written).
 *   tok_score is not differentiable (selection), as in the reference.
 */
typedef struct sast_layer_grads {
  float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  float *qkv_w, *qkv_b, *proj_w, *proj_b, *gamma1, *gamma2, *mlp1_w, *mlp1_b, *mlp2_w, *mlp2_b;
} sast_layer_grads;
size_t sast_layer_bwd_workspace_bytes(int64_t P, int32_t C, int32_t I);
int sast_layer_bwd(const sast_layer_args* a, const float* d_out, float* dx, const sast_layer_grads* grads, void* stream);
size_t sast_score_bwd_workspace_bytes(int64_t P, int32_t C, int32_t B);
int sast_score_bwd(const sast_score_args* a, const float* d_xw, float* dx, float* d_score_w, float* d_score_b,
                   float* d_ctrl_w, void* workspace, size_t workspace_bytes, void* stream);

/* Standalone gather / scatter of the selected tokens (a9 / a13), for tests and the
 * HBM-roofline microbenchmark: rows [S,C] <-> map tokens, through sel.row_tok. */
int sast_gather(const sast_geom* g, int32_t flavor, const float* x, const sast_selection* sel,
                float* rows, void* stream);
int sast_scatter(const sast_geom* g, int32_t flavor, const float* rows, const sast_selection* sel,
                 float* x, void* stream);

/*
 * Memory-bound glue of the dense callers on either side of the block.
 * sast_pad_input : [B,Cin,H,W] NCHW of `dtype` (SAST_U8/I32/F32) -> fp32 NHWC [B,H+2p,W+2p,Cin]
 *                  with replicate padding (ref: sast_rnn.py:153 x.float() + ops.py:77-89 conv
 *                  padding_mode='replicate').
 * sast_pad_nhwc  : fp32 NHWC map with element strides (stride_b, stride_y, stride_x; channels
 *                  contiguous) -> dense replicate-padded NHWC.
 * sast_layernorm : LayerNorm over the last dim of [P,C] fp32 (ref: ops.py:85,90); weight/bias may be NULL.
 * sast_lstm_gates: mix [P,4C] (1x1-conv output, channels [f,i,o | g]) + bias [4C] (may be NULL), c_prev [P,C] or NULL (zero state)
 *                  -> h [P,C], c [P,C]   (ref: models/layers/rnn.py:58-69).
 */
int sast_pad_input(const void* x, int32_t dtype, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t pad,
                   float* out, void* stream);
int sast_pad_nhwc(const float* x, int32_t B, int32_t H, int32_t W, int32_t C, int32_t pad, int64_t stride_b,
                  int64_t stride_y, int64_t stride_x, float* out, void* stream);
int sast_layernorm(const float* x, const float* weight, const float* bias, float eps, int64_t P, int32_t C,
                   float* out, void* stream);
int sast_lstm_gates(const float* mix, const float* bias, const float* c_prev, int64_t P, int32_t C, float* h_out,
                    float* c_out, void* stream);

/*
 * Stem as one kernel (ref: sast_rnn.py:153 x.float(), ops.py:54-91 ConvDownsampling_Cf2Cl with patch_size 4):
 * x [B,Cin,H,W] uint8 NCHW -> out [B,H/4,W/4,Cout] fp32 NHWC = LayerNorm(conv 7x7, stride 4, replicate padding 3,
 * no bias), an implicit GEMM on tcgen05.  Weights come pre-packed: fp16 [Cout, n_groups_pad*8], K ordered
 * (c, ky, kx) with kx padded 7 -> 8 and (c,ky) groups padded to a multiple of 8 (zeros), split w = w_hi + w_lo
 * so that the product is fp32-grade (event counts are exact in fp16).  H, W % 4 == 0; Cout % 32 == 0, <= 256.
 */
int sast_stem_fwd(const uint8_t* x, int32_t B, int32_t Cin, int32_t H, int32_t W, const uint16_t* w_hi,
                  const uint16_t* w_lo, int32_t Cout, int32_t n_groups_pad, const float* ln_w, const float* ln_b,
                  float eps, float* out, void* stream);

/*
 * TMA-fed stem for the 16-bit mode (same reference lines as sast_stem_fwd), two calls:
 *
 * sast_events_nhwc: the event histogram src -- bit-packed along x (bits = 1 / 4, the format of
 * sast_unpack_nonzero_ratio) or plain uint8 (bits = 8), NCHW [B,Cin,H,W*bits/8] -- becomes xh = fp16 NHWC
 * [B, H+8, W+8, Cin] with the stem's replicate padding of 3 materialised (row yy = source row clamp(yy-3), column xx =
 * source column clamp(xx-3); the trailing 5 rows / columns are zeros).  With r != NULL it also computes the scene
 * sparsity ratios r [4,B,Cin] exactly as sast_nonzero_ratio does (scratch: B*Cin*4 zeroed int32, left zeroed; not needed --
 * may be NULL -- for bits = 1, where one CTA counts a whole plane and writes its ratios directly).
 * xh == NULL computes only r (the caller feeds sast_stem_bits_fwd).  W % 32 == 0, Cin == 20 (the stem it feeds).
 *
 * sast_stem_nhwc_fwd: xh -> out [B,H/4,W/4,Cout] fp32 NHWC = LayerNorm(conv 7x7, stride 4, no bias).  The im2col operand
 * is read by TMA straight from xh (a 5-D overlapping-stride tensor map), the fp16 weights stay resident in shared
 * memory: w16 is fp16 [7*Cout, 144], row ky*Cout + n = conv.weight[n, :, ky, :] ordered (kx, c) followed by 4 zeros.
 * One fp16 rounding per weight (finer than the TF32 rounding of the cuDNN convolution it replaces); event counts are
 * exact.  Geometry: sast_stem_nhwc_supported (Cin 20, Cout 64, H % 4 == 0, W % 32 == 0) -- else SAST_E_UNSUPPORTED and
 * the caller takes sast_stem_fwd.
 */
int sast_events_nhwc(const uint8_t* src, int32_t bits, int32_t B, int32_t Cin, int32_t H, int32_t W, uint16_t* xh, float* r,
                     int32_t* scratch, void* stream);
int sast_stem_nhwc_supported(int32_t Cin, int32_t H, int32_t W, int32_t Cout);
int sast_stem_nhwc_fwd(const uint16_t* xh, int32_t B, int32_t Cin, int32_t H, int32_t W, const uint16_t* w16, int32_t Cout,
                       const float* ln_w, const float* ln_b, float eps, float* out, void* stream);

/*
 * Stem straight from the 1-bit packed histogram (same reference lines as sast_stem_fwd; the benchmark's binary input,
 * benchmark.py:58-60): packed uint8 [B,Cin,H,W/8] -> out [B,H/4,W/4,Cout] fp32 NHWC = LayerNorm(conv 7x7, stride 4,
 * replicate padding 3, no bias).  Nothing is unpacked to memory: the producer warps stage the bits a tile touches in
 * shared memory and expand 8-bit windows into fp16 operand chunks through a 256-entry table; fp16 weights resident in
 * shared memory: w16 is fp16 [7*Cout, 160], row ky*Cout + n holds conv.weight[n, c, ky, kx] at column c*8 + kx (column
 * c*8 + 7 zero).  One fp16 rounding per weight, event bits exact.  Geometry: sast_stem_bits_supported (bits 1, Cin 20,
 * Cout 64, H % 4 == 0, W % 32 == 0, W >= 64) -- else SAST_E_UNSUPPORTED and the caller takes the NHWC or uint8 stem.
 */
int sast_stem_bits_supported(int32_t bits, int32_t Cin, int32_t H, int32_t W, int32_t Cout);
int sast_stem_bits_fwd(const uint8_t* packed, int32_t bits, int32_t B, int32_t Cin, int32_t H, int32_t W, const uint16_t* w16,
                       int32_t Cout, const float* ln_w, const float* ln_b, float eps, float* out, void* stream);

/*
 * Downsample stem of stages 2-3 as one kernel (ref: ops.py:54-95 ConvDownsampling_Cf2Cl with downsample factor 2: Conv2d k=3,
 * s=2, replicate padding 1, no bias + NCHW->NHWC + LayerNorm), 16-bit mode: xp bf16 [B,H+2,W+2,Cin] (replicate-padded NHWC,
 * sast_pad_nhwc_bf16: the fp32 map of the previous stage, element strides as in sast_pad_nhwc) -> out [B,H/2,W/2,Cout] fp32
 * NHWC.  Implicit GEMM on tcgen05 with bf16 operands (like every GEMM of the 16-bit mode): im2col by an overlapping-stride
 * 5-D tensor map, weights resident (Cin 64) or streamed per k-block (Cin 128), LayerNorm in the epilogue.
 * w9: bf16 [Cout, 9*Cin], column ky*3*Cin + kx*Cin + c = conv.weight[n,c,ky,kx].
 * Geometry: sast_downsample_supported ((Cin,Cout) = (64,128) or (128,256), H and W even) -- else SAST_E_UNSUPPORTED.
 */
int sast_pad_nhwc_bf16(const float* x, int32_t B, int32_t H, int32_t W, int32_t C, int32_t pad, int64_t stride_b,
                       int64_t stride_y, int64_t stride_x, uint16_t* out, void* stream);
int sast_downsample_supported(int32_t Cin, int32_t H, int32_t W, int32_t Cout);
int sast_downsample_fwd(const uint16_t* xp, int32_t B, int32_t Cin, int32_t H, int32_t W, const uint16_t* w9, int32_t Cout,
                        const float* ln_w, const float* ln_b, float eps, float* out, void* stream);

/*
 * Conv-LSTM cell as one kernel (ref: models/layers/rnn.py:36-69 with dws_conv False): the 1x1 conv of
 * [x | h_prev] as a TF32 tcgen05 GEMM straight from the fp32 NHWC maps, gates in the epilogue.
 * x, h_prev, c_prev, h_out, c_out: [P,C] fp32 (NHWC rows); h_prev/c_prev both NULL = zero state.
 * w_packed [4C,K] fp32, K = C (zero state) or 2C, rows interleaved 4*c + {forget,input,output,cell};
 * bias_packed [4C] in the same order (may be NULL).
 */
int sast_lstm_fwd(const float* x, const float* h_prev, const float* c_prev, const float* w_packed,
                  const float* bias_packed, int64_t P, int32_t C, float* h_out, float* c_out, void* stream);

/* D[M,N] = A[M,K] W[N,K]^T (+bias): bf16 in, fp32 accumulate on tcgen05, fp32 or bf16 out.
 * Exposed for unit tests of the tensor-core path. */
int sast_gemm_bf16(const uint16_t* A, const uint16_t* W, const float* bias, void* D, int32_t d_is_bf16,
                   int32_t M, int32_t N, int32_t K, void* stream);

/* Number of kernels this library has launched (or recorded into a CUDA graph) since it was
 * loaded: a monotonically increasing statistics counter, the only process-wide state. */
uint64_t sast_launch_count(void);

/* D[M,N/2] (bf16) = value * gelu_erf(gate) of A W^T + bias, W rows interleaved value_j, gate_j: the layer's
 * GLU GEMM (ops.py:135-137) standalone, for unit tests and the roofline measurement of bench.py. */
int sast_gemm_bf16_glu(const uint16_t* A, const uint16_t* W, const float* bias, uint16_t* D, int32_t M, int32_t N,
                       int32_t K, void* stream);

/* Debug aid (tools/attn_trace.py, tools/gemm_trace.py): while buf is non-null, launches of the tensor-core
 * attention and GEMM kernels write thread-level clock64 stamps of their phase boundaries into buf
 * (attention: [CTA][16], GEMM and scoring: [CTA][128] int64; last slot = SM id).  Pass NULL to switch it off (the default).
 * Only the instrumented build (`make -C sast_b200/csrc trace`) stamps; the regular library ignores the call. */
void sast_debug_trace(long long* buf, int32_t which /* 1 attention, 2 GEMM, 3 scoring, 4 fused layer, 5 group layer */);

/* Library / build info. */
int sast_abi_version(void);
/* sizeof of an ABI struct: 0 sast_geom, 1 sast_selection, 2 sast_score_args, 3 sast_select_args,
 * 4 sast_layer_weights, 5 sast_layer_args, 6 sast_layer_grads (lets a foreign-language binding verify its mirror). */
size_t sast_struct_size(int32_t which);
const char* sast_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* SAST_B200_H_ */
