// Internal declarations shared by the MS-WSA layer pipeline (layer.cu) and its
// tensor-core kernels (gemm_tc.cu, attn_tc.cu).
#pragma once
#include "common.cuh"

namespace sast {

// Epilogue flavours of the token-wise GEMMs D = A W^T (+ bias):
enum Epi {
  EPI_STORE = 0,    // D = acc + bias                                   (QKV)
  EPI_RESID = 1,    // D = resid + gamma * (acc + bias)                 (proj + LayerScale + shortcut)
  EPI_GLU = 2,      // D[:, j] = (acc[2j]+b[2j]) * gelu(acc[2j+1]+b[2j+1])  (GLU, interleaved weight rows)
  EPI_SCATTER = 3,  // map[pixel(row)] = resid + gamma * (acc + bias)   (MLP out + LayerScale + residual + scatter-back)
  EPI_LSTM = 4      // conv-LSTM gates: columns interleaved [f,i,o,g] per channel; TF32 operands straight from fp32
};

struct EpiParams {
  float* out_f32;          // fp32 destination (rows or NHWC map); EPI_LSTM: h
  float* out2_f32;         // EPI_LSTM: c
  __nv_bfloat16* out_bf16; // optional bf16 copy of the destination rows (operand of the next GEMM)
  int ldo;                 // leading dimension of the row destinations
  const float* resid;      // [rows, ldr] fp32
  int ldr;
  const float* gamma;      // [N] or nullptr (identity)
  const int* row_pix;      // compacted row -> NHWC pixel index, for EPI_SCATTER
  int C;                   // channels of the NHWC map (EPI_SCATTER)
  long long* trace;        // debug stamps (sast_debug_trace), normally null
};

struct LayerWorkspace {
  float* n2f;              // [P,C]   LN2(LN1(x)) of the selected tokens (shortcut)
  __nv_bfloat16* n2h;      // [P,C]   bf16 copy (SAST_BF16)
  void* qkv;               // [P,3C]  fp32 or bf16
  void* att;               // [P,C]   fp32 or bf16
  float* yf;               // [P,C]   y = n2 + g1*proj
  __nv_bfloat16* yh;       // [P,C]
  void* hid;               // [P,I]   GLU output, fp32 or bf16
  float* mtmp;             // [P,C]   MLP output before context broadcast
  float* cbmean;           // [B,C]
};

size_t layer_workspace_layout(long long P, int C, int I, int B, int precision, void* base, LayerWorkspace* ws);

// tensor-core launchers (gemm_tc.cu / attn_tc.cu); M is read on the device from counts[1]
int launch_gemm_tc(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, int N, int K,
                   const int* counts, long long max_rows, int epi, const EpiParams& ep, cudaStream_t st);
// variant 0: tcgen05 tiles; 1: CUDA-core kernel (debug knob SAST_B200_ATTN=simt)
int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* att, int C, const sast_selection& sel, int NW, int B, int T,
                        long long max_rows, int variant, cudaStream_t st);

// fused one-kernel layer (fused_layer.cu): bf16 path, C = 64 / 128, no context broadcast
bool fused_layer_supported(const sast_layer_args& a);
int launch_layer_fused(const sast_layer_args& a, const Geom& g, cudaStream_t st);

// one kernel per layer with a group of C/64 CTAs per tile (group_layer.cu): bf16 path, C = 256 / 512
bool group_layer_supported(const sast_layer_args& a);
int launch_layer_group(const sast_layer_args& a, const Geom& g, cudaStream_t st);
size_t group_layer_workspace_bytes(long long P, int C, int I);

}  // namespace sast
