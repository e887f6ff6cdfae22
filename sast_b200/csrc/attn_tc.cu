// Attention over the compacted rows (SAST_BF16 path): per selected window and head,
// softmax(q k^T / sqrt(32)) v over the window's selected tokens.   (replaces SAST.py:219-229)
//
// v1: bf16 in / bf16 out with fp32 CUDA-core math (one thread per query row, K/V staged in
// shared memory as fp32).  The compacted buffer has no padding rows, so no column mask exists.
// TODO(round 2): tcgen05 S = Q K^T / P V tiles over greedy 128-row window groups (sel.tiles).
#include "layer.cuh"

namespace sast {

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float* o) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(p[i]);
    o[2 * i] = f.x; o[2 * i + 1] = f.y;
  }
}

__global__ void __launch_bounds__(128) attention_bf16_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                             __nv_bfloat16* __restrict__ att, int C,
                                                             const int* __restrict__ win_K,
                                                             const int* __restrict__ win_row0) {
  extern __shared__ __align__(16) float kv[];      // k [K][32] then v [K][32]
  const int w = blockIdx.x, h = blockIdx.y;
  const int K = win_K[w];
  if (K == 0) return;
  const int row0 = win_row0[w];
  const int ld = 3 * C;
  float* ks = kv;
  float* vs = kv + (size_t)K * 32;
  for (int i = threadIdx.x; i < K * 8; i += blockDim.x) {      // 8 x 16-byte chunks per row: 4 of k, 4 of v
    const int r = i >> 3, c = i & 7;
    const uint4 u = *reinterpret_cast<const uint4*>(qkv + (size_t)(row0 + r) * ld + h * 96 + 32 + c * 8);
    bf16x8_to_f32(u, c < 4 ? ks + r * 32 + c * 8 : vs + r * 32 + (c - 4) * 8);
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= K) return;
  float q[32], o[32];
  const float scale = 0.17677669529663688110f;   // 32^-0.5
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(qkv + (size_t)(row0 + i) * ld + h * 96 + c * 8);
    bf16x8_to_f32(u, q + c * 8);
  }
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
  float mx = -INFINITY, l = 0.f;
  for (int j = 0; j < K; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) s = fmaf(q[d], ks[j * 32 + d], s);
    s *= scale;
    const float mn = fmaxf(mx, s);
    const float corr = __expf(mx - mn), p = __expf(s - mn);
    l = l * corr + p;
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = fmaf(p, vs[j * 32 + d], o[d] * corr);
    mx = mn;
  }
  const float il = 1.0f / l;
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] *= il;
  __nv_bfloat16* dst = att + (size_t)(row0 + i) * C + h * 32;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 pk;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(o[c * 8 + 0], o[c * 8 + 1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(o[c * 8 + 2], o[c * 8 + 3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(o[c * 8 + 4], o[c * 8 + 5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(o[c * 8 + 6], o[c * 8 + 7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(dst + c * 8) = pk;
  }
}

int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* att, int C, const sast_selection& sel, int NW, int T,
                        cudaStream_t st) {
  const size_t smem = (size_t)T * 64 * sizeof(float);
  attention_bf16_kernel<<<dim3(NW, C / 32), 128, smem, st>>>(qkv, att, C, sel.win_K, sel.win_row0);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

}  // namespace sast
