// a1: scene sparsity ratio r  (replaces sast_rnn.py:45-60 non_zero_ratio).
//
// The planes are cut in bands of 32 pixel rows (= 8 rows of 4x4 level-0 cells = one row of level-3 cells),
// one CTA per (plane, band): every thread max-reduces 4x4 cells with
// 128-bit / 32-bit row loads (coalesced: neighbouring threads own neighbouring cells), the
// level-0 maxima go to shared memory and levels 1..3 (8x8, 16x16, 32x32 pixels) are
// max-reduced from there.  "Cell != 0" is counted per level exactly like the cascaded
// F.max_pool2d of the reference (floor semantics on ragged sizes, true max so that a cell of
// {-1, 0} counts as zero), the counts wrap to int16 as the reference's accumulator does, and
// the result is fp32(B / numel(pooled)) * fp32(count): bit-exact.
#include "common.cuh"
#include <cuda_fp16.h>

namespace sast {

template <typename T> struct Load4;
template <> struct Load4<uint8_t> {
  static __device__ __forceinline__ void ld(const uint8_t* p, bool vec, float v[4]) {
    if (vec) {
      const uchar4 u = *reinterpret_cast<const uchar4*>(p);
      v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
    } else {
      v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
    }
  }
};
template <> struct Load4<int32_t> {
  static __device__ __forceinline__ void ld(const int32_t* p, bool vec, float v[4]) {
    if (vec) {
      const int4 u = *reinterpret_cast<const int4*>(p);
      v[0] = (float)u.x; v[1] = (float)u.y; v[2] = (float)u.z; v[3] = (float)u.w;
    } else {
      v[0] = (float)p[0]; v[1] = (float)p[1]; v[2] = (float)p[2]; v[3] = (float)p[3];
    }
  }
};
template <> struct Load4<float> {
  static __device__ __forceinline__ void ld(const float* p, bool vec, float v[4]) {
    if (vec) {
      const float4 u = *reinterpret_cast<const float4*>(p);
      v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
    } else {
      v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
    }
  }
};

// One CTA per (plane, band of 32 pixel rows): 1920 CTAs at 1 Mpx B=8 instead of 160 plane-sized ones.  Integer
// counts are accumulated with atomicAdd (order independent, hence still bit-exact) into a zeroed scratch array;
// a second tiny kernel turns them into ratios.
template <typename T>
__global__ void __launch_bounds__(256) nonzero_count_kernel(const T* __restrict__ x, int H, int W, int* __restrict__ counts) {
  pdl_entry();
  extern __shared__ float cell0[];        // [8][w0] level-0 maxima of this band
  __shared__ int red[4][8];
  const int plane = blockIdx.x, band = blockIdx.y;
  const T* xp = x + (size_t)plane * H * W;
  const int h0 = H / 4, w0 = W / 4;
  const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  int cnt[4] = {0, 0, 0, 0};
  const int rows0 = min(8, h0 - band * 8);          // level-0 cell rows in this band
  for (int i = threadIdx.x; i < rows0 * w0; i += blockDim.x) {
    const int cy = i / w0, cx = i - cy * w0;
    const T* p = xp + (size_t)((band * 8 + cy) * 4) * W + cx * 4;
    float v[4][4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) Load4<T>::ld(p + (size_t)rr * W, vec, v[rr]);     // 4 independent row loads in flight
    float m = -INFINITY;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) m = fmaxf(m, fmaxf(fmaxf(v[rr][0], v[rr][1]), fmaxf(v[rr][2], v[rr][3])));
    cell0[cy * w0 + cx] = m;
    cnt[0] += (m != 0.0f);
  }
  __syncthreads();
#pragma unroll
  for (int lvl = 1; lvl < 4; ++lvl) {
    const int s = 1 << lvl;                           // level-0 cells per side
    const int rows = rows0 / s, cols = w0 / s;        // complete cells only (floor)
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
      const int cy = i / cols, cx = i - cy * cols;
      float m = -INFINITY;
      for (int yy = 0; yy < s; ++yy)
        for (int xx = 0; xx < s; ++xx) m = fmaxf(m, cell0[(cy * s + yy) * w0 + cx * s + xx]);
      cnt[lvl] += (m != 0.0f);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int v = warp_sum_i(cnt[lvl]);
    if (lane == 0) red[lvl][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    int tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[threadIdx.x][w];
    if (tot) atomicAdd(&counts[plane * 4 + threadIdx.x], tot);
  }
}


// Bit-packed event histograms (sast_unpack_nonzero_ratio): the same band decomposition, reading BITS (1 or 4) bits per
// bin and writing the bins back as uint8 for the stem -- the unpack rides on the pass that has to read the input anyway.
// Packing runs along x, little endian: BITS = 1: bit k of byte j is column 8 j + k; BITS = 4: the low nibble of byte j is
// column 2 j, the high nibble column 2 j + 1.
template <int BITS>
__global__ void __launch_bounds__(256) unpack_count_kernel(const uint8_t* __restrict__ packed, int H, int W,
                                                            uint8_t* __restrict__ out, int* __restrict__ counts) {
  pdl_entry();
  extern __shared__ float cell0[];        // [8][w0] level-0 maxima of this band
  __shared__ int red[4][8];
  const int plane = blockIdx.x, band = blockIdx.y;
  const int pitch = W * BITS / 8;         // bytes per packed row
  const uint8_t* xp = packed + (size_t)plane * H * pitch;
  uint8_t* op = out + (size_t)plane * H * W;
  const int h0 = H / 4, w0 = W / 4;
  int cnt[4] = {0, 0, 0, 0};
  const int rows0 = min(8, h0 - band * 8);
  for (int i = threadIdx.x; i < rows0 * w0; i += blockDim.x) {
    const int cy = i / w0, cx = i - cy * w0;
    const int y0 = (band * 8 + cy) * 4;
    uint32_t w[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {        // 4 independent row loads in flight; 4 bins -> one uchar4
      const uint8_t* row = xp + (size_t)(y0 + rr) * pitch;
      if (BITS == 1) {
        const uint32_t nib = (row[cx >> 1] >> ((cx & 1) * 4)) & 15u;
        w[rr] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
      } else {
        const uint32_t h = *reinterpret_cast<const uint16_t*>(row + cx * 2);
        w[rr] = (h & 15u) | ((h & 0xF0u) << 4) | ((h & 0xF00u) << 8) | ((h & 0xF000u) << 12);
      }
    }
    uint32_t any = 0;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      if (out) *reinterpret_cast<uint32_t*>(op + (size_t)(y0 + rr) * W + cx * 4) = w[rr];
      any |= w[rr];
    }
    const float m = any ? 1.0f : 0.0f;      // bins are >= 0: "max != 0" is "any bin set"
    cell0[cy * w0 + cx] = m;
    cnt[0] += any != 0;
  }
  // rows of the plane below the last complete 4-row cell (H % 4) hold no cell but must still be unpacked
  if (out && band == gridDim.y - 1) {
    for (int y = h0 * 4; y < H; ++y)
      for (int x = threadIdx.x; x < W; x += blockDim.x)
        op[(size_t)y * W + x] = BITS == 1 ? ((xp[(size_t)y * pitch + (x >> 3)] >> (x & 7)) & 1u)
                                          : ((xp[(size_t)y * pitch + (x >> 1)] >> ((x & 1) * 4)) & 15u);
  }
  __syncthreads();
#pragma unroll
  for (int lvl = 1; lvl < 4; ++lvl) {
    const int s = 1 << lvl;
    const int rows = rows0 / s, cols = w0 / s;
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
      const int cy = i / cols, cx = i - cy * cols;
      float m = 0.f;
      for (int yy = 0; yy < s; ++yy)
        for (int xx = 0; xx < s; ++xx) m = fmaxf(m, cell0[(cy * s + yy) * w0 + cx * s + xx]);
      cnt[lvl] += (m != 0.0f);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int v = warp_sum_i(cnt[lvl]);
    if (lane == 0) red[lvl][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    int tot = 0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) tot += red[threadIdx.x][w2];
    if (tot) atomicAdd(&counts[plane * 4 + threadIdx.x], tot);
  }
}

// Count-only pass over a 1-bit packed histogram (W % 32 == 0): one thread per (band of 32 rows, 32-bit word column) ORs the
// four rows of each level-0 cell row; a level-0 / 1 / 2 / 3 cell (4 / 8 / 16 / 32 columns) is then a nibble / byte / half /
// word of the OR of 1 / 2 / 4 / 8 such cell rows, so "cell != 0" is a few bit operations and a popcount.  Same counts as
// nonzero_count_kernel (complete cells only: floor semantics on ragged heights), 32 independent coalesced loads per thread.
// One CTA per plane, so the ratios are written directly (what nonzero_finalize_kernel does for the other formats).
__global__ void __launch_bounds__(256) packed1_count_kernel(const uint32_t* __restrict__ packed, int H, int W, int B, int Cin, float f0,
                                                             float f1, float f2, float f3, float* __restrict__ r) {
  pdl_entry();
  __shared__ int red[4][8];
  const int plane = blockIdx.x;
  const int words = W / 32, h0 = H / 4, bands = (h0 + 7) / 8;
  const uint32_t* xp = packed + (size_t)plane * H * words;
  int cnt[4] = {0, 0, 0, 0};
  for (int i = threadIdx.x; i < bands * words; i += blockDim.x) {
    const int band = i / words, w = i - band * words;
    const int rows0 = min(8, h0 - band * 8);
    uint32_t o[8];
#pragma unroll
    for (int cy = 0; cy < 8; ++cy) {
      uint32_t v = 0;
      if (cy < rows0) {
        const uint32_t* r = xp + (size_t)((band * 8 + cy) * 4) * words + w;
        v = __ldg(r) | __ldg(r + words) | __ldg(r + 2 * words) | __ldg(r + 3 * words);
      }
      o[cy] = v;
    }
#pragma unroll
    for (int cy = 0; cy < 8; ++cy) {                        // (rows beyond rows0 are zero: they count nothing)
      uint32_t u = o[cy] | (o[cy] >> 1);
      u |= u >> 2;
      cnt[0] += __popc(u & 0x11111111u);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (2 * j + 1 < rows0) {
        uint32_t u = o[2 * j] | o[2 * j + 1];
        u |= u >> 1; u |= u >> 2; u |= u >> 4;
        cnt[1] += __popc(u & 0x01010101u);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (4 * j + 3 < rows0) {
        const uint32_t u = o[4 * j] | o[4 * j + 1] | o[4 * j + 2] | o[4 * j + 3];
        cnt[2] += ((u & 0xFFFFu) != 0) + ((u >> 16) != 0);
      }
    }
    if (rows0 == 8) cnt[3] += (o[0] | o[1] | o[2] | o[3] | o[4] | o[5] | o[6] | o[7]) != 0;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int v = warp_sum_i(cnt[lvl]);
    if (lane == 0) red[lvl][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {       // one CTA owns the whole plane: its totals ARE the plane's counts -- no scratch, no finalize launch
    int tot = 0;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) tot += red[threadIdx.x][w2];
    const int lvl = threadIdx.x, b = plane / Cin, c = plane - b * Cin;
    const float f = lvl == 0 ? f0 : lvl == 1 ? f1 : lvl == 2 ? f2 : f3;
    r[((size_t)lvl * B + b) * Cin + c] = f * (float)(int16_t)tot;            // int16 wrap as in the reference
  }
}

// Event histogram (bit-packed, BITS = 1 / 4, or plain uint8, BITS = 8; NCHW) -> fp16 NHWC with the stem's replicate
// padding materialised: xh[b][yy][xx][c], yy in [0, H+8) = source row clamp(yy-3), xx in [0, W+8) = source column
// clamp(xx-3); rows >= H+3 / columns >= W+3 are never multiplied by a non-zero weight and are written as zeros (they
// must be finite).  One CTA per padded row: the Cin packed source rows are staged in shared memory, a thread builds two
// neighbouring pixels (2 x Cin halves = 80 bytes at Cin = 20) and stores them as 16-byte words.  This is the layout
// the TMA-fed stem (stem_nhwc.cu) reads its A operand from: the 7 x Cin window of an output pixel is contiguous.
template <int BITS, int CIN>
__global__ void __launch_bounds__(256) events_nhwc_kernel(const uint8_t* __restrict__ src, int H, int W, uint16_t* __restrict__ xh) {
  pdl_entry();
  static_assert(CIN % 4 == 0, "a pixel pair is a whole number of 16-byte words");
  constexpr int QP = 2 * CIN / 8;         // 16-byte words per pixel pair (5 at CIN = 20: lane stride 5 words -> conflict-free STS.128)
  extern __shared__ __align__(16) uint8_t ev_smem[];
  const int yy = blockIdx.x, b = blockIdx.y;
  const int Hp = H + 8, Wp = W + 8;
  const int pitch = W * BITS / 8;
  uint8_t* const rows_s = ev_smem;                                    // [CIN][pitch] packed source rows
  uint4* const out_s = reinterpret_cast<uint4*>(ev_smem + (size_t)CIN * pitch);      // the padded output row, Wp * CIN halves
  const bool live = yy < H + 3;
  const int iy = min(max(yy - 3, 0), H - 1);
  if (live) {
    const int words = pitch / 4;          // W % 32 == 0: whole 32-bit words
    for (int i = threadIdx.x; i < CIN * words; i += blockDim.x) {
      const int c = i / words, w = i - c * words;
      reinterpret_cast<uint32_t*>(rows_s)[i] = __ldg(reinterpret_cast<const uint32_t*>(src + (((size_t)b * CIN + c) * H + iy) * pitch) + w);
    }
  }
  __syncthreads();
  for (int xp = threadIdx.x; xp < Wp / 2; xp += blockDim.x) {
    uint32_t hv[2 * CIN];                 // [pixel of the pair][channel] as fp16 bit patterns (exact: counts <= 255)
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int xx = 2 * xp + px;
      const bool ok = live && xx < W + 3;
      const int ix = min(max(xx - 3, 0), W - 1);
      const int byte = BITS == 1 ? ix >> 3 : BITS == 4 ? ix >> 1 : ix;
      const int sh = BITS == 1 ? ix & 7 : BITS == 4 ? (ix & 1) * 4 : 0;
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const uint32_t v = ok ? ((uint32_t)rows_s[c * pitch + byte] >> sh) & (BITS == 1 ? 1u : BITS == 4 ? 15u : 255u) : 0u;
        hv[px * CIN + c] = BITS == 1 ? v * 0x3C00u : (uint32_t)__half_as_ushort(__uint2half_rn(v));
      }
    }
#pragma unroll
    for (int q = 0; q < QP; ++q)
      out_s[xp * QP + q] = make_uint4(hv[8 * q] | (hv[8 * q + 1] << 16), hv[8 * q + 2] | (hv[8 * q + 3] << 16),
                                      hv[8 * q + 4] | (hv[8 * q + 5] << 16), hv[8 * q + 6] | (hv[8 * q + 7] << 16));
  }
  __syncthreads();
  uint4* const orow = reinterpret_cast<uint4*>(xh + ((size_t)b * Hp + yy) * Wp * CIN);
  for (int i = threadIdx.x; i < Wp / 2 * QP; i += blockDim.x) orow[i] = out_s[i];      // dense 16-byte stores
}

__global__ void nonzero_finalize_kernel(int* __restrict__ counts, int planes, int B, int Cin, float f0, float f1, float f2, float f3,
                                        float* __restrict__ r) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= planes * 4) return;
  const int plane = i >> 2, lvl = i & 3;
  const int b = plane / Cin, c = plane - b * Cin;
  const float f = lvl == 0 ? f0 : lvl == 1 ? f1 : lvl == 2 ? f2 : f3;
  r[((size_t)lvl * B + b) * Cin + c] = f * (float)(int16_t)counts[i];       // int16 wrap as in the reference
  counts[i] = 0;                                                             // leave the scratch zeroed for the next call
}

}  // namespace sast

extern "C" int sast_nonzero_ratio(const void* x, int32_t dtype, int32_t B, int32_t Cin, int32_t H, int32_t W,
                                  float* r, int32_t* scratch, void* stream) {
  SAST_CHECK_PTR(x); SAST_CHECK_PTR(r); SAST_CHECK_PTR(scratch);
  if (B <= 0 || Cin <= 0 || H < 32 || W < 32) return SAST_E_SHAPE;
  float f[4];
  for (int l = 0; l < 4; ++l) {
    const long long cs = 4ll << l;
    const double numel = (double)B * Cin * (H / cs) * (W / cs);
    f[l] = (float)((double)B / numel);
  }
  const size_t smem = (size_t)8 * (W / 4) * sizeof(float);
  if (smem > 48 * 1024) return SAST_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int planes = B * Cin;
  const dim3 grid(planes, (H / 4 + 7) / 8), block(256);
  switch (dtype) {
    case SAST_U8: sast::launch_k(sast::nonzero_count_kernel<uint8_t>, grid, block, smem, st, (const uint8_t*)x, H, W, scratch); break;
    case SAST_I32: sast::launch_k(sast::nonzero_count_kernel<int32_t>, grid, block, smem, st, (const int32_t*)x, H, W, scratch); break;
    case SAST_F32: sast::launch_k(sast::nonzero_count_kernel<float>, grid, block, smem, st, (const float*)x, H, W, scratch); break;
    default: return SAST_E_UNSUPPORTED;
  }
  SAST_LAUNCH_CHECK();
  sast::launch_k(sast::nonzero_finalize_kernel, dim3((planes * 4 + 127) / 128), dim3(128), 0, st, scratch, planes, B, Cin, f[0], f[1], f[2],
                 f[3], r);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

// Bit-packed twin of sast_nonzero_ratio: unpacks to uint8 [B,Cin,H,W] (the stem's input) and computes r in the same pass.
extern "C" int sast_unpack_nonzero_ratio(const uint8_t* packed, int32_t bits, int32_t B, int32_t Cin, int32_t H, int32_t W,
                                         uint8_t* x_out, float* r, int32_t* scratch, void* stream) {
  SAST_CHECK_PTR(packed); SAST_CHECK_PTR(x_out); SAST_CHECK_PTR(r); SAST_CHECK_PTR(scratch);
  if (B <= 0 || Cin <= 0 || H < 32 || W < 32 || W % 8 != 0) return SAST_E_SHAPE;
  if (bits != 1 && bits != 4) return SAST_E_UNSUPPORTED;
  float f[4];
  for (int l = 0; l < 4; ++l) {
    const long long cs = 4ll << l;
    const double numel = (double)B * Cin * (H / cs) * (W / cs);
    f[l] = (float)((double)B / numel);
  }
  const size_t smem = (size_t)8 * (W / 4) * sizeof(float);
  if (smem > 48 * 1024) return SAST_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int planes = B * Cin;
  const dim3 grid(planes, (H / 4 + 7) / 8), block(256);
  if (bits == 1) sast::launch_k(sast::unpack_count_kernel<1>, grid, block, smem, st, packed, H, W, x_out, scratch);
  else sast::launch_k(sast::unpack_count_kernel<4>, grid, block, smem, st, packed, H, W, x_out, scratch);
  SAST_LAUNCH_CHECK();
  sast::launch_k(sast::nonzero_finalize_kernel, dim3((planes * 4 + 127) / 128), dim3(128), 0, st, scratch, planes, B, Cin, f[0], f[1], f[2],
                 f[3], r);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

// Input pass of the TMA-fed stem (sast_stem_nhwc_fwd): src is the event histogram, bit-packed (bits = 1 / 4, as
// sast_unpack_nonzero_ratio takes it) or plain uint8 (bits = 8), NCHW.  Writes xh = fp16 [B, H+8, W+8, Cin] (NHWC, replicate
// padding of 3 materialised, see events_nhwc_kernel) and, when r != NULL, the scene sparsity ratios r [4,B,Cin] exactly as
// sast_nonzero_ratio does (scratch: B*Cin*4 zeroed int32).  W % 32 == 0, Cin == 20.
extern "C" int sast_events_nhwc(const uint8_t* src, int32_t bits, int32_t B, int32_t Cin, int32_t H, int32_t W, uint16_t* xh,
                                float* r, int32_t* scratch, void* stream) {
  SAST_CHECK_PTR(src);
  if (!xh && !r) return SAST_E_NULL;                  // xh == NULL: only the ratios (the stem reads the packed bits itself)
  if (B <= 0 || Cin <= 0 || H < 32 || W < 32 || W % 32 != 0) return SAST_E_SHAPE;
  if ((bits != 1 && bits != 4 && bits != 8) || Cin != 20) return SAST_E_UNSUPPORTED;      // the stem it feeds is built for 20 event bins
  if ((reinterpret_cast<uintptr_t>(src) & 3) || (reinterpret_cast<uintptr_t>(xh) & 15)) return SAST_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (r) {
    if (bits != 1) SAST_CHECK_PTR(scratch);               // the 1-bit kernel needs no scratch
    float f[4];
    for (int l = 0; l < 4; ++l) {
      const long long cs = 4ll << l;
      f[l] = (float)((double)B / ((double)B * Cin * (H / cs) * (W / cs)));
    }
    const size_t smem = (size_t)8 * (W / 4) * sizeof(float);
    if (smem > 48 * 1024) return SAST_E_UNSUPPORTED;
    const int planes = B * Cin;
    const dim3 grid(planes, (H / 4 + 7) / 8), block(256);
    if (bits == 1) {
      sast::launch_k(sast::packed1_count_kernel, dim3(planes), block, 0, st, (const uint32_t*)src, H, W, B, Cin, f[0], f[1], f[2], f[3], r);
      SAST_LAUNCH_CHECK();
    } else {
    if (bits == 4) sast::launch_k(sast::unpack_count_kernel<4>, grid, block, smem, st, src, H, W, (uint8_t*)nullptr, scratch);
    else sast::launch_k(sast::nonzero_count_kernel<uint8_t>, grid, block, smem, st, src, H, W, scratch);
    SAST_LAUNCH_CHECK();
    sast::launch_k(sast::nonzero_finalize_kernel, dim3((planes * 4 + 127) / 128), dim3(128), 0, st, scratch, planes, B, Cin, f[0],
                   f[1], f[2], f[3], r);
    SAST_LAUNCH_CHECK();
    }
  }
  if (!xh) return SAST_OK;
  const size_t smem2 = (size_t)Cin * W * bits / 8 + (size_t)(W + 8) * Cin * 2;
  if (smem2 > 48 * 1024) return SAST_E_UNSUPPORTED;
  const dim3 grid2(H + 8, B), block2(256);
  if (bits == 1) sast::launch_k(sast::events_nhwc_kernel<1, 20>, grid2, block2, smem2, st, src, H, W, xh);
  else if (bits == 4) sast::launch_k(sast::events_nhwc_kernel<4, 20>, grid2, block2, smem2, st, src, H, W, xh);
  else sast::launch_k(sast::events_nhwc_kernel<8, 20>, grid2, block2, smem2, st, src, H, W, xh);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}
