// Stem of the backbone as ONE kernel (replaces x.float() + replicate padding +
// Conv2d(k=7, s=4, no bias) + NCHW->NHWC + LayerNorm, ops.py:54-91 / sast_rnn.py:153):
//   uint8 event histogram [B,Cin,H,W] (NCHW)  ->  LayerNorm(conv(x)) as fp32 NHWC [B,H/4,W/4,64].
//
// Implicit GEMM on tcgen05: M = output pixels (128 per tile), N = Cout, K ordered (c, ky, kx) with kx padded
// 7 -> 8 so that one (c, ky) group is 8 consecutive input bytes = one 16-byte fp16 chunk of the A operand.
// Event counts are small integers: exact in fp16.  The fp32 weights are split w = w_hi + w_lo (both fp16,
// |w - w_hi - w_lo| <= ~2^-22 |w| for the weight magnitudes of a conv layer) and both halves are multiplied, so
// the result is fp32-grade (tighter than the TF32 path cuDNN takes by default) at 16-bit tensor-core speed.
//
// Persistent CTAs, 14 warps:
//   warps 0-7   A producers: two threads per output pixel, four (c,ky) groups each per k-block: three coalesced
//               32-bit loads (L1-resident input rows) -> 8 bytes -> 8 fp16 -> one STS.128 into the SWIZZLE_128B tile;
//               an elected lane of warp 0 also TMA-loads the packed weight tiles
//   warp 8      TMEM allocator + MMA issuer (whole warp on uniform values, elected lane issues; two accumulators)
//   warp 9      idle (keeps the epilogue warps on TMEM lane quarters 2,3,0,1)
//   warps 10-13 epilogue: LayerNorm over the Cout channels of each pixel straight from TMEM, swizzled
//               shared-memory transpose, dense 128-byte row stores
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_fp16.h>

namespace sast {

constexpr int ST_BM = 128, ST_STAGES = 4;
constexpr int ST_THREADS = 14 * 32;

struct StSmem {
  uint64_t full[ST_STAGES];
  uint64_t empty[ST_STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// four bytes (one 32-bit word, little endian) -> two packed f16x2 words, exactly: the f16 with bits 0x6400|b is
// 1024 + b, so one PRMT builds two such halves and one HSUB2 removes the 1024 (2 instructions per 2 bytes)
__device__ __forceinline__ void bytes4_to_f16x4(uint32_t w, uint32_t& lo, uint32_t& hi) {
  const uint32_t a = __byte_perm(w, 0x64646464u, 0x4140), b = __byte_perm(w, 0x64646464u, 0x4342);
  const __half2 k = __half2half2(__ushort_as_half((unsigned short)0x6400));
  const __half2 ra = __hsub2(*reinterpret_cast<const __half2*>(&a), k), rb = __hsub2(*reinterpret_cast<const __half2*>(&b), k);
  lo = *reinterpret_cast<const uint32_t*>(&ra);
  hi = *reinterpret_cast<const uint32_t*>(&rb);
}

__global__ void __launch_bounds__(ST_THREADS, 1) stem_tc_kernel(const __grid_constant__ CUtensorMap map_whi,
                                                                const __grid_constant__ CUtensorMap map_wlo,
                                                                const uint8_t* __restrict__ x, int B, int Cin, int H, int W,
                                                                int Cout, int n_groups_pad, const float* __restrict__ ln_w,
                                                                const float* __restrict__ ln_b, float eps,
                                                                float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(16) float stage_smem[4][32 * 32];
  const int Ho = H / 4, Wo = W / 4;
  const int P = B * Ho * Wo;
  const int total_tiles = (P + ST_BM - 1) / ST_BM;
  const uint32_t a_bytes = ST_BM * 128, w_bytes = (uint32_t)Cout * 128;
  const uint32_t stage_bytes = a_bytes + 2 * w_bytes;                 // A | W_hi | W_lo
  uint8_t* base = smem_raw;
  StSmem* sm = reinterpret_cast<StSmem*>(smem_raw + (size_t)ST_STAGES * stage_bytes);
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;    // warp: provably uniform
  const int nkb = n_groups_pad / 8;                                    // k-blocks of 8 groups x 8 taps
  const int n_groups = Cin * 7;

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&map_whi);
    ptx::tma_prefetch_desc(&map_wlo);
    for (int s = 0; s < ST_STAGES; ++s) { ptx::mbar_init(&sm->full[s], 9 /* 8 producer warps + the TMA expect_tx arrive */); ptx::mbar_init(&sm->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&sm->tmem_full[a], 1); ptx::mbar_init(&sm->tmem_empty[a], 4); }
    ptx::fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * Cout)) tmem_cols <<= 1;
  if (warp == 8) ptx::tmem_alloc(&sm->tmem_base, tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sm->tmem_base;
  // PDL: everything above (barriers, TMEM, index loads of data written >= 2 kernels ago) overlapped the tail of the
  // preceding kernel; its output is read only from here on
  pdl_entry();

  if (warp < 8) {
    // ---------------- A producers ----------------
    const int t = threadIdx.x;                 // 0..255
    const int m = t & 127;                     // pixel row of the tile
    const int gh = t >> 7;                     // which four of the eight groups of a k-block
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int p = tile * ST_BM + m;
      const bool ok = p < P;
      const int pp = ok ? p : P - 1;
      const int b = pp / (Ho * Wo), rem = pp - b * (Ho * Wo);
      const int oy = rem / Wo, ox = rem - oy * Wo;
      const uint8_t* xb = x + (size_t)b * Cin * H * W;
      // input words of one k-block: 4 groups x {left, centre, right} aligned words; the NEXT k-block's words are
      // requested before the current ones are converted (software pipelining: the loads never sit on the
      // critical path of the conversion)
      uint32_t w0[4], w1[4], w2[4], n0[4], n1[4], n2[4], m0[4], m1[4], m2[4];
      auto load_block = [&](int kb, uint32_t (&a0)[4], uint32_t (&a1)[4], uint32_t (&a2)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int gidx = kb * 8 + gh * 4 + j;
          if (gidx >= n_groups) gidx = n_groups - 1;          // padding groups: finite data, zero weights
          const int c = gidx / 7, ky = gidx - c * 7;
          const int iy = min(max(4 * oy - 3 + ky, 0), H - 1);  // replicate padding (rows)
          const uint32_t* row = reinterpret_cast<const uint32_t*>(xb + ((size_t)c * H + iy) * W);
          // three aligned words; nothing here may depend on a loaded value (the loads must stay in flight):
          // the left-edge replicate is patched in at conversion time
          a1[j] = row[ox];                                      // columns 4ox .. 4ox+3
          a0[j] = row[ox > 0 ? ox - 1 : ox];
          a2[j] = row[ox + 1 < Wo ? ox + 1 : ox];               // only tap 7 (zero weight) reads it
        }
      };
      // prefetch distance 2: the producers are bound by the latency of these loads (first touch of the input comes from
      // DRAM), not by instruction issue, so two k-blocks of words are kept in flight behind the one being converted
      load_block(0, w0, w1, w2);
      if (nkb > 1) load_block(1, n0, n1, n2);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % ST_STAGES, round = it / ST_STAGES;
        if (kb + 2 < nkb) load_block(kb + 2, m0, m1, m2);
        ptx::mbar_wait(&sm->empty[s], (round & 1) ^ 1);
        uint8_t* st = base + (size_t)s * stage_bytes;
        if (warp == 0 && ptx::elect_one()) {     // warp-uniform operands, elected lane: no R2UR waterfall per TMA
          ptx::mbar_arrive_expect_tx(&sm->full[s], 2 * w_bytes);
          ptx::tma_load_2d(st + a_bytes, &map_whi, &sm->full[s], kb * 64, 0);
          ptx::tma_load_2d(st + a_bytes + w_bytes, &map_wlo, &sm->full[s], kb * 64, 0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // taps kx = 0..7 are input columns 4ox-3 .. 4ox+4: bytes 1..3 of w0, all of w1, byte 0 of w2
          if (ox == 0) w0[j] = (w1[j] & 0xFFu) * 0x01010101u;          // replicate padding (left edge)
          const uint32_t lo4 = __byte_perm(w0[j], w1[j], 0x4321);      // columns 4ox-3 .. 4ox
          const uint32_t hi4 = __byte_perm(w1[j], w2[j], 0x4321);      // columns 4ox+1 .. 4ox+4
          uint4 v;
          bytes4_to_f16x4(lo4, v.x, v.y);
          bytes4_to_f16x4(hi4, v.z, v.w);
          const int g = gh * 4 + j;
          *reinterpret_cast<uint4*>(st + (uint32_t)(m >> 3) * 1024 + (uint32_t)(m & 7) * 128 + (uint32_t)((g ^ (m & 7)) * 16)) = v;
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm->full[s]);        // one arrive per warp: 256 serialized smem atomics per k-block otherwise
#pragma unroll
        for (int j = 0; j < 4; ++j) { w0[j] = n0[j]; w1[j] = n1[j]; w2[j] = n2[j]; n0[j] = m0[j]; n1[j] = m1[j]; n2[j] = m2[j]; }
      }
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer: whole warp on uniform values, one elected lane issues (no R2UR waterfall per UTCHMMA) ----------------
    const bool leader = ptx::elect_one();
    const uint32_t idesc = (1u << 4) | (((uint32_t)Cout >> 3) << 17) | ((ST_BM >> 4) << 24);   // kind::f16: A = B = F16, D = F32, K-major
    uint32_t it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti & 1, use = ti >> 1;
      ptx::mbar_wait(&sm->tmem_empty[acc], (use & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)Cout;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % ST_STAGES, round = it / ST_STAGES;
        ptx::mbar_wait(&sm->full[s], round & 1);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(base + (size_t)s * stage_bytes);
        const uint64_t da = ptx::umma_desc_sw128_kmajor(sa);
        const uint64_t dh = ptx::umma_desc_sw128_kmajor(sa + a_bytes), dl = ptx::umma_desc_sw128_kmajor(sa + a_bytes + w_bytes);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t o = (uint64_t)(k * 2);
            ptx::umma_f16_ss(tmem_d, da + o, dl + o, idesc, (kb | k) ? 1u : 0u);      // low half first
            ptx::umma_f16_ss(tmem_d, da + o, dh + o, idesc, 1u);
          }
          ptx::umma_commit(&sm->empty[s]);
        }
      }
      if (leader) ptx::umma_commit(&sm->tmem_full[acc]);
    }
  } else if (warp >= 10) {
    // ---------------- epilogue: LayerNorm over channels (thread = pixel row), transposed store ----------------
    const int quarter = warp & 3;
    float* stage = &stage_smem[warp - 10][0];
    const int r_sub = lane >> 3, gq = lane & 7, c4 = gq * 4;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti & 1, use = ti >> 1;
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)Cout + ((uint32_t)(quarter * 32) << 16);
      ptx::mbar_wait(&sm->tmem_full[acc], use & 1);
      ptx::tc_fence_after();
      // pass 1: statistics of this thread's row (two-pass variance over the TMEM-resident row)
      float sum = 0.f;
      for (int c0 = 0; c0 < Cout; c0 += 32) {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tmem_d + (uint32_t)c0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += __uint_as_float(raw[j]);
      }
      const float mean = sum / (float)Cout;
      float ss = 0.f;
      for (int c0 = 0; c0 < Cout; c0 += 32) {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tmem_d + (uint32_t)c0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = __uint_as_float(raw[j]) - mean; ss += d * d; }
      }
      const float rstd = rsqrtf(ss / (float)Cout + eps);
      // pass 2: normalise, transpose through shared memory, dense row stores
      for (int c0 = 0; c0 < Cout; c0 += 32) {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tmem_d + (uint32_t)c0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 v;
          v.x = (__uint_as_float(raw[4 * k]) - mean) * rstd; v.y = (__uint_as_float(raw[4 * k + 1]) - mean) * rstd;
          v.z = (__uint_as_float(raw[4 * k + 2]) - mean) * rstd; v.w = (__uint_as_float(raw[4 * k + 3]) - mean) * rstd;
          *reinterpret_cast<float4*>(stage + lane * 32 + ((k ^ (lane & 7)) << 2)) = v;
        }
        __syncwarp();
        const int n = c0 + c4;
        float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ln_w) g4 = __ldg(reinterpret_cast<const float4*>(ln_w + n));
        if (ln_b) b4 = __ldg(reinterpret_cast<const float4*>(ln_b + n));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + r_sub;
          const int p = tile * ST_BM + quarter * 32 + r;
          if (p >= P) continue;
          const float4 a4 = *reinterpret_cast<const float4*>(stage + r * 32 + ((gq ^ (r & 7)) << 2));
          *reinterpret_cast<float4*>(out + (size_t)p * Cout + n) =
              make_float4(a4.x * g4.x + b4.x, a4.y * g4.y + b4.y, a4.z * g4.z + b4.z, a4.w * g4.w + b4.w);
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm->tmem_empty[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

int make_tmap_bf16_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows,
                       int swizzle_bytes);

}  // namespace sast

// x [B,Cin,H,W] uint8 NCHW -> out [B,H/4,W/4,Cout] fp32 NHWC = LayerNorm(conv7x7 stride 4, replicate padding 3, no bias).
// w_hi / w_lo: fp16 [Cout, Kp] with K ordered (c, ky, kx padded to 8), Kp = n_groups_pad*8, zero for padding entries;
// w_hi + w_lo ~= conv.weight (fp16 split).  H, W multiples of 4; Cout multiple of 32, <= 256.
extern "C" int sast_stem_fwd(const uint8_t* x, int32_t B, int32_t Cin, int32_t H, int32_t W, const uint16_t* w_hi,
                             const uint16_t* w_lo, int32_t Cout, int32_t n_groups_pad, const float* ln_w, const float* ln_b,
                             float eps, float* out, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(x); SAST_CHECK_PTR(w_hi); SAST_CHECK_PTR(w_lo); SAST_CHECK_PTR(out);
  if (B <= 0 || Cin <= 0 || H < 8 || W < 8 || H % 4 || W % 4 || Cout % 32 || Cout > 256 || Cout <= 0) return SAST_E_SHAPE;
  if (n_groups_pad % 8 != 0 || n_groups_pad < Cin * 7) return SAST_E_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) & 3) != 0) return SAST_E_SHAPE;
  const int Kp = n_groups_pad * 8;
  CUtensorMap mh, ml;
  int rc = make_tmap_bf16_box(&mh, w_hi, Cout, Kp, Kp, 64, Cout, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_box(&ml, w_lo, Cout, Kp, Kp, 64, Cout, 128);
  if (rc) return rc;
  const size_t smem = (size_t)ST_STAGES * (ST_BM * 128 + 2 * (size_t)Cout * 128) + sizeof(StSmem) + 64;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  if (smem > 200 * 1024) return SAST_E_UNSUPPORTED;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long P = (long long)B * (H / 4) * (W / 4);
  if (P >= (1ll << 31) / 256) return SAST_E_UNSUPPORTED;
  const long long tiles = (P + ST_BM - 1) / ST_BM;
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  launch_k(stem_tc_kernel, dim3(grid), dim3(ST_THREADS), smem, (cudaStream_t)stream, mh, ml, x, B, Cin, H, W, Cout, n_groups_pad,
           ln_w, ln_b, eps, out);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}
