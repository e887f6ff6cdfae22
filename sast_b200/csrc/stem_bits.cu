// Stem of the backbone straight from the 1-bit packed event histogram (replaces x.float() +
// replicate padding + Conv2d(k=7, s=4, no bias) + NCHW->NHWC + LayerNorm, ops.py:54-91 / sast_rnn.py:153) for the 16-bit mode:
//   packed uint8 [B, 20, H, W/8] (bit k of byte j = column 8 j + k)  ->  LayerNorm(conv(x)) fp32 NHWC [B, H/4, W/4, 64].
//
// The whole input is 4.9 MB at 1 Mpx B = 8 -- nothing is unpacked to memory.  Per output tile of 8 oy x 16 ox pixels
// (M = 128) the 35 input rows x 20 bins x 72 columns the tile touches are 8.4 KB of bits: a stager warp keeps the next
// tile's bits in shared memory (shifted so that the 8-column window of output pixel ox_i starts at bit 4 ox_i, replicate
// padding applied; double buffer).  The A operand never touches shared memory: with N = 64 an SS-form tcgen05.mma re-reads 4 KB of A per
// 2 KB of B and is bound by shared-memory bandwidth (measured: ~100 clk per M128 N64 K16 instruction against the 32 clk
// floor), so the operand lives in TENSOR MEMORY.  A producer thread owns one tile row (= TMEM lane): per kernel row ky and
// bin c it reads the 8-bit window of its pixel (two LDS.32 + a funnel shift), expands it to 8 fp16 values (kx = 0..6 and
// a zero-weight 8th tap) with a handful of integer instructions, and writes the 20 bins of a ky (80 TMEM columns) with
// tcgen05.st; the MMA warp issues TS-form instructions (A from TMEM, B = weights from shared memory).
// K is ordered (ky, c, kx8): one A buffer = one ky = 10 K steps of 2 bins.
//
// The fp16 weights (one rounding of 2^-12 relative per weight -- finer than the TF32 operand rounding of the cuDNN
// convolution this replaces; event bits are exact) stay RESIDENT in shared memory: 7 x 20 KB.
//
//   warps 0-7   producers: two per TMEM lane quarter (10 bins each); 4 A buffers of 80 TMEM columns
//   warp 8      TMEM allocator + MMA issuer (two 64-column accumulators: the epilogue of tile i overlaps tile i+1);
//               also TMA-loads the weights once
//   warp 9      stager: global -> shared memory copy of the NEXT tile's bits (double buffer)
//   warps 10-13 epilogue: LayerNorm over the channels of each pixel straight from TMEM, swizzled shared-memory transpose,
//               dense 256-byte row stores
#include "common.cuh"
#include "ptx.cuh"
#include "fused_common.cuh"

namespace sast {
extern long long* g_trace;
extern int g_trace_which;
namespace sb {

constexpr int kCout = 64, kCin = 20;
constexpr int kKRow = 160;                          // halves per (output channel, ky): 20 bins x 8 taps
constexpr int kABufs = 4;                           // A operand buffers in tensor memory
constexpr int kACols = 80;                          // one buffer: the 20 bins of one ky = 160 halves = 80 columns
constexpr int kChains = 1;                          // accumulation chains per tile (summed by the epilogue).  2 was measured:
                                                    // no effect -- back-to-back tcgen05.mma on ONE accumulator already run at the
                                                    // instruction floor (tools/mma_bench.cu), the issuer just waited for operands
constexpr int kAcc0 = 0, kA0 = 2 * kChains * kCout; // TMEM columns: two sets of accumulators, then the A buffers
static_assert(kA0 + kABufs * kACols <= 512, "tensor memory");
constexpr int kWRow = 2 * 8192 + 4096;              // weights of one ky: two [64 x 64] SW128 tiles + one [64 x 32] SW64 tile
constexpr int kInRows = 35, kInWords = 3;           // staged bits: 35 input rows x 20 bins x 3 words (72 columns used)
constexpr int kInBytes = (kInRows * kCin * kInWords * 4 + 1023) / 1024 * 1024;
constexpr int kItems = kInRows * kCin;              // (row, bin) staging items of a tile
constexpr int kProducers = 256;
constexpr int kThreads = 14 * 32;

struct Ctl {
  uint64_t full[kABufs];
  uint64_t empty[kABufs];
  uint64_t wbar;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1)
stem_bits_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w64,
                 const uint32_t* __restrict__ packed, int B, int H, int W, const float* __restrict__ ln_w,
                 const float* __restrict__ ln_b, float eps, float* __restrict__ out, long long* trace) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(16) float stage_smem[4][32 * 32];
  [[maybe_unused]] long long* const trc = trace ? trace + (size_t)blockIdx.x * 32 : nullptr;     // trace build: stamps of the 4th tile
  SAST_STAMP(trc, threadIdx.x == 0, 20);
  uint8_t* const wsm = smem_raw;                                       // 7 x kWRow
  uint32_t* const inb = reinterpret_cast<uint32_t*>(wsm + 7 * kWRow);  // 2 x kInBytes
  Ctl* const ctl = reinterpret_cast<Ctl*>(reinterpret_cast<uint8_t*>(inb) + 2 * kInBytes);
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int Ho = H / 4, Wo = W / 4, words = W / 32;
  const int tiles_x = (Wo + 15) / 16, tiles_y = (Ho + 7) / 8;
  const int total_tiles = B * tiles_y * tiles_x;

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&map_w); ptx::tma_prefetch_desc(&map_w64);
    for (int s = 0; s < kABufs; ++s) { ptx::mbar_init(&ctl->full[s], kProducers / 32); ptx::mbar_init(&ctl->empty[s], 1); }
    ptx::mbar_init(&ctl->wbar, 1);
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&ctl->tmem_full[a], 1); ptx::mbar_init(&ctl->tmem_empty[a], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 8) ptx::tmem_alloc(&ctl->tmem_base, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 8 && ptx::elect_one()) {             // the weights are parameters, not a predecessor's output: before the PDL wait
    ptx::mbar_arrive_expect_tx(&ctl->wbar, 7 * kWRow);
    for (int ky = 0; ky < 7; ++ky) {
      ptx::tma_load_2d(wsm + ky * kWRow, &map_w, &ctl->wbar, 0, ky * kCout);
      ptx::tma_load_2d(wsm + ky * kWRow + 8192, &map_w, &ctl->wbar, 64, ky * kCout);
      ptx::tma_load_2d(wsm + ky * kWRow + 16384, &map_w64, &ctl->wbar, 128, ky * kCout);
    }
  }
  __syncwarp();
  pdl_entry();
  SAST_STAMP(trc, threadIdx.x == 0, 21);

  if (warp < 8) {
    // ---------------- producers ----------------
    [[maybe_unused]] const int t = threadIdx.x;
    const int m = (warp & 3) * 32 + lane;                             // tile row = TMEM lane (a warp reaches its own lane quarter)
    const int h = warp >> 2;                                          // which 10 of the 20 bins
    const int oxi = m & 15;
    const uint32_t src0 = (uint32_t)((4 * (m >> 4) * kCin + 10 * h) * kInWords + (oxi >> 3));     // + ky * kCin * kInWords
    const uint32_t sh = (uint32_t)((oxi & 7) * 4);
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t it = 0, ti = 0;
    asm volatile("bar.sync 1, %0;" ::"n"(kProducers + 32) : "memory");                // tile 0 staged (warp 9)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t* in = inb + (ti & 1) * (kInBytes / 4);
      SAST_STAMP(trc, t == 0 && ti == 3, 0);
      for (int ky = 0; ky < 7; ++ky, ++it) {
        const uint32_t ab = it % kABufs, round = it / kABufs;
        const uint32_t* wp = in + src0 + (uint32_t)(ky * kCin * kInWords);
        uint32_t v[40];                                                // 10 bins x 8 halves
#pragma unroll
        for (int c = 0; c < 10; ++c) {
          const uint32_t bits = __funnelshift_r(wp[c * kInWords], wp[c * kInWords + 1], sh);
          // 4 bits -> 4 bytes of 0 / 1 (multiply-spread), x 0x3C -> the high bytes of fp16 1.0 / 0.0, two per word
          const uint32_t lo = (((bits & 15u) * 0x00204081u) & 0x01010101u) * 0x3Cu;
          const uint32_t hi = ((((bits >> 4) & 15u) * 0x00204081u) & 0x01010101u) * 0x3Cu;
          v[4 * c] = __byte_perm(lo, 0u, 0x1404);
          v[4 * c + 1] = __byte_perm(lo, 0u, 0x3424);
          v[4 * c + 2] = __byte_perm(hi, 0u, 0x1404);
          v[4 * c + 3] = __byte_perm(hi, 0u, 0x3424);
        }
        ptx::mbar_wait(&ctl->empty[ab], (round & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t ta = tmem_base + lane_sel + (uint32_t)(kA0 + ab * kACols + 40 * h);
        ptx::tmem_st_32x32(ta, v);
        ptx::tmem_st_32x8(ta + 32u, v + 32);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ctl->full[ab]);
      }
      SAST_STAMP(trc, t == 0 && ti == 3, 5);
      asm volatile("bar.sync 1, %0;" ::"n"(kProducers + 32) : "memory");              // next tile staged, this tile's bits dead
      SAST_STAMP(trc, t == 0 && ti == 3, 6);
    }
  } else if (warp == 9) {
    // ---------------- stager: the bits of the NEXT tile, global -> shared memory, while the producers expand this one ----------------
    // (row, bin) item = the four global words around the tile's 72 columns -> three staged words: word j holds columns
    // 64 tx - 3 + 32 j .., so the window of output pixel ox_i starts at bit 4 ox_i.  All 88 loads of a lane are issued
    // back to back (one latency per tile); a separate warp so that the producers' scoreboards never wait on global loads.
    constexpr int NB = (kItems + 31) / 32;                             // items per lane (22)
    auto stage_tile = [&](int tile, uint32_t* dst) {
      const int tx = tile % tiles_x, rest = tile / tiles_x;
      const int ty = rest % tiles_y, b = rest / tiles_y;
      uint32_t g[NB][4];
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int item = min(lane + 32 * u, kItems - 1);
        const int yy = item / kCin, c = item - yy * kCin;
        const int y = min(max(32 * ty - 3 + yy, 0), H - 1);            // replicate padding (rows)
        const uint32_t* row = packed + (((size_t)b * kCin + c) * H + y) * words;
#pragma unroll
        for (int k = 0; k < 4; ++k) g[u][k] = __ldg(row + min(max(2 * tx - 1 + k, 0), words - 1));
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int item = lane + 32 * u;
        if (item < kItems) {
          if (tx == 0) g[u][0] = (g[u][1] & 1u) ? 0xFFFFFFFFu : 0u;    // replicate padding (left edge): column 0 repeated
#pragma unroll
          for (int j = 0; j < 3; ++j) dst[item * kInWords + j] = __funnelshift_r(g[u][j], g[u][j + 1], 29);
        }
      }
    };
    uint32_t ti = 0;
    if ((int)blockIdx.x < total_tiles) stage_tile(blockIdx.x, inb);
    asm volatile("bar.sync 1, %0;" ::"n"(kProducers + 32) : "memory");
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int next = tile + gridDim.x;
      SAST_STAMP(trc, lane == 0 && ti == 3, 16);
      if (next < total_tiles) stage_tile(next, inb + ((ti + 1) & 1) * (kInBytes / 4));
      SAST_STAMP(trc, lane == 0 && ti == 3, 17);
      asm volatile("bar.sync 1, %0;" ::"n"(kProducers + 32) : "memory");
    }
  } else if (warp == 8) {
    // ---------------- MMA issuer ----------------
    const bool leader = ptx::elect_one();
    const uint32_t idesc = (1u << 4) | (((uint32_t)kCout >> 3) << 17) | ((128u >> 4) << 24);     // kind::f16: A = B = F16, D = F32, K-major
    ptx::mbar_wait(&ctl->wbar, 0);
    const uint32_t w0 = ptx::smem_u32(wsm);
    uint32_t it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti & 1, use = ti >> 1;
      SAST_STAMP(trc, lane == 0 && ti == 3, 8);
      ptx::mbar_wait(&ctl->tmem_empty[acc], (use & 1) ^ 1);
      ptx::tc_fence_after();
      SAST_STAMP(trc, lane == 0 && ti == 3, 9);
      const uint32_t tmem_d = tmem_base + (uint32_t)kAcc0 + acc * (uint32_t)(kChains * kCout);
      for (int ky = 0; ky < 7; ++ky, ++it) {
        const uint32_t ab = it % kABufs, round = it / kABufs;
        SAST_STAMP(trc, lane == 0 && ti == 3 && ky < 4, 24 + 2 * ky);
        ptx::mbar_wait(&ctl->full[ab], round & 1);
        ptx::tc_fence_after();
        SAST_STAMP(trc, lane == 0 && ti == 3 && ky < 4, 25 + 2 * ky);
        // whole warp on warp-uniform operands, only the instruction is guarded (no R2UR waterfall per UTCHMMA)
        const uint32_t ta = tmem_base + (uint32_t)(kA0 + ab * kACols);
        const uint32_t wt = w0 + (uint32_t)ky * kWRow;
        const uint64_t d0 = ptx::umma_desc_sw128_kmajor(wt), d1 = ptx::umma_desc_sw128_kmajor(wt + 8192), d2 = fl::desc_sw64_k(wt + 16384);
#pragma unroll
        for (int k = 0; k < 10; ++k) {                                 // K steps of 16 = 2 bins = 8 TMEM columns
          const uint64_t dw = (k < 4 ? d0 : k < 8 ? d1 : d2) + (uint64_t)((k & 3) * 2);
          if (leader)
            ptx::umma_f16_ts(tmem_d + (uint32_t)((k % kChains) * kCout), ta + (uint32_t)(8 * k), dw, idesc, (ky | (k / kChains)) ? 1u : 0u);
        }
        if (leader) ptx::umma_commit(&ctl->empty[ab]);
        __syncwarp();
      }
      if (leader) ptx::umma_commit(&ctl->tmem_full[acc]);
      __syncwarp();
      SAST_STAMP(trc, lane == 0 && ti == 3, 10);
    }
  } else if (warp >= 10) {
    // ---------------- epilogue: LayerNorm over channels (thread = pixel row), transposed store ----------------
    const int quarter = warp & 3;
    float* stage = &stage_smem[quarter][0];
    const int r_sub = lane >> 3, gq = lane & 7, c4 = gq * 4;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int tx = tile % tiles_x, rest = tile / tiles_x;
      const int ty = rest % tiles_y, b = rest / tiles_y;
      const uint32_t acc = ti & 1, use = ti >> 1;
      const uint32_t tmem_d = tmem_base + (uint32_t)kAcc0 + acc * (uint32_t)(kChains * kCout) + ((uint32_t)(quarter * 32) << 16);
      SAST_STAMP(trc, threadIdx.x == 320 && ti == 3, 12);
      ptx::mbar_wait(&ctl->tmem_full[acc], use & 1);
      ptx::tc_fence_after();
      SAST_STAMP(trc, threadIdx.x == 320 && ti == 3, 13);
      uint32_t raw0[32], raw1[32];
      ptx::tmem_ld_32x32(tmem_d, raw0);
      ptx::tmem_ld_32x32(tmem_d + 32u, raw1);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int ch = 1; ch < kChains; ++ch) {                           // sum of the accumulation chains
        uint32_t part[32];
        ptx::tmem_ld_32x32(tmem_d + (uint32_t)(ch * kCout), part);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) raw0[j] = __float_as_uint(__uint_as_float(raw0[j]) + __uint_as_float(part[j]));
        ptx::tmem_ld_32x32(tmem_d + (uint32_t)(ch * kCout) + 32u, part);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) raw1[j] = __float_as_uint(__uint_as_float(raw1[j]) + __uint_as_float(part[j]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->tmem_empty[acc]);          // the whole row sits in registers: the accumulator is free
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += __uint_as_float(raw0[j]) + __uint_as_float(raw1[j]);
      const float mean = sum * (1.0f / kCout);
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d0 = __uint_as_float(raw0[j]) - mean, d1 = __uint_as_float(raw1[j]) - mean;
        ss += d0 * d0 + d1 * d1;
      }
      const float rstd = rsqrtf(ss * (1.0f / kCout) + eps);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t* raw = half ? raw1 : raw0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 v;
          v.x = (__uint_as_float(raw[4 * k]) - mean) * rstd; v.y = (__uint_as_float(raw[4 * k + 1]) - mean) * rstd;
          v.z = (__uint_as_float(raw[4 * k + 2]) - mean) * rstd; v.w = (__uint_as_float(raw[4 * k + 3]) - mean) * rstd;
          *reinterpret_cast<float4*>(stage + lane * 32 + ((k ^ (lane & 7)) << 2)) = v;
        }
        __syncwarp();
        const int n = half * 32 + c4;
        float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ln_w) g4 = __ldg(reinterpret_cast<const float4*>(ln_w + n));
        if (ln_b) b4 = __ldg(reinterpret_cast<const float4*>(ln_b + n));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + r_sub;                                  // row of this warp's 32: tile row quarter*32 + r
          const int m = quarter * 32 + r;
          const int oy = 8 * ty + (m >> 4), ox = 16 * tx + (m & 15);
          if (oy >= Ho || ox >= Wo) continue;
          const float4 a4 = *reinterpret_cast<const float4*>(stage + r * 32 + ((gq ^ (r & 7)) << 2));
          *reinterpret_cast<float4*>(out + (((size_t)b * Ho + oy) * Wo + ox) * kCout + n) =
              make_float4(a4.x * g4.x + b4.x, a4.y * g4.y + b4.y, a4.z * g4.z + b4.z, a4.w * g4.w + b4.w);
        }
        __syncwarp();
      }
      SAST_STAMP(trc, threadIdx.x == 320 && ti == 3, 14);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  SAST_STAMP(trc, threadIdx.x == 0, 22);
  if (warp == 8) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sb

int make_tmap_bf16_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows,
                       int swizzle_bytes);

}  // namespace sast

// true if sast_stem_bits_fwd takes this geometry (the reference's stems: 20 event bins, embed_dim 64, patch_size 4)
extern "C" int sast_stem_bits_supported(int32_t bits, int32_t Cin, int32_t H, int32_t W, int32_t Cout) {
  return bits == 1 && Cin == sast::sb::kCin && Cout == sast::sb::kCout && H >= 32 && W >= 64 && H % 4 == 0 && W % 32 == 0;
}

// packed: 1-bit packed histogram uint8 [B, Cin, H, W/8] (the format of sast_unpack_nonzero_ratio) -> out [B,H/4,W/4,Cout] fp32
// NHWC = LayerNorm(conv7x7 stride 4, replicate padding 3, no bias).  w16: fp16 [7 * Cout, 160]: row ky * Cout + n holds
// conv.weight[n, c, ky, kx] at column c * 8 + kx (column c * 8 + 7 is zero).
extern "C" int sast_stem_bits_fwd(const uint8_t* packed, int32_t bits, int32_t B, int32_t Cin, int32_t H, int32_t W,
                                  const uint16_t* w16, int32_t Cout, const float* ln_w, const float* ln_b, float eps, float* out,
                                  void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(packed); SAST_CHECK_PTR(w16); SAST_CHECK_PTR(out);
  if (B <= 0) return SAST_E_SHAPE;
  if (!sast_stem_bits_supported(bits, Cin, H, W, Cout)) return SAST_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(packed) & 3) || (reinterpret_cast<uintptr_t>(w16) & 15)) return SAST_E_SHAPE;
  CUtensorMap mw, mw64;
  int rc;       // fp16 and bf16 are both 2-byte types without arithmetic in the copy engine: the bf16 map builder serves
  if ((rc = make_tmap_bf16_box(&mw, w16, 7ll * Cout, sb::kKRow, sb::kKRow, 64, Cout, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&mw64, w16, 7ll * Cout, sb::kKRow, sb::kKRow, 32, Cout, 64))) return rc;
  const size_t smem = (size_t)7 * sb::kWRow + 2 * sb::kInBytes + sizeof(sb::Ctl) + 64;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(sb::stem_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int Ho = H / 4, Wo = W / 4;
  const long long tiles = (long long)B * ((Ho + 7) / 8) * ((Wo + 15) / 16);
  if (tiles >= (1ll << 31)) return SAST_E_UNSUPPORTED;
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  launch_k(sb::stem_bits_kernel, dim3(grid), dim3(sb::kThreads), smem, (cudaStream_t)stream, mw, mw64, (const uint32_t*)packed, B, H, W,
           ln_w, ln_b, eps, out, g_trace_which == 6 ? g_trace : nullptr);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}
