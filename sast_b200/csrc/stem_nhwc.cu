// Stem of the backbone, TMA-fed (replaces x.float() + replicate padding + Conv2d(k=7, s=4,
// no bias) + NCHW->NHWC + LayerNorm, ops.py:54-91 / sast_rnn.py:153) for the 16-bit mode:
//   xh  fp16 [B, H+8, W+8, Cin] NHWC with the padding materialised (sast_events_nhwc)  ->  LayerNorm(conv(x)) fp32 NHWC.
//
// In that layout the 7 x Cin window of one output pixel and one kernel row ky is CONTIGUOUS (140 halves at Cin = 20) and the
// windows of neighbouring output pixels start 4 Cin halves apart, so the im2col A operand is a plain (overlapping-stride)
// 5-D tensor map  {k: 144, ox: Wo (stride 4 Cin), ky: 4 (stride row), oy: Ho+1 (stride 4 rows), b}  and one TMA box
// {64 | 16, 16 ox, 1, 9 oy, 1} lands as a ready SWIZZLE_128B (SWIZZLE_32B for the 16-wide K tail) operand tile: no
// producer warps, no conversions.  The 9-row box serves TWO kernel rows: rows 0..127 of the tile are the operand of ky,
// rows 16..143 (one output row further down = 4 input rows) the operand of ky + 4, so every input row is fetched once.
// An output tile is 8 oy x 16 ox pixels (M = 128).  K per kernel row = 144 = 64 + 64 + 16 (4 zero-weight columns).
//
// The weights are fp16 (one rounding of 2^-12 relative per weight -- finer than the TF32 operand rounding of the cuDNN
// convolution this replaces, torch.backends.cudnn.allow_tf32 = True) and stay RESIDENT in shared memory: 7 x 18 KB.
// Event counts are exact in fp16.  The fp32-grade path (split weights) is stem_tc.cu.
//
//   warp 0      TMA producer (weights once, then 12 operand boxes per tile through a 4-stage ring)
//   warp 1      TMEM allocator + MMA issuer (two 64-column accumulators: the epilogue of tile i overlaps tile i+1)
//   warps 2-5   epilogue: LayerNorm over the channels of each pixel straight from TMEM, swizzled shared-memory transpose,
//               dense 256-byte row stores
#include "common.cuh"
#include "ptx.cuh"

namespace sast {
namespace sn {

constexpr int kCout = 64, kCin = 20;
constexpr int kKRow = 144;                          // halves per (pixel, ky): 7 * 20 = 140 padded to 144
constexpr int kStages = 4;
constexpr int kABytes = 144 * 128;                  // one operand box: 9 oy x 16 ox rows of 128 bytes
constexpr int kWRow = 2 * 8192 + 2048;              // weights of one ky: two [64 x 64] SW128 tiles + one [64 x 16] SW32 tile
constexpr int kThreads = 6 * 32;

struct Ctl {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t wbar;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
// K-major operand tile in the canonical SWIZZLE_32B layout: rows of 32 bytes (16 halves = one K step), 8-row groups 256 B apart
__device__ __forceinline__ uint64_t desc_sw32_k(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)6 << 61);
}

__global__ void __launch_bounds__(kThreads, 1)
stem_nhwc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a16,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w16, int B, int Ho, int Wo,
                 const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(16) float stage_smem[4][32 * 32];
  uint8_t* const wsm = smem_raw;                                       // 7 x kWRow
  uint8_t* const ring = smem_raw + 7 * kWRow;                          // kStages x kABytes
  Ctl* const ctl = reinterpret_cast<Ctl*>(ring + kStages * kABytes);
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int tiles_x = (Wo + 15) / 16, tiles_y = (Ho + 7) / 8;
  const int total_tiles = B * tiles_y * tiles_x;

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&map_a); ptx::tma_prefetch_desc(&map_a16);
    ptx::tma_prefetch_desc(&map_w); ptx::tma_prefetch_desc(&map_w16);
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&ctl->full[s], 1); ptx::mbar_init(&ctl->empty[s], 1); }
    ptx::mbar_init(&ctl->wbar, 1);
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&ctl->tmem_full[a], 1); ptx::mbar_init(&ctl->tmem_empty[a], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(&ctl->tmem_base, 128);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0 && ptx::elect_one()) {             // the weights are parameters, not a predecessor's output: before the PDL wait
    ptx::mbar_arrive_expect_tx(&ctl->wbar, 7 * kWRow);
    for (int ky = 0; ky < 7; ++ky) {
      ptx::tma_load_2d(wsm + ky * kWRow, &map_w, &ctl->wbar, 0, ky * kCout);
      ptx::tma_load_2d(wsm + ky * kWRow + 8192, &map_w, &ctl->wbar, 64, ky * kCout);
      ptx::tma_load_2d(wsm + ky * kWRow + 16384, &map_w16, &ctl->wbar, 128, ky * kCout);
    }
  }
  __syncwarp();
  pdl_entry();

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    const bool leader = ptx::elect_one();
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tx = tile % tiles_x, rest = tile / tiles_x;
      const int ty = rest % tiles_y, b = rest / tiles_y;
      for (int ky = 0; ky < 4; ++ky) {
        for (int part = 0; part < 3; ++part, ++it) {
          const uint32_t s = it % kStages, round = it / kStages;
          ptx::mbar_wait(&ctl->empty[s], (round & 1) ^ 1);
          if (leader) {
            ptx::mbar_arrive_expect_tx(&ctl->full[s], part < 2 ? 144 * 128 : 144 * 32);
            tma_load_5d(ring + s * kABytes, part < 2 ? &map_a : &map_a16, &ctl->full[s], 64 * part, 16 * tx, ky, 8 * ty, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const bool leader = ptx::elect_one();
    const uint32_t idesc = (1u << 4) | (((uint32_t)kCout >> 3) << 17) | ((128u >> 4) << 24);     // kind::f16: A = B = F16, D = F32, K-major
    ptx::mbar_wait(&ctl->wbar, 0);
    const uint32_t w0 = ptx::smem_u32(wsm), r0 = ptx::smem_u32(ring);
    uint32_t it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti & 1, use = ti >> 1;
      ptx::mbar_wait(&ctl->tmem_empty[acc], (use & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)kCout;
      for (int ky4 = 0; ky4 < 4; ++ky4) {
        for (int part = 0; part < 3; ++part, ++it) {
          const uint32_t s = it % kStages, round = it / kStages;
          ptx::mbar_wait(&ctl->full[s], round & 1);
          ptx::tc_fence_after();
          const uint32_t sa = r0 + s * kABytes;
          // whole warp on warp-uniform descriptors, only the instruction is guarded (no R2UR waterfall per UTCHMMA)
          for (int g = 0; g < (ky4 < 3 ? 2 : 1); ++g) {             // kernel rows ky4 and ky4 + 4 share the operand box
            const uint32_t wt = w0 + (uint32_t)(ky4 + 4 * g) * kWRow + (uint32_t)part * 8192;
            const uint64_t da = part < 2 ? ptx::umma_desc_sw128_kmajor(sa + (uint32_t)g * 16 * 128) : desc_sw32_k(sa + (uint32_t)g * 16 * 32);
            const uint64_t dw = part < 2 ? ptx::umma_desc_sw128_kmajor(wt) : desc_sw32_k(wt);
            const int steps = part < 2 ? 4 : 1;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < steps && leader)
                ptx::umma_f16_ss(tmem_d, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (ky4 | part | g | k) ? 1u : 0u);
          }
          if (leader) ptx::umma_commit(&ctl->empty[s]);
          __syncwarp();
        }
      }
      if (leader) ptx::umma_commit(&ctl->tmem_full[acc]);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: LayerNorm over channels (thread = pixel row), transposed store ----------------
    const int quarter = warp & 3;
    float* stage = &stage_smem[quarter][0];
    const int r_sub = lane >> 3, gq = lane & 7, c4 = gq * 4;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int tx = tile % tiles_x, rest = tile / tiles_x;
      const int ty = rest % tiles_y, b = rest / tiles_y;
      const uint32_t acc = ti & 1, use = ti >> 1;
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)kCout + ((uint32_t)(quarter * 32) << 16);
      ptx::mbar_wait(&ctl->tmem_full[acc], use & 1);
      ptx::tc_fence_after();
      uint32_t raw0[32], raw1[32];
      ptx::tmem_ld_32x32(tmem_d, raw0);
      ptx::tmem_ld_32x32(tmem_d + 32u, raw1);
      ptx::tmem_ld_wait();
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
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace sn

typedef CUresult (*EncodeTiledFnS)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnS stem_encode_fn() {
  static EncodeTiledFnS fn = nullptr;     // idempotent lookup; benign race
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFnS)p;
  }
  return fn;
}

int make_tmap_bf16_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows,
                       int swizzle_bytes);

}  // namespace sast

// true if sast_stem_nhwc_fwd takes this geometry (the reference's stems: 20 event bins, embed_dim 64, patch_size 4)
extern "C" int sast_stem_nhwc_supported(int32_t Cin, int32_t H, int32_t W, int32_t Cout) {
  // (the last term: sast_events_nhwc stages the source rows and one padded output row of the widest format in 48 KB)
  return Cin == sast::sn::kCin && Cout == sast::sn::kCout && H >= 32 && W >= 32 && H % 4 == 0 && W % 32 == 0 &&
         (size_t)Cin * W + (size_t)(W + 8) * Cin * 2 <= 48 * 1024;
}

// xh fp16 [B, H+8, W+8, Cin] (sast_events_nhwc) -> out [B,H/4,W/4,Cout] fp32 NHWC = LayerNorm(conv7x7 stride 4, replicate
// padding 3, no bias).  w16: fp16 [7 * Cout, 144]: row ky * Cout + n holds conv.weight[n, :, ky, :] ordered (kx, c), then 4 zeros.
extern "C" int sast_stem_nhwc_fwd(const uint16_t* xh, int32_t B, int32_t Cin, int32_t H, int32_t W, const uint16_t* w16,
                                  int32_t Cout, const float* ln_w, const float* ln_b, float eps, float* out, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(xh); SAST_CHECK_PTR(w16); SAST_CHECK_PTR(out);
  if (B <= 0) return SAST_E_SHAPE;
  if (!sast_stem_nhwc_supported(Cin, H, W, Cout)) return SAST_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(xh) & 15) || (reinterpret_cast<uintptr_t>(w16) & 15)) return SAST_E_SHAPE;
  EncodeTiledFnS enc = stem_encode_fn();
  if (!enc) return (int)cudaErrorNotSupported;
  const int Ho = H / 4, Wo = W / 4, Hp = H + 8, Wp = W + 8;
  const cuuint64_t pitch = (cuuint64_t)Wp * Cin * 2;                  // bytes per padded input row
  // {k, ox, ky, oy, b}: overlapping windows (stride 4 Cin halves < 144), every stride a multiple of the one before
  const cuuint64_t dims[5] = {(cuuint64_t)sn::kKRow, (cuuint64_t)Wo, 4, (cuuint64_t)Ho + 1, (cuuint64_t)B};
  const cuuint64_t strides[4] = {(cuuint64_t)4 * Cin * 2, pitch, 4 * pitch, (cuuint64_t)Hp * pitch};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap ma, ma16, mw, mw16;
  {
    const cuuint32_t box[5] = {64, 16, 1, 9, 1};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<uint16_t*>(xh), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
    const cuuint32_t box16[5] = {16, 16, 1, 9, 1};
    if (enc(&ma16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<uint16_t*>(xh), dims, strides, box16, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }
  int rc;       // fp16 and bf16 are both 2-byte types without arithmetic in the copy engine: the bf16 map builder serves
  if ((rc = make_tmap_bf16_box(&mw, w16, 7ll * Cout, sn::kKRow, sn::kKRow, 64, Cout, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&mw16, w16, 7ll * Cout, sn::kKRow, sn::kKRow, 16, Cout, 32))) return rc;
  const size_t smem = (size_t)7 * sn::kWRow + (size_t)sn::kStages * sn::kABytes + sizeof(sn::Ctl) + 64;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(sn::stem_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (long long)B * ((Ho + 7) / 8) * ((Wo + 15) / 16);
  if (tiles >= (1ll << 31)) return SAST_E_UNSUPPORTED;
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  launch_k(sn::stem_nhwc_kernel, dim3(grid), dim3(sn::kThreads), smem, (cudaStream_t)stream, ma, ma16, mw, mw16, B, Ho, Wo, ln_w, ln_b,
           eps, out);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}
