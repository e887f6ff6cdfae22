// Downsample stem of stages 2-3 as ONE kernel (replaces Conv2d(k=3, s=2, replicate padding 1, no bias) +
// NCHW->NHWC + LayerNorm of ConvDownsampling_Cf2Cl, ops.py:54-95) for the 16-bit mode:
//   xp bf16 [B, H+2, W+2, Cin] NHWC, replicate-padded (sast_pad_nhwc_bf16)  ->  LayerNorm(conv(x)) fp32 NHWC [B, H/2, W/2, Cout].
//
// Implicit GEMM on tcgen05 (kind::f16, bf16 operands like every other GEMM of the 16-bit mode): in NHWC the 3 x Cin window of
// one output pixel and one kernel row ky is CONTIGUOUS and the windows of neighbouring output pixels start 2 Cin elements
// apart, so im2col is a plain overlapping-stride 5-D tensor map
//   {k: 3 Cin, ox: Wo (stride 2 Cin), ky: 3 (stride row), oy: Ho (stride 2 rows), b}
// and one TMA box {64 elements, ox_n, 1, oy_n, 1} (ox_n x oy_n = 128 output pixels) lands as a ready SWIZZLE_128B operand tile.
// K is ordered (ky, kx, c): 9 Cin / 64 k-blocks.  Cin = 64: the 147 KB of bf16 weights stay RESIDENT in shared memory and the
// ring carries operand boxes only; Cin = 128 (590 KB): the weight k-block [Cout x 64] streams through the ring next to the
// operand box (L2 resident).  A first version on kind::tf32 straight from the fp32 map moved 4 bytes per element twice
// (2.25x im2col duplication + the weights per tile) and was bound by the L2 rate: 18.5 / 22.6 us against 8 / 11 us here.
// All Cout channels of a pixel sit in one accumulator, so the LayerNorm runs in the epilogue straight from tensor memory.
//
//   warp 0      TMA producer (resident weights once; operand box (+ weight box) per k-block, 3- / 4-stage ring)
//   warp 1      TMEM allocator + MMA issuer (two Cout-column accumulators: the epilogue of tile i overlaps tile i+1)
//   warps 2-5   epilogue: LayerNorm over channels (thread = pixel; three passes over TMEM in 32-column chunks), swizzled
//               shared-memory transpose, dense 128-byte row stores
#include "common.cuh"
#include "ptx.cuh"

namespace sast {
namespace ds {

constexpr int kMaxStages = 4;
constexpr int kThreads = 6 * 32;

struct Ctl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
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

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1)
downsample_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int B, int Ho, int Wo,
                  int ox_shift, const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, float* __restrict__ out) {
  constexpr int kKbPerRow = 3 * CIN / 64;                               // k-blocks of 64 elements per kernel row
  constexpr int kKb = 3 * kKbPerRow;
  constexpr bool kResident = CIN == 64;                                 // all weight k-blocks stay in shared memory
  constexpr int kStages = kResident ? 3 : 4;
  constexpr uint32_t kABytes = 128 * 128, kWBytes = COUT * 128, kStageBytes = kResident ? kABytes : kABytes + kWBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(16) float stage_smem[4][32 * 32];
  uint8_t* const wres = smem_raw;                                       // resident weights: kKb tiles of [COUT x 64] (or nothing)
  uint8_t* const ring = smem_raw + (kResident ? kKb * kWBytes : 0);
  Ctl* const ctl = reinterpret_cast<Ctl*>(ring + kStages * kStageBytes);
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int ox_n = 1 << ox_shift, oy_n = 128 >> ox_shift;
  const int tiles_x = (Wo + ox_n - 1) >> ox_shift, tiles_y = (Ho + oy_n - 1) / oy_n;
  const int total_tiles = B * tiles_y * tiles_x;

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&map_a); ptx::tma_prefetch_desc(&map_w);
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&ctl->full[s], 1); ptx::mbar_init(&ctl->empty[s], 1); }
    ptx::mbar_init(&ctl->wbar, 1);
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&ctl->tmem_full[a], 1); ptx::mbar_init(&ctl->tmem_empty[a], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(&ctl->tmem_base, 2 * COUT);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  if (kResident && warp == 0 && ptx::elect_one()) {   // the weights are parameters, not a predecessor's output: before the PDL wait
    ptx::mbar_arrive_expect_tx(&ctl->wbar, kKb * kWBytes);
    for (int kb = 0; kb < kKb; ++kb) ptx::tma_load_2d(wres + kb * kWBytes, &map_w, &ctl->wbar, kb * 64, 0);
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
      for (int ky = 0; ky < 3; ++ky) {
        for (int kb = 0; kb < kKbPerRow; ++kb, ++it) {
          const uint32_t s = it % kStages, round = it / kStages;
          ptx::mbar_wait(&ctl->empty[s], (round & 1) ^ 1);
          if (leader) {
            uint8_t* st = ring + s * kStageBytes;
            ptx::mbar_arrive_expect_tx(&ctl->full[s], kStageBytes);
            tma_load_5d(st, &map_a, &ctl->full[s], 64 * kb, tx << ox_shift, ky, ty * oy_n, b);
            if (!kResident) ptx::tma_load_2d(st + kABytes, &map_w, &ctl->full[s], (ky * kKbPerRow + kb) * 64, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const bool leader = ptx::elect_one();
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, COUT);
    const uint32_t r0 = ptx::smem_u32(ring), w0 = ptx::smem_u32(wres);
    if (kResident) ptx::mbar_wait(&ctl->wbar, 0);
    uint32_t it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti & 1, use = ti >> 1;
      ptx::mbar_wait(&ctl->tmem_empty[acc], (use & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)COUT;
      for (int kb = 0; kb < kKb; ++kb, ++it) {
        const uint32_t s = it % kStages, round = it / kStages;
        ptx::mbar_wait(&ctl->full[s], round & 1);
        ptx::tc_fence_after();
        // whole warp on warp-uniform descriptors, only the instruction is guarded (no R2UR waterfall per UTCHMMA)
        const uint32_t sa = r0 + s * kStageBytes;
        const uint64_t da = ptx::umma_desc_sw128_kmajor(sa);
        const uint64_t dw = ptx::umma_desc_sw128_kmajor(kResident ? w0 + (uint32_t)kb * kWBytes : sa + kABytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)                                    // 16 bf16 = 32 bytes per K step
          if (leader) ptx::umma_f16_ss(tmem_d, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        if (leader) ptx::umma_commit(&ctl->empty[s]);
        __syncwarp();
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
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)COUT + ((uint32_t)(quarter * 32) << 16);
      ptx::mbar_wait(&ctl->tmem_full[acc], use & 1);
      ptx::tc_fence_after();
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tmem_d + (uint32_t)c0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += __uint_as_float(raw[j]);
      }
      const float mean = sum * (1.0f / COUT);
      float ss = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tmem_d + (uint32_t)c0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = __uint_as_float(raw[j]) - mean; ss += d * d; }
      }
      const float rstd = rsqrtf(ss * (1.0f / COUT) + eps);
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += 32) {
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
          const int r = i * 4 + r_sub;                                  // row of this warp's 32: tile row quarter*32 + r
          const int m = quarter * 32 + r;
          const int oy = ty * oy_n + (m >> ox_shift), ox = (tx << ox_shift) + (m & (ox_n - 1));
          if (oy >= Ho || ox >= Wo) continue;
          const float4 a4 = *reinterpret_cast<const float4*>(stage + r * 32 + ((gq ^ (r & 7)) << 2));
          *reinterpret_cast<float4*>(out + (((size_t)b * Ho + oy) * Wo + ox) * COUT + n) =
              make_float4(a4.x * g4.x + b4.x, a4.y * g4.y + b4.y, a4.z * g4.z + b4.z, a4.w * g4.w + b4.w);
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->tmem_empty[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 2 * COUT);
  }
}

typedef CUresult (*EncodeTiledFnD)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnD encode_fn() {
  static EncodeTiledFnD fn = nullptr;     // idempotent lookup; benign race
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFnD)p;
  }
  return fn;
}

template <int CIN, int COUT>
static int launch(const uint16_t* xp, int B, int H, int W, const uint16_t* w9, const float* ln_w, const float* ln_b, float eps, float* out,
                  cudaStream_t st) {
  EncodeTiledFnD enc = encode_fn();
  if (!enc) return (int)cudaErrorNotSupported;
  const int Ho = H / 2, Wo = W / 2, Hp = H + 2, Wp = W + 2;
  const int ox_shift = Wo % 16 == 0 ? 4 : 3;
  const cuuint64_t pitch = (cuuint64_t)Wp * CIN * 2;                  // bytes per padded input row
  // {k, ox, ky, oy, b}: overlapping windows (stride 2 Cin elements < 3 Cin), every stride a multiple of the one before
  const cuuint64_t dims[5] = {(cuuint64_t)3 * CIN, (cuuint64_t)Wo, 3, (cuuint64_t)Ho, (cuuint64_t)B};
  const cuuint64_t strides[4] = {(cuuint64_t)2 * CIN * 2, pitch, 2 * pitch, (cuuint64_t)Hp * pitch};
  const cuuint32_t box[5] = {64, (cuuint32_t)(1 << ox_shift), 1, (cuuint32_t)(128 >> ox_shift), 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap ma, mw;
  if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<uint16_t*>(xp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  const cuuint64_t wdims[2] = {(cuuint64_t)9 * CIN, (cuuint64_t)COUT};
  const cuuint64_t wstr[1] = {(cuuint64_t)9 * CIN * 2};
  const cuuint32_t wbox[2] = {64, (cuuint32_t)COUT};
  if (enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(w9), wdims, wstr, wbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  const size_t smem = (CIN == 64 ? (size_t)(9 * CIN / 64) * COUT * 128 + 3 * (128 * 128) : (size_t)4 * (128 * 128 + COUT * 128)) + sizeof(Ctl) + 64;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(downsample_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ox_n = 1 << ox_shift, oy_n = 128 >> ox_shift;
  const long long tiles = (long long)B * ((Ho + oy_n - 1) / oy_n) * ((Wo + ox_n - 1) / ox_n);
  if (tiles >= (1ll << 31)) return SAST_E_UNSUPPORTED;
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  launch_k(downsample_kernel<CIN, COUT>, dim3(grid), dim3(kThreads), smem, st, ma, mw, B, Ho, Wo, ox_shift, ln_w, ln_b, eps, out);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

}  // namespace ds
}  // namespace sast

// true if sast_downsample_fwd takes this geometry (the base model's stages 2 and 3: 64 -> 128 and 128 -> 256 channels)
extern "C" int sast_downsample_supported(int32_t Cin, int32_t H, int32_t W, int32_t Cout) {
  return ((Cin == 64 && Cout == 128) || (Cin == 128 && Cout == 256)) && H >= 2 && W >= 16 && H % 2 == 0 && W % 2 == 0;
}

// xp bf16 [B, H+2, W+2, Cin] (replicate-padded NHWC, sast_pad_nhwc_bf16) -> out [B,H/2,W/2,Cout] fp32 NHWC = LayerNorm(conv 3x3,
// stride 2, no bias).  w9: bf16 [Cout, 9*Cin], column ky*3*Cin + kx*Cin + c = conv.weight[n, c, ky, kx].
extern "C" int sast_downsample_fwd(const uint16_t* xp, int32_t B, int32_t Cin, int32_t H, int32_t W, const uint16_t* w9, int32_t Cout,
                                   const float* ln_w, const float* ln_b, float eps, float* out, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(xp); SAST_CHECK_PTR(w9); SAST_CHECK_PTR(out);
  if (B <= 0) return SAST_E_SHAPE;
  if (!sast_downsample_supported(Cin, H, W, Cout)) return SAST_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(xp) & 15) || (reinterpret_cast<uintptr_t>(w9) & 15)) return SAST_E_SHAPE;
  if (Cin == 64) return ds::launch<64, 128>(xp, B, H, W, w9, ln_w, ln_b, eps, out, (cudaStream_t)stream);
  return ds::launch<128, 256>(xp, B, H, W, w9, ln_w, ln_b, eps, out, (cudaStream_t)stream);
}
