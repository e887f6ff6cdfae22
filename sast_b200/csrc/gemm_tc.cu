// Token-wise GEMMs of the MS-WSA layer on the 5th-gen tensor cores (SAST_BF16 path):
//   D[M,N] = A[M,K] W[N,K]^T (+ bias) with the layer's fused epilogues (layer.cuh: Epi).
// A = compacted activations (bf16, row-major = K-major), W = nn.Linear weight [N,K] (K-major):
// the canonical TN shape, so both operands are TMA-loaded as SWIZZLE_128B tiles and fed to
// tcgen05.mma straight from shared memory; the fp32 accumulator lives in TMEM.
//
// Persistent CTAs (one per SM), each walking 128 x BN output tiles, 14 warps:
//   warp 0   TMA producer   (elected lane; ring of 4 {A 128x64, W BNx64} bf16 stages, runs ahead across tiles)
//   warp 1   TMEM allocator + MMA issuer (elected lane; tcgen05.mma M=128,N=BN,K=16, 4 per k-block; three
//            accumulators so the next tiles are multiplied while the previous ones are drained)
//   warps 2-13 epilogue (three groups of 4, one per accumulator): tcgen05.ld 32 lanes x 32 columns per warp
//            (warp w owns TMEM lanes 32*(w%4)..+31 = output rows); bias / LayerScale / residual / GLU / scatter
//            in registers.
// M (= number of selected tokens) is read from device memory; CTAs past it exit at once.
// K tails (K % 64 != 0) rely on TMA zero fill and issue only the k-steps that hold data.
#include "layer.cuh"
#include "ptx.cuh"

namespace sast {

constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 4;
constexpr int TC_MAX_GROUPS = 3;
// Epilogue groups of 4 warps, one TMEM accumulator each.  The epilogues dominate these kernels (a 128 x BN tile is
// multiplied in ~650 clk and drained in 5-10k clk, tools/gemm_trace.py): three groups for every flavour -- the
// math-heavy ones (GLU, LSTM gates) are issue bound, the others latency bound, and a third group helps both.
constexpr int tc_groups(int epi) { return 3 + 0 * epi; }
constexpr int tc_threads(int epi) { return 64 + tc_groups(epi) * 128; }   // warp 0 TMA, warp 1 MMA, then the epilogue warps

struct TcSmem {            // lives after the operand ring (which needs 1024-byte alignment)
  uint64_t full[TC_STAGES];
  uint64_t empty[TC_STAGES];
  uint64_t tmem_full[TC_MAX_GROUPS];
  uint64_t tmem_empty[TC_MAX_GROUPS];
  uint32_t tmem_base;
};

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float* v) {
  uint4 pk;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(v[0], v[1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[2], v[3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[4], v[5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(v[6], v[7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
  *reinterpret_cast<uint4*>(dst) = pk;
}

// Persistent: every CTA walks tiles t = blockIdx.x, +gridDim.x, ... (n-tile fastest, so CTAs that
// run side by side share the A tile in L2).  The TMA ring runs ahead across tile boundaries, the MMA
// warp cycles through three TMEM accumulators, and the three epilogue groups (4 warps each, one per
// accumulator) drain tiles i, i+1 while tile i+2 is being loaded and multiplied.
template <int EPI>
__global__ void __launch_bounds__(tc_threads(EPI), 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_a2, int k_split,
                                                                const __grid_constant__ CUtensorMap map_w,
                                                                const float* __restrict__ bias, int N, int K, int BN, int stages,
                                                                const int* __restrict__ counts, int m_static, EpiParams ep) {
  // trace build only: [CTA][128] clock64: 0 entry, 1 set-up done, 2 end; per tile ti < 8: MMA warp 8+4ti {accumulator free,
  // first k-block landed, last commit issued}, epilogue (quarter 0) 48+4ti {accumulator full, drained}, producer 100+ti; 127 SM id
  [[maybe_unused]] long long* const trc = ep.trace ? ep.trace + (size_t)blockIdx.x * 128 : nullptr;
  SAST_STAMP(trc, threadIdx.x == 0, 0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int NG = tc_groups(EPI);
  __shared__ __align__(16) float stage_smem[NG * 4][32 * 32];   // epilogue transpose tiles (XOR-swizzled 16-byte groups)
  const int M = counts ? counts[1] : m_static;
  const int m_tiles = (M + TC_BM - 1) / TC_BM, n_tiles = N / BN;
  const int total_tiles = m_tiles * n_tiles;
  if ((int)blockIdx.x >= total_tiles) return;

  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TC_BM * TC_BK * 2, w_bytes = (uint32_t)BN * TC_BK * 2;
  const uint32_t stage_bytes = a_bytes + w_bytes;
  TcSmem* sm = reinterpret_cast<TcSmem*>(base + (size_t)stages * stage_bytes);

  const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;    // warp: provably uniform
  // operand element: bf16 (64 per 128-byte k-block row) or, for EPI_LSTM, fp32 read as TF32 (32 per row)
  constexpr bool kTf32 = (EPI == EPI_LSTM);
  constexpr int kBK = kTf32 ? 32 : TC_BK;
  constexpr int kKStep = kTf32 ? 8 : 16;
  const int nkb = (K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&map_a);
    ptx::tma_prefetch_desc(&map_w);
    for (int s = 0; s < stages; ++s) { ptx::mbar_init(&sm->full[s], 1); ptx::mbar_init(&sm->empty[s], 1); }
    for (int a = 0; a < NG; ++a) { ptx::mbar_init(&sm->tmem_full[a], 1); ptx::mbar_init(&sm->tmem_empty[a], 4); }
    ptx::fence_barrier_init();
  }
  uint32_t tmem_cols = 32;                                          // NG accumulators of BN columns, power of two
  while (tmem_cols < (uint32_t)(NG * BN)) tmem_cols <<= 1;
  if (warp == 1) ptx::tmem_alloc(&sm->tmem_base, tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sm->tmem_base;
  // PDL: everything above (barriers, TMEM, index loads of data written >= 2 kernels ago) overlapped the tail of the
  // preceding kernel; its output is read only from here on
  pdl_entry();
  SAST_STAMP(trc, threadIdx.x == 0, 1);
#ifdef SAST_TRACE
  if (trc && threadIdx.x == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); trc[127] = smid; }
#endif

  // The producer and MMA warps run their loops as WHOLE warps on warp-uniform values and guard only the TMA /
  // tcgen05 instructions with an elected lane: under `if (lane == 0)` the descriptors sit in per-thread registers
  // and every UTCHMMA pays an ELECT / R2UR waterfall loop (~120 clk per instruction, measured in attn_tc).
  if (warp == 0) {
    const bool leader = ptx::elect_one();
    uint32_t it = 0, pti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++pti) {
      const int m0 = (t / n_tiles) * TC_BM, n0 = (t % n_tiles) * BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % (uint32_t)stages, round = it / (uint32_t)stages;
        ptx::mbar_wait(&sm->empty[s], (round & 1) ^ 1);
        SAST_STAMP(trc, leader && kb == 0 && pti < 8, 100 + pti);
        uint8_t* sa = base + (size_t)s * stage_bytes;
        const int k0 = kb * kBK;
        if (leader) {
          ptx::mbar_arrive_expect_tx(&sm->full[s], stage_bytes);
          if (k0 < k_split) ptx::tma_load_2d(sa, &map_a, &sm->full[s], k0, m0);
          else ptx::tma_load_2d(sa, &map_a2, &sm->full[s], k0 - k_split, m0);      // second A source ([x | h_prev])
          ptx::tma_load_2d(sa + a_bytes, &map_w, &sm->full[s], k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = ptx::elect_one();
    const uint32_t idesc = kTf32 ? ((1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)BN >> 3) << 17) | ((TC_BM >> 4) << 24))
                                 : ptx::umma_idesc_bf16(TC_BM, (uint32_t)BN);
    uint32_t it = 0, ti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      const uint32_t acc = ti % NG, use = ti / NG;
      ptx::mbar_wait(&sm->tmem_empty[acc], (use & 1) ^ 1);       // epilogue has drained this accumulator
      SAST_STAMP(trc, leader && ti < 8, 8 + 4 * ti);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % (uint32_t)stages, round = it / (uint32_t)stages;
        ptx::mbar_wait(&sm->full[s], round & 1);
        SAST_STAMP(trc, leader && kb == 0 && ti < 8, 9 + 4 * ti);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(base + (size_t)s * stage_bytes);
        const uint64_t da = ptx::umma_desc_sw128_kmajor(sa);
        const uint64_t db = ptx::umma_desc_sw128_kmajor(sa + a_bytes);
        const int krem = K - kb * kBK;
        const int ksteps = krem >= kBK ? 4 : (krem + kKStep - 1) / kKStep;
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < ksteps) {
              // advance one k-step = 32 bytes along K inside the 128-byte swizzle atom: +2 in the >>4 address field
              if (kTf32) ptx::umma_tf32_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
              else ptx::umma_f16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            }
          }
          ptx::umma_commit(&sm->empty[s]);          // frees the smem slot when these MMAs retire
        }
      }
      if (leader) ptx::umma_commit(&sm->tmem_full[acc]);      // accumulator complete
      SAST_STAMP(trc, leader && ti < 8, 10 + 4 * ti);
    }
  } else {
    // ---------------- epilogue: group g = (warp-2)/4 drains accumulator g; TMEM lane quarter = warp % 4 ----------------
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;
    // tcgen05.ld gives every lane 32 consecutive accumulator columns of ITS row; global memory wants the
    // opposite (a warp instruction covering whole row segments).  Phase A: each lane drops its 32 values
    // into a [32 rows x 128 B] shared-memory tile with 8 x STS.128, 16-byte groups XOR-swizzled by
    // (row % 8) -- conflict free.  Phase B: lane = (row-in-group r_sub, 16-byte column group gq); per
    // iteration a warp handles 4 full rows: LDS.128, epilogue math on 4 columns whose bias / gamma are
    // lane constants, one 4..16-byte store per output -- every global access is a dense row segment.
    float* stage = &stage_smem[warp - 2][0];
    const int r_sub = lane >> 3, gq = lane & 7, c4 = gq * 4;
    // Global addresses are (warp-uniform base of the tile / chunk) + (32-bit lane offset that never changes):
    // the offsets of this lane's 8 phase-B rows are computed once per kernel, so a store costs one instruction
    // instead of a 64-bit multiply-add and a bounds test (these were a quarter of the GLU epilogue's instructions).
    constexpr int kColShift = EPI == EPI_GLU ? 1 : EPI == EPI_LSTM ? 2 : 0;     // accumulator column -> output column
    uint32_t ooff[8], roff[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint32_t r = (uint32_t)(quarter * 32 + it * 4 + r_sub);
      ooff[it] = r * (uint32_t)ep.ldo + (uint32_t)(c4 >> kColShift);
      roff[it] = r * (uint32_t)ep.ldr + (uint32_t)(c4 >> kColShift);
    }
    uint32_t ti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      if ((int)(ti % NG) != group) continue;
      const uint32_t use = ti / NG;
      const int m0 = (t / n_tiles) * TC_BM, n0 = (t % n_tiles) * BN;
      const bool full = m0 + TC_BM <= M;                   // warp-uniform: only the last row tile tests rows
      const int row_lim = M - m0 - quarter * 32 - r_sub;   // row it is valid iff it * 4 < row_lim
      const uint32_t tmem_d = tmem_base + (uint32_t)group * (uint32_t)BN + ((uint32_t)(quarter * 32) << 16);
      ptx::mbar_wait(&sm->tmem_full[group], use & 1);
      SAST_STAMP(trc, quarter == 0 && lane == 0 && ti < 8, 48 + 4 * ti);
      ptx::tc_fence_after();
      {
        long long pixo[8];
        if (EPI == EPI_SCATTER) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = m0 + quarter * 32 + it * 4 + r_sub;
            pixo[it] = row < M ? (long long)ep.row_pix[row] * ep.C : -1;
          }
        }
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t raw[32];
          ptx::tmem_ld_32x32(tmem_d + (uint32_t)c0, raw);
          const int n = n0 + c0 + c4;
          const size_t ocol = (size_t)m0 * ep.ldo + (size_t)((n0 + c0) >> kColShift);    // warp-uniform
          const size_t rcol = (size_t)m0 * ep.ldr + (size_t)((n0 + c0) >> kColShift);
          float4 r4[8];
          if (EPI == EPI_RESID || EPI == EPI_SCATTER) {    // residual rows in flight while the TMEM load completes
            const float* rb = ep.resid + rcol;
            if (full) {
#pragma unroll
              for (int it = 0; it < 8; ++it) r4[it] = *reinterpret_cast<const float4*>(rb + roff[it]);
            } else {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                r4[it] = it * 4 < row_lim ? *reinterpret_cast<const float4*>(rb + roff[it]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          float cprev[8];
          if (EPI == EPI_LSTM) {
            const float* pb = ep.resid ? ep.resid + rcol : nullptr;
#pragma unroll
            for (int it = 0; it < 8; ++it) cprev[it] = (pb && (full || it * 4 < row_lim)) ? pb[roff[it]] : 0.f;
          }
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
          if ((EPI == EPI_RESID || EPI == EPI_SCATTER) && ep.gamma) g4 = __ldg(reinterpret_cast<const float4*>(ep.gamma + n));
          ptx::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<uint4*>(stage + lane * 32 + ((k ^ (lane & 7)) << 2)) = make_uint4(raw[4 * k], raw[4 * k + 1], raw[4 * k + 2], raw[4 * k + 3]);
          __syncwarp();
          // math for all 8 row-iterations is straight-line (independent chains interleave)
          float4 a4[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + r_sub;
            a4[it] = *reinterpret_cast<const float4*>(stage + r * 32 + ((gq ^ (r & 7)) << 2));
          }
          if (EPI == EPI_GLU) {                // columns interleaved value_j, gate_j
            __nv_bfloat162 o[8];
            const float2 hb = make_float2(0.5f * b4.x, 0.5f * b4.z);     // the 0.5 of the GELU rides on the value branch
#pragma unroll
            for (int it = 0; it < 8; ++it)
              o[it] = __floats2bfloat162_rn(glu_tanh_fit(fmaf(a4[it].x, 0.5f, hb.x), a4[it].y + b4.y),
                                            glu_tanh_fit(fmaf(a4[it].z, 0.5f, hb.y), a4[it].w + b4.w));
            __nv_bfloat16* ob = ep.out_bf16 + ocol;
            if (full) {
#pragma unroll
              for (int it = 0; it < 8; ++it) *reinterpret_cast<__nv_bfloat162*>(ob + ooff[it]) = o[it];
            } else {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (it * 4 < row_lim) *reinterpret_cast<__nv_bfloat162*>(ob + ooff[it]) = o[it];
            }
          } else if (EPI == EPI_LSTM) {        // columns interleaved forget, input, output, cell-input of one channel
            float hn[8], cn[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              // one MUFU.TANH per gate (10 MUFU per channel with exp + divide made this epilogue MUFU bound)
              const float f = sigmoid_fast(a4[it].x + b4.x), ig = sigmoid_fast(a4[it].y + b4.y), og = sigmoid_fast(a4[it].z + b4.z);
              const float g = tanh_fast(a4[it].w + b4.w);
              cn[it] = f * cprev[it] + ig * g;
              hn[it] = og * tanh_fast(cn[it]);
            }
            float* hb_ = ep.out_f32 + ocol;
            float* cb = ep.out2_f32 + ocol;
            if (full) {
#pragma unroll
              for (int it = 0; it < 8; ++it) { cb[ooff[it]] = cn[it]; hb_[ooff[it]] = hn[it]; }
            } else {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (it * 4 < row_lim) { cb[ooff[it]] = cn[it]; hb_[ooff[it]] = hn[it]; }
            }
          } else {
            float4 v4[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              v4[it] = make_float4(a4[it].x + b4.x, a4[it].y + b4.y, a4[it].z + b4.z, a4[it].w + b4.w);
              if (EPI == EPI_RESID || EPI == EPI_SCATTER) {
                v4[it].x = r4[it].x + g4.x * v4[it].x; v4[it].y = r4[it].y + g4.y * v4[it].y;
                v4[it].z = r4[it].z + g4.z * v4[it].z; v4[it].w = r4[it].w + g4.w * v4[it].w;
              }
            }
            if (EPI == EPI_SCATTER) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (pixo[it] >= 0) *reinterpret_cast<float4*>(ep.out_f32 + pixo[it] + n) = v4[it];
            } else {
              if (ep.out_f32) {
                float* of = ep.out_f32 + ocol;
#pragma unroll
                for (int it = 0; it < 8; ++it)
                  if (full || it * 4 < row_lim) *reinterpret_cast<float4*>(of + ooff[it]) = v4[it];
              }
              if (ep.out_bf16) {
                __nv_bfloat16* oh = ep.out_bf16 + ocol;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const __nv_bfloat162 lo = __floats2bfloat162_rn(v4[it].x, v4[it].y), hi = __floats2bfloat162_rn(v4[it].z, v4[it].w);
                  uint2 pk;
                  pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                  pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                  if (full || it * 4 < row_lim) *reinterpret_cast<uint2*>(oh + ooff[it]) = pk;
                }
              }
            }
          }
          __syncwarp();
        }
      }
      // this warp's TMEM reads are complete: hand the accumulator back to the MMA warp
      SAST_STAMP(trc, quarter == 0 && lane == 0 && ti < 8, 49 + 4 * ti);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm->tmem_empty[group]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
  SAST_STAMP(trc, threadIdx.x == 0, 2);
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;     // idempotent lookup; benign race
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] with leading dimension ld (elements); box = box_cols x box_rows,
// swizzle span = box_cols * 2 bytes (64 or 128)
int make_tmap_bf16_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows,
                       int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return (int)cudaErrorNotSupported;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_32B;
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// 2-D fp32 row-major [rows, cols]; box = box_cols (<= 32: 128-byte swizzle span) x box_rows
int make_tmap_f32_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return (int)cudaErrorNotSupported;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

int make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_rows) {
  return make_tmap_bf16_box(m, ptr, rows, cols, ld, TC_BK, box_rows, 128);
}

static int sm_count() {   // of the CURRENT device (no process-wide cache: one process may drive several GPUs)
  int n = 0, dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n;
}

template <int EPI>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mw, const float* bias, int N, int K, int BN, const int* counts,
                     long long max_rows, int m_static, const EpiParams& ep, cudaStream_t st,
                     const CUtensorMap* ma2 = nullptr, int k_split = 1 << 30) {
  const int stages = BN > 128 ? 3 : TC_STAGES;                          // 3 x 48 KB or 4 x <=32 KB of operand ring
  const size_t smem = 1024 + (size_t)stages * (TC_BM * TC_BK * 2 + (size_t)BN * TC_BK * 2) + 128;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);   // + <= 34 KB static
    if (e != cudaSuccess) return (int)e;
  }
  const long long tiles = ((max_rows + TC_BM - 1) / TC_BM) * (N / BN);       // worst case; the kernel clips to counts[1]
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  EpiParams ept = ep;
  ept.trace = g_trace_which == 2 ? g_trace : nullptr;
  sast::launch_k(gemm_tc_kernel<EPI>, grid, tc_threads(EPI), smem, st, ma, ma2 ? *ma2 : ma, k_split, mw, bias, N, K, BN, stages, counts, m_static, ept);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

// Widest N tile (multiple of 32, <= 256 so that two accumulators fit the 512 TMEM columns) that divides N:
// every extra n-tile re-reads the A tile and pays the per-tile hand-shakes again.
// ... unless that leaves SMs idle.  With few row tiles (late stages, small batches) the choice is a trade: every
// n-tile re-reads its A tile through L2 (measured: 1920x512x512 with BN=32 spends 7 of its 11 us on aggregate L2
// bandwidth), so below a full wave the tile stays >= 128 wide (>= 64 when even that leaves under ~40 tiles).
static int pick_bn(int N, long long m_tiles = 1 << 20, int groups = 2) {
  const int cap = 512 / groups / 32 * 32;                          // `groups` accumulators must fit the 512 TMEM columns
  int widest = 0, le128 = 0, le64 = 0;
  for (int bn = cap; bn >= 32; bn -= 32) {
    if (N % bn != 0) continue;
    if (!widest) widest = bn;
    if (m_tiles * (N / bn) >= sm_count()) return bn;               // widest tile that still fills the chip
    if (bn <= 128 && !le128) le128 = bn;
    if (bn <= 64 && !le64) le64 = bn;
  }
  if (le128 && m_tiles * (N / le128) >= 40) return le128;
  if (le64) return le64;
  return le128 ? le128 : widest;
}

int launch_gemm_tc(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, const float* bias, int N, int K, const int* counts,
                   long long max_rows, int epi, const EpiParams& ep, cudaStream_t st) {
  const int BN = pick_bn(N, (max_rows + TC_BM - 1) / TC_BM, tc_groups(epi));
  if (BN == 0 || K % 8 != 0 || lda % 8 != 0) return SAST_E_SHAPE;
  CUtensorMap ma, mw;
  int rc = make_tmap_bf16_2d(&ma, A, max_rows, K, lda, TC_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&mw, W, N, K, K, BN);
  if (rc) return rc;
  switch (epi) {
    case EPI_STORE: return launch_tc<EPI_STORE>(ma, mw, bias, N, K, BN, counts, max_rows, 0, ep, st);
    case EPI_RESID: return launch_tc<EPI_RESID>(ma, mw, bias, N, K, BN, counts, max_rows, 0, ep, st);
    case EPI_GLU: return launch_tc<EPI_GLU>(ma, mw, bias, N, K, BN, counts, max_rows, 0, ep, st);
    case EPI_SCATTER: return launch_tc<EPI_SCATTER>(ma, mw, bias, N, K, BN, counts, max_rows, 0, ep, st);
  }
  return SAST_E_UNSUPPORTED;
}

}  // namespace sast

// h, c = LSTM cell on the 1x1-conv of [x | h_prev]  (models/layers/rnn.py:36-69, dws_conv False) as ONE kernel:
// TF32 tcgen05 GEMM straight from the fp32 NHWC maps + gate epilogue.  w_packed [4C, K] fp32 with rows interleaved
// 4*c + {f,i,o,g} (K = C for a zero initial state, 2C otherwise), bias_packed [4C] likewise.
extern "C" int sast_lstm_fwd(const float* x, const float* h_prev, const float* c_prev, const float* w_packed,
                             const float* bias_packed, int64_t P, int32_t C, float* h_out, float* c_out, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(x); SAST_CHECK_PTR(w_packed); SAST_CHECK_PTR(h_out); SAST_CHECK_PTR(c_out);
  if (P <= 0 || C <= 0 || C % 8 != 0 || P >= (1ll << 31)) return SAST_E_SHAPE;
  if ((h_prev == nullptr) != (c_prev == nullptr)) return SAST_E_NULL;
  const int K = h_prev ? 2 * C : C, N = 4 * C;
  const int BN = pick_bn(N, (P + TC_BM - 1) / TC_BM, tc_groups(EPI_LSTM));
  if (BN == 0) return SAST_E_SHAPE;
  CUtensorMap ma, ma2, mw;
  int rc = make_tmap_f32_box(&ma, x, P, C, C, 32, TC_BM);
  if (rc) return rc;
  if (h_prev) { rc = make_tmap_f32_box(&ma2, h_prev, P, C, C, 32, TC_BM); if (rc) return rc; }
  rc = make_tmap_f32_box(&mw, w_packed, N, K, K, 32, BN);
  if (rc) return rc;
  EpiParams ep{};
  ep.out_f32 = h_out; ep.out2_f32 = c_out; ep.ldo = C; ep.resid = c_prev; ep.ldr = C;
  return launch_tc<EPI_LSTM>(ma, mw, bias_packed, N, K, BN, nullptr, P, (int)P, ep, (cudaStream_t)stream, h_prev ? &ma2 : nullptr,
                             h_prev ? C : (1 << 30));
}

// D[M, N/2] (bf16) = GLU(A W^T + bias) with W rows interleaved value_j, gate_j: the layer's MLP-in GEMM, standalone
extern "C" int sast_gemm_bf16_glu(const uint16_t* A, const uint16_t* W, const float* bias, uint16_t* D, int32_t M, int32_t N,
                                  int32_t K, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(A); SAST_CHECK_PTR(W); SAST_CHECK_PTR(D);
  if (M <= 0 || N <= 0 || K <= 0 || N % 64 != 0) return SAST_E_SHAPE;
  const int BN = pick_bn(N, (M + TC_BM - 1) / TC_BM, tc_groups(EPI_GLU));
  if (BN == 0 || K % 8 != 0) return SAST_E_SHAPE;
  CUtensorMap ma, mw;
  int rc = make_tmap_bf16_2d(&ma, A, M, K, K, TC_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&mw, W, N, K, K, BN);
  if (rc) return rc;
  EpiParams ep{};
  ep.ldo = N / 2;
  ep.out_bf16 = (__nv_bfloat16*)D;
  return launch_tc<EPI_GLU>(ma, mw, bias, N, K, BN, nullptr, M, M, ep, (cudaStream_t)stream);
}

extern "C" int sast_gemm_bf16(const uint16_t* A, const uint16_t* W, const float* bias, void* D, int32_t d_is_bf16, int32_t M,
                              int32_t N, int32_t K, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(A); SAST_CHECK_PTR(W); SAST_CHECK_PTR(D);
  if (M <= 0 || N <= 0 || K <= 0) return SAST_E_SHAPE;
  const int BN = pick_bn(N, (M + TC_BM - 1) / TC_BM, tc_groups(EPI_STORE));
  if (BN == 0 || K % 8 != 0) return SAST_E_SHAPE;
  CUtensorMap ma, mw;
  int rc = make_tmap_bf16_2d(&ma, A, M, K, K, TC_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&mw, W, N, K, K, BN);
  if (rc) return rc;
  EpiParams ep{};
  ep.ldo = N;
  if (d_is_bf16) ep.out_bf16 = (__nv_bfloat16*)D; else ep.out_f32 = (float*)D;
  return launch_tc<EPI_STORE>(ma, mw, bias, N, K, BN, nullptr, M, M, ep, (cudaStream_t)stream);
}
