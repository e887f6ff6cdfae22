// a4 on the tensor cores: scoring GEMM  s = relu((x + pos) Ws^T + b)  as 3xTF32 tcgen05.mma.
//
// Selection thresholds a softmax of sum_c |amp/ctrl_c * s_c| (SURVEY.md "hard part 1"): a bf16 GEMM
// flips ~1e-3 of the tokens, so the product is kept fp32-accurate by error compensation:
//     a = a_hi + a_lo,  w = w_hi + w_lo   (each half rounded to TF32, 1+10-bit significand)
//     a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi          (dropped term ~2^-22 |a||w|)
// Three kind::tf32 MMAs per 8-wide k-step, fp32 accumulation in TMEM.
//
// Persistent CTAs, 13 warps:
//   warps 0-3  loaders: x + pos formed on the fly (coalesced float4 rows), split into hi/lo and written
//              to shared memory in the SWIZZLE_128B K-major operand layout; an elected lane of warp 0 also
//              TMA-loads the pre-split weight tiles (W_hi, W_lo: [C,C] fp32, K-major like nn.Linear.weight)
//   warps 4-11 epilogue, two groups of 4 (one TMEM accumulator each, tiles round-robin): TMEM -> swizzled
//              128-bit shared-memory transpose -> coalesced rows: STP-weighted map xw = sigmoid(ctrl) sigmoid(s) x0
//              and the per-token L1  sum_c |amp/ctrl_c s_c|; every global load of a 32-column chunk is issued
//              before anything waits on it.
//   warp 12    TMEM allocator + MMA issuer (whole warp on uniform values, elected lane issues)
// Measured (tools/gemm_trace.py score): loaders and epilogue each need 7-8 k clk per 128-row tile, the MMAs ~0.8 k;
// with only two ring stages the loaders pay one DRAM round trip per k-block.  A deeper ring plus prefetching
// loaders is the next step (tried with 17 warps: the 96-register budget then starved the epilogue).
#include "common.cuh"
#include "ptx.cuh"

namespace sast {

constexpr int SC_BM = 128, SC_BK = 32, SC_STAGES = 2;
constexpr int SC_GROUPS = 2;                     // epilogue groups = TMEM accumulators
constexpr int SC_MMA_WARP = 4 + 4 * SC_GROUPS;
constexpr int SC_THREADS = (SC_MMA_WARP + 1) * 32;

struct ScSmem {
  uint64_t full[SC_STAGES];
  uint64_t empty[SC_STAGES];
  uint64_t tmem_full[SC_GROUPS];
  uint64_t tmem_empty[SC_GROUPS];
  uint32_t tmem_base;
};

// round to TF32 (10-bit mantissa), nearest with ties away from zero == cvt.rna.tf32.f32 on finite values, as two integer
// instructions: the conversion unit retires 16 cvt per clock and SM, and the loaders need 8 192 of them per k-block (a
// quarter of their time in the phase tracer)
__device__ __forceinline__ float to_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u); }
using ptx::umma_tf32_ss;
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);   // D=F32, A=B=TF32, K-major
}

__global__ void __launch_bounds__(SC_THREADS, 1) score_tc_kernel(const __grid_constant__ CUtensorMap map_whi,
                                                                 const __grid_constant__ CUtensorMap map_wlo,
                                                                 const float* __restrict__ x, const float* __restrict__ pos,
                                                                 long long pos_bstride, const float* __restrict__ bs,
                                                                 const float* __restrict__ r, const float* __restrict__ ctrl_w,
                                                                 int n_bins, float amp, int B,
                                                                 int HW, int C, int BN, long long P, float* __restrict__ xw,
                                                                 float* __restrict__ l1_out, long long* __restrict__ trace) {
  // trace build only: same [CTA][128] slot layout as gemm_tc_kernel (0 entry, 1 set-up done, 2 end; per tile ti < 8: MMA
  // 8+4ti {accumulator free, first k-block ready, last commit}, epilogue quarter 0: 48+4ti {full, drained}, loaders 100+ti)
  [[maybe_unused]] long long* const trc = trace ? trace + (size_t)blockIdx.x * 128 : nullptr;
  SAST_STAMP(trc, threadIdx.x == 0, 0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];     // ring | ScSmem | epilogue transpose tiles
  const int m_tiles = (int)((P + SC_BM - 1) / SC_BM), n_tiles = C / BN;
  const int total_tiles = m_tiles * n_tiles;
  uint8_t* base = smem_raw;                                       // 1024-byte aligned (checked below)
  const uint32_t a_bytes = SC_BM * SC_BK * 4;                     // 16 KB per half
  const uint32_t w_bytes = (uint32_t)BN * SC_BK * 4;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * w_bytes;         // A_hi, A_lo, W_hi, W_lo
  ScSmem* sm = reinterpret_cast<ScSmem*>(smem_raw + (size_t)SC_STAGES * stage_bytes);
  float* stage_all = reinterpret_cast<float*>(smem_raw + (size_t)SC_STAGES * stage_bytes + 256);
  // controls of the STP weighting, per (frame, channel): sig = sigmoid(ctrl), inv = amp / ctrl (SAST.py:105-119 with the
  // PositiveLinear of :305-328: ctrl[b,c] = sum_j exp(Wc[c,j]) (r[b,j] + 1e-6)).  Every CTA computes the whole table (B x C
  // x n_bins exp: a few hundred per thread) into shared memory BEFORE the PDL wait -- r comes from the very first kernels of
  // the forward -- instead of a separate 5 us launch per block in front of this kernel.
  float* const sig = stage_all + (size_t)4 * SC_GROUPS * 32 * 33;
  float* const inv = sig + (size_t)B * C;
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;    // warp: provably uniform
  const int nkb = C / SC_BK;

  if (threadIdx.x == 0) {
    ptx::tma_prefetch_desc(&map_whi);
    ptx::tma_prefetch_desc(&map_wlo);
    for (int s = 0; s < SC_STAGES; ++s) { ptx::mbar_init(&sm->full[s], 5 /* 4 loader warps + the TMA expect_tx arrive */); ptx::mbar_init(&sm->empty[s], 1); }
    for (int a = 0; a < SC_GROUPS; ++a) { ptx::mbar_init(&sm->tmem_full[a], 1); ptx::mbar_init(&sm->tmem_empty[a], 4); }
    ptx::fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(SC_GROUPS * BN)) tmem_cols <<= 1;
  if (warp == SC_MMA_WARP) ptx::tmem_alloc(&sm->tmem_base, tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = sm->tmem_base;
  for (int e = threadIdx.x; e < B * C; e += blockDim.x) {
    const int b = e / C, c = e - b * C;
    float acc = 0.f;
    for (int j = 0; j < n_bins; ++j) acc += expf(ctrl_w[c * n_bins + j]) * (r[b * n_bins + j] + 1e-6f);
    sig[e] = sigmoidf_acc(acc);
    float iv = amp / acc;
    if (isinf(iv)) iv = 0.f;
    inv[e] = iv;
  }
  __syncthreads();
  // PDL: everything above (barriers, TMEM, the control table from data written >= 2 kernels ago) overlapped the tail of the
  // preceding kernel; its output is read only from here on
  pdl_entry();

  SAST_STAMP(trc, threadIdx.x == 0, 1);
  if (warp < 4) {
    // ---------------- loaders ----------------
    const int t = threadIdx.x;                 // 0..127
    const int chunk = t & 7;                   // 16-byte chunk of the 128-byte k-block row
    const int rbase = t >> 3;                  // rows rbase, rbase+16, ... (8 rows per thread)
    uint32_t it = 0;
    [[maybe_unused]] uint32_t pti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++pti) {
      const int m0 = (tile / n_tiles) * SC_BM;
      const int n0 = (tile % n_tiles) * BN;
      int xo[8], po[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int tok = m0 + rbase + 16 * j;
        const bool ok = tok < (int)P;
        const int b = ok ? tok / HW : 0;
        xo[j] = ok ? tok * C + chunk * 4 : -1;
        po[j] = (int)(b * pos_bstride) + (ok ? tok - b * HW : 0) * C + chunk * 4;
      }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % SC_STAGES, round = it / SC_STAGES;
        SAST_STAMP(trc, threadIdx.x == 0 && pti == 2 && kb < 2, 110 + 4 * kb);
        float4 xv[8], pv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {         // global loads first: they do not depend on the ring slot
          if (xo[j] >= 0) {
            xv[j] = *reinterpret_cast<const float4*>(x + xo[j] + kb * SC_BK);
            pv[j] = *reinterpret_cast<const float4*>(pos + po[j] + kb * SC_BK);
          } else {
            xv[j] = make_float4(0.f, 0.f, 0.f, 0.f); pv[j] = xv[j];
          }
        }
        ptx::mbar_wait(&sm->empty[s], (round & 1) ^ 1);
        SAST_STAMP(trc, threadIdx.x == 0 && kb == 0 && pti < 8, 100 + pti);
        SAST_STAMP(trc, threadIdx.x == 0 && pti == 2 && kb < 2, 111 + 4 * kb);
        uint8_t* st = base + (size_t)s * stage_bytes;
        if (warp == 0 && ptx::elect_one()) {     // warp-uniform operands, elected lane: no R2UR waterfall per TMA
          ptx::mbar_arrive_expect_tx(&sm->full[s], 2 * w_bytes);
          ptx::tma_load_2d(st + 2 * a_bytes, &map_whi, &sm->full[s], kb * SC_BK, n0);
          ptx::tma_load_2d(st + 2 * a_bytes + w_bytes, &map_wlo, &sm->full[s], kb * SC_BK, n0);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = rbase + 16 * j;
          const float a0 = xv[j].x + pv[j].x, a1 = xv[j].y + pv[j].y, a2 = xv[j].z + pv[j].z, a3 = xv[j].w + pv[j].w;
          const float h0 = to_tf32(a0), h1 = to_tf32(a1), h2 = to_tf32(a2), h3 = to_tf32(a3);
          const uint32_t off = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128 + (uint32_t)((chunk ^ (r & 7)) * 16);
          *reinterpret_cast<float4*>(st + off) = make_float4(h0, h1, h2, h3);
          *reinterpret_cast<float4*>(st + a_bytes + off) = make_float4(to_tf32(a0 - h0), to_tf32(a1 - h1), to_tf32(a2 - h2), to_tf32(a3 - h3));
        }
        SAST_STAMP(trc, threadIdx.x == 0 && pti == 2 && kb < 2, 112 + 4 * kb);
        ptx::fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&sm->full[s]);     // one arrive per warp, not per thread
        SAST_STAMP(trc, threadIdx.x == 0 && pti == 2 && kb < 2, 113 + 4 * kb);
      }
    }
  } else if (warp == SC_MMA_WARP) {
    // ---------------- MMA issuer: whole warp on uniform values, one elected lane issues (no R2UR waterfall per UTCHMMA) ----------------
    const bool leader = ptx::elect_one();
    const uint32_t idesc = idesc_tf32(SC_BM, (uint32_t)BN);
    uint32_t it = 0, ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti % SC_GROUPS, use = ti / SC_GROUPS;
      ptx::mbar_wait(&sm->tmem_empty[acc], (use & 1) ^ 1);
      SAST_STAMP(trc, leader && ti < 8, 8 + 4 * ti);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % SC_STAGES, round = it / SC_STAGES;
        ptx::mbar_wait(&sm->full[s], round & 1);
        SAST_STAMP(trc, leader && kb == 0 && ti < 8, 9 + 4 * ti);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(base + (size_t)s * stage_bytes);
        const uint64_t dah = ptx::umma_desc_sw128_kmajor(sa), dal = ptx::umma_desc_sw128_kmajor(sa + a_bytes);
        const uint64_t dwh = ptx::umma_desc_sw128_kmajor(sa + 2 * a_bytes), dwl = ptx::umma_desc_sw128_kmajor(sa + 2 * a_bytes + w_bytes);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {          // 8 tf32 = 32 bytes per k-step: +2 in the >>4 address field
            const uint64_t o = (uint64_t)(k * 2);
            umma_tf32_ss(tmem_d, dal + o, dwh + o, idesc, (kb | k) ? 1u : 0u);    // small terms first
            umma_tf32_ss(tmem_d, dah + o, dwl + o, idesc, 1u);
            umma_tf32_ss(tmem_d, dah + o, dwh + o, idesc, 1u);
          }
          ptx::umma_commit(&sm->empty[s]);
        }
      }
      if (leader) ptx::umma_commit(&sm->tmem_full[acc]);
      SAST_STAMP(trc, leader && ti < 8, 10 + 4 * ti);
    }
  } else {
    // ---------------- epilogue (warps 4..15: group = (warp-4)/4, TMEM lane quarter = warp % 4) ----------------
    const int group = (warp - 4) >> 2;
    const int quarter = warp & 3;
    float* stage = stage_all + (warp - 4) * (32 * 33);
    const int r_sub = lane >> 3, c4 = (lane & 7) * 4;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      if ((int)(ti % SC_GROUPS) != group) continue;
      const uint32_t use = ti / SC_GROUPS;
      const int m0 = (tile / n_tiles) * SC_BM;
      const int n_tile = tile % n_tiles, n0 = n_tile * BN;
      const uint32_t tmem_d = tmem_base + (uint32_t)group * (uint32_t)BN + ((uint32_t)(quarter * 32) << 16);
      // per-row element offsets of this lane's 8 rows (hoisted out of the column loop; < 2^31, checked on the host)
      int xo[8], po[8], co[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int tok = m0 + quarter * 32 + i * 4 + r_sub;
        const bool ok = tok < (int)P;
        const int b = ok ? tok / HW : 0;
        xo[i] = ok ? tok * C + c4 : -1;
        po[i] = (int)(b * pos_bstride) + (ok ? tok - b * HW : 0) * C + c4;
        co[i] = b * C + c4;
      }
      ptx::mbar_wait(&sm->tmem_full[group], use & 1);
      SAST_STAMP(trc, quarter == 0 && lane == 0 && ti < 8, 48 + 4 * ti);
      ptx::tc_fence_after();
      float l1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) l1[i] = 0.f;
      // sig / inv are per (frame, channel): one load per chunk unless the tile straddles two frames (late stages only)
      const bool one_frame = (m0 / HW) == ((min(m0 + SC_BM, (int)P) - 1) / HW);
      const int gq = lane & 7;
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tmem_d + (uint32_t)c0, raw);
        const int nc = n0 + c0;
        // every global load of the chunk is issued before anything waits (a `continue` inside the row loop used to
        // serialise 8 x 3 dependent L2 round trips per chunk: ~11k clk per tile)
        float4 xv[8], pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (xo[i] >= 0) {
            xv[i] = *reinterpret_cast<const float4*>(x + xo[i] + nc);
            pv[i] = __ldg(reinterpret_cast<const float4*>(pos + po[i] + nc));      // small table, cache resident
          } else {
            xv[i] = make_float4(0.f, 0.f, 0.f, 0.f); pv[i] = xv[i];
          }
        }
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bs + nc + c4));
        float4 sg0 = *reinterpret_cast<const float4*>(sig + co[0] + nc);
        float4 iv0 = *reinterpret_cast<const float4*>(inv + co[0] + nc);
        ptx::tmem_ld_wait();
        // transpose through shared memory: 8 x STS.128 / 8 x LDS.128, 16-byte groups XOR-swizzled by (row % 8)
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<uint4*>(stage + lane * 32 + ((k ^ (lane & 7)) << 2)) = make_uint4(raw[4 * k], raw[4 * k + 1], raw[4 * k + 2], raw[4 * k + 3]);
        __syncwarp();
        float4 a4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + r_sub;
          a4[i] = *reinterpret_cast<const float4*>(stage + r * 32 + ((gq ^ (r & 7)) << 2));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 sg = sg0, iv = iv0;
          if (!one_frame) {                       // warp-uniform
            sg = *reinterpret_cast<const float4*>(sig + co[i] + nc);
            iv = *reinterpret_cast<const float4*>(inv + co[i] + nc);
          }
          const float s0 = fmaxf(a4[i].x + b4.x, 0.f), s1 = fmaxf(a4[i].y + b4.y, 0.f), s2 = fmaxf(a4[i].z + b4.z, 0.f),
                      s3 = fmaxf(a4[i].w + b4.w, 0.f);
          float4 o;
          o.x = (sg.x * sigmoid_fast(s0)) * (xv[i].x + pv[i].x);
          o.y = (sg.y * sigmoid_fast(s1)) * (xv[i].y + pv[i].y);
          o.z = (sg.z * sigmoid_fast(s2)) * (xv[i].z + pv[i].z);
          o.w = (sg.w * sigmoid_fast(s3)) * (xv[i].w + pv[i].w);
          if (xo[i] >= 0) *reinterpret_cast<float4*>(xw + xo[i] + nc) = o;
          l1[i] += (fabsf(iv.x * s0) + fabsf(iv.y * s1)) + (fabsf(iv.z * s2) + fabsf(iv.w * s3));
        }
        __syncwarp();
      }
      SAST_STAMP(trc, quarter == 0 && lane == 0 && ti < 8, 49 + 4 * ti);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm->tmem_empty[group]);
      // per-token L1 over this tile's BN channels: reduce the 8 lanes that share a row
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = l1[i];
        v += __shfl_xor_sync(kFull, v, 1);
        v += __shfl_xor_sync(kFull, v, 2);
        v += __shfl_xor_sync(kFull, v, 4);
        const int tok = m0 + quarter * 32 + i * 4 + r_sub;
        if ((lane & 7) == 0 && tok < (int)P) l1_out[(long long)n_tile * P + tok] = v;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == SC_MMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
  SAST_STAMP(trc, threadIdx.x == 0, 2);
}

int make_tmap_f32_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows);

// returns the number of channel slices (partials) written to l1_part, or <0 / >0 on error (negated for CUDA errors)
int launch_score_tc(const sast_score_args* a, float* l1_part, int* n_slices, cudaStream_t st) {
  const sast_geom& g = a->g;
  const long long P = (long long)g.B * g.H * g.W;
  const int C = g.C;
  const int BN = C % 128 == 0 ? 128 : (C % 64 == 0 ? 64 : 32);
  CUtensorMap mh, ml;
  int rc = make_tmap_f32_box(&mh, a->score_w_hi, C, C, C, SC_BK, BN);
  if (rc) return rc;
  rc = make_tmap_f32_box(&ml, a->score_w_lo, C, C, C, SC_BK, BN);
  if (rc) return rc;
  if (P * C >= (1ll << 31)) return SAST_E_UNSUPPORTED;        // 32-bit element offsets inside the kernel
  const size_t smem = (size_t)SC_STAGES * (2 * SC_BM * SC_BK * 4 + 2 * (size_t)BN * SC_BK * 4) + 256 +
                      (size_t)4 * SC_GROUPS * 32 * 33 * sizeof(float) + (size_t)2 * g.B * C * sizeof(float);
  if (smem > 200 * 1024) return SAST_E_UNSUPPORTED;             // (B x C control table: 32 KB at B = 8, C = 512)
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
  }
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = ((P + SC_BM - 1) / SC_BM) * (C / BN);
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  sast::launch_k(score_tc_kernel, grid, SC_THREADS, smem, st, mh, ml, a->x, a->pos, a->pos_batch_stride, a->score_b, a->r, a->ctrl_w, a->n_bins, a->amp, g.B, g.H * g.W, C, BN, P,
                                                  a->xw, l1_part, g_trace_which == 3 ? g_trace : nullptr);
  SAST_LAUNCH_CHECK();
  *n_slices = C / BN;
  return SAST_OK;
}

}  // namespace sast
