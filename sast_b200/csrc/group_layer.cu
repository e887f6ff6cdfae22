// One MS-WSA layer as ONE kernel for the late stages (SAST_BF16 path, C = 256 or 512: 7 680 / 1 920 tokens at 1 Mpx B=8).
//
// There are only 64 / 16 tiles of 128 compacted rows here and 1.5 / 6.2 MB of bf16 weights per layer: one CTA per tile
// (fused_layer.cu) would leave most SMs idle while each busy one streams every weight matrix.  Instead a GROUP of
// G = C / 64 CTAs works on one tile, CTA s of the group owning heads 2s, 2s+1 and a 64-column slice of every
// token-wise GEMM:
//   LN            rows 128/G .. of the tile: x[row_pix] -> LN1 -> LN2 -> n2 (bf16 operand rows + fp32 shortcut)  -> scratch
//   QKV           n2[128 x C] (streamed) x Wqkv[192 rows of slice s]^T          -> Q,K,V tiles of 2 heads in shared memory
//   attention     exactly the per-pair code of the one-CTA kernel (S in TMEM, P back to TMEM, ts-form P V)  -> att slice -> scratch
//   proj          att[128 x C] (streamed) x Wp[64 rows]^T -> y slice = n2 + g1 (o + b) (fp32 in registers; bf16 -> scratch)
//   GLU           y[128 x C] (streamed, two passes) x W1[336 rows]^T -> val * gelu(gate) -> hid slice [128 x 168]      -> scratch
//   MLP out       hid[128 x I] (streamed) x W2[64 rows]^T -> out[row_pix][64-column slice] = y + g2 (m + b)
// Operands stream through a 3 x 40 KB TMA ring (A k-block 16 KB + weight k-block <= 24 KB) fed by a 17th warp; the
// slices the CTAs exchange (n2, att, y, hid: bf16, <= 344 KB per tile) go through an L2-resident scratch area, with a
// group barrier (global atomic flag, all CTAs of the <= 148-CTA grid are co-resident) between the phases.
// Every weight byte is read once per tile and the tile's activations G times -- from L2.
#include "fused_common.cuh"
#include <cstdlib>

namespace sast {

int make_tmap_bf16_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows,
                       int swizzle_bytes);

namespace gl {

using namespace fl;

constexpr int kStages = 3, kStageBytes = 40960, kWOff = 16384;      // stage = A [128 x 64] bf16 + W [<= 192 x 64] bf16
constexpr int kGluA = 176, kGluB = 160;                              // accumulator columns of the two GLU passes (336 per slice)

template <int C_>
struct Cfg {
  static constexpr int C = C_;
  static constexpr int G = C / 64;                  // CTAs per tile
  static constexpr int I = (C * 8 / 3) / 32 * 32;   // GLU width for mlp_ratio 4: 672 / 1344
  static constexpr int IS = I / G;                  // hid columns per slice (168)
  static constexpr int KB = C / 64;                 // k-blocks of the C-deep GEMMs
  static constexpr int KBI = (I + 63) / 64;         // k-blocks of the MLP-out GEMM
  static constexpr int RPC = 128 / G;               // LN rows per CTA
  static constexpr int LPT = 512 / RPC;             // lanes per LN row (16 / 32), 4 float4 each
  static constexpr int R_BYTES = 6 * 8192;          // Q0 K0 Q1 K1 V0 V1 (SWIZZLE_64B); later the staged output rows
  static constexpr int OUT_PITCH = 64 * 4 + 16;
  static constexpr int PV_LN = 0, PV_QKVB = 4 * C, PV_PROJB = PV_QKVB + 192, PV_G1 = PV_PROJB + 64, PV_B1 = PV_G1 + 64,
                       PV_B2 = PV_B1 + 2 * IS, PV_G2 = PV_B2 + 64, PV_FLOATS = PV_G2 + 64;
  static constexpr int CHUNKS = 4 * KB + KBI;       // ring chunks per tile: QKV, proj, GLU a, GLU b (KB each), out (KBI)
  static_assert(C == 256 || C == 512, "group layer kernel: C = 256 or 512");
  static_assert(2 * IS == kGluA + kGluB && IS % 8 == 0, "GLU slice");
  static_assert(128 * OUT_PITCH <= R_BYTES && 128 * IS * 2 <= R_BYTES, "output / hid staging");
};

struct Ctl {
  uint64_t full[kStages], empty[kStages];
  uint64_t mma_bar;
  uint64_t go[4];              // group barrier k passed: the producer may stream the phase's A operand
  uint32_t tmem_base;
  int pix[128];
  uint8_t lo[128], hi[128];
  float pmax[2][2][128], psum[2][2][128];
};

struct Params {
  const float* x;
  float* out;
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *qkv_b, *proj_b, *gamma1, *gamma2, *mlp1_b, *mlp2_b;
  float eps;
  const int *counts, *tile_list, *row_pix, *row_win, *tok_row;
  Geom g;
  int flavor;
  // scratch, per group: rows [group * 128, +128)
  __nv_bfloat16 *n2h, *att, *yh, *hid;
  float* n2f;
  int* flags;                  // [n_groups][4], zero on entry
  long long* trace;            // debug stamps (sast_debug_trace which = 5), normally null; trace build only
};

// trace build only: [CTA][32] clock64 stamps of thread 0 at the phase boundaries of the CTA's FIRST tile; 20 kernel entry,
// 21 set-up done, 22 tiles done, 23 unselected pass done
#define GL_STAMP(i) SAST_STAMP(trc, tid == 0 && tile_it == 0, (i))

template <int LPT>
__device__ __forceinline__ float lanes_sum(float v) {
#pragma unroll
  for (int o = LPT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// LayerNorm of one row spread over LPT lanes (4 float4 each: float4 index l + LPT i)
template <int C, int LPT>
__device__ __forceinline__ void ln_row(float4 (&v)[4], int l, const float* __restrict__ w, const float* __restrict__ b, float eps) {
  const float inv_c = 1.0f / (float)C;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = lanes_sum<LPT>(s) * inv_c;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = v[i].x - mean, bq = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + bq * bq) + (c * c + d * d);
  }
  const float rstd = rsqrtf(lanes_sum<LPT>(ss) * inv_c + eps);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 w4 = *reinterpret_cast<const float4*>(w + (l + LPT * i) * 4);
    const float4 b4 = *reinterpret_cast<const float4*>(b + (l + LPT * i) * 4);
    v[i].x = (v[i].x - mean) * rstd * w4.x + b4.x; v[i].y = (v[i].y - mean) * rstd * w4.y + b4.y;
    v[i].z = (v[i].z - mean) * rstd * w4.z + b4.z; v[i].w = (v[i].w - mean) * rstd * w4.w + b4.w;
  }
}

template <int C>
__global__ void __launch_bounds__(544, 1)
layer_group_kernel(const __grid_constant__ CUtensorMap map_n2, const __grid_constant__ CUtensorMap map_att,
                   const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_hid,
                   const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_proj,
                   const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
                   const __grid_constant__ CUtensorMap map_hid_st, const Params p) {
  using K = Cfg<C>;
  constexpr int G = K::G, I = K::I, IS = K::IS, KB = K::KB, KBI = K::KBI, LPT = K::LPT;
  extern __shared__ __align__(1024) uint8_t base[];
  const uint32_t sRing = ptx::smem_u32(base);
  if ((sRing & 1023u) != 0) __trap();
  const uint32_t sR = sRing + kStages * kStageBytes;
  Ctl* ctl = reinterpret_cast<Ctl*>(base + kStages * kStageBytes + K::R_BYTES);
  float* const spv = reinterpret_cast<float*>(base + kStages * kStageBytes + K::R_BYTES + ((sizeof(Ctl) + 15) / 16) * 16);
  const uint32_t sPV = sR + K::R_BYTES + ((sizeof(Ctl) + 15) / 16) * 16;

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(kFull, tid >> 5, 0), lane = tid & 31;
  [[maybe_unused]] long long* const trc = p.trace ? p.trace + (size_t)blockIdx.x * 32 : nullptr;
  SAST_STAMP(trc, tid == 0, 20);
  const int group = blockIdx.x / G, slice = blockIdx.x % G, n_groups = gridDim.x / G;

  if (tid == 0) {
    ptx::tma_prefetch_desc(&map_n2); ptx::tma_prefetch_desc(&map_att); ptx::tma_prefetch_desc(&map_y); ptx::tma_prefetch_desc(&map_hid);
    ptx::tma_prefetch_desc(&map_qkv); ptx::tma_prefetch_desc(&map_proj); ptx::tma_prefetch_desc(&map_w1); ptx::tma_prefetch_desc(&map_w2); ptx::tma_prefetch_desc(&map_hid_st);
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&ctl->full[s], 1); ptx::mbar_init(&ctl->empty[s], 1); }
    ptx::mbar_init(&ctl->mma_bar, 1);
    for (int k = 0; k < 4; ++k) ptx::mbar_init(&ctl->go[k], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc(&ctl->tmem_base, 512);
  for (int i = tid; i < K::PV_FLOATS; i += blockDim.x) {   // per-channel parameters of this slice (constants of the forward)
    float v;
    if (i < C) v = p.ln1_w[i];
    else if (i < 2 * C) v = p.ln1_b[i - C];
    else if (i < 3 * C) v = p.ln2_w[i - 2 * C];
    else if (i < 4 * C) v = p.ln2_b[i - 3 * C];
    else if (i < K::PV_PROJB) v = p.qkv_b ? p.qkv_b[192 * slice + i - K::PV_QKVB] : 0.f;
    else if (i < K::PV_G1) v = p.proj_b ? p.proj_b[64 * slice + i - K::PV_PROJB] : 0.f;
    else if (i < K::PV_B1) v = p.gamma1 ? p.gamma1[64 * slice + i - K::PV_G1] : 1.f;
    else if (i < K::PV_B2) v = p.mlp1_b ? p.mlp1_b[2 * IS * slice + i - K::PV_B1] : 0.f;
    else if (i < K::PV_G2) v = p.mlp2_b ? p.mlp2_b[64 * slice + i - K::PV_B2] : 0.f;
    else v = p.gamma2 ? p.gamma2[64 * slice + i - K::PV_G2] : 1.f;
    spv[i] = v;
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  pdl_entry();
  const int n_tiles = p.counts[3];
  SAST_STAMP(trc, tid == 0, 21);
  const int srow = group * 128;                              // this group's rows of the scratch buffers

  if (warp == 16) {
    // ---------------- operand producer: per tile 4 KB + KBI chunks, each gated by the group barrier of its phase ----------------
    const bool leader = ptx::elect_one();
    uint32_t it = 0, tile_it = 0;
    for (int t = group; t < n_tiles; t += n_groups, ++tile_it) {
      for (int ph = 0; ph < 5; ++ph) {                       // QKV, proj, GLU a, GLU b, out
        const CUtensorMap* ma = ph == 0 ? &map_n2 : ph == 1 ? &map_att : ph == 4 ? &map_hid : &map_y;
        const CUtensorMap* mw = ph == 0 ? &map_qkv : ph == 1 ? &map_proj : ph == 4 ? &map_w2 : &map_w1;
        const int wrow = ph == 0 ? 192 * slice : ph == 1 ? 64 * slice : ph == 2 ? 2 * IS * slice : ph == 3 ? 2 * IS * slice + kGluA : 64 * slice;
        const uint32_t wbytes = ph == 0 ? 192 * 128 : (ph == 2 || ph == 3) ? kGluA * 128 : 64 * 128;
        const int nkb = ph == 4 ? KBI : KB;
        // the weight chunks do not depend on the partners: the first ring-full of them is in flight BEFORE the phase's group
        // barrier opens; the operand chunks (written by the partners) follow it
        const int pre = ph == 3 ? 0 : (nkb < kStages ? nkb : kStages);        // GLU b reads the same operand as GLU a: no barrier
        for (int kb = 0; kb < pre; ++kb) {
          const uint32_t s = (it + kb) % kStages, round = (it + kb) / kStages;
          ptx::mbar_wait(&ctl->empty[s], (round & 1) ^ 1);
          if (leader) {
            ptx::mbar_arrive_expect_tx(&ctl->full[s], 16384 + wbytes);
            ptx::tma_load_2d(base + s * kStageBytes + kWOff, mw, &ctl->full[s], 64 * kb, wrow);
          }
        }
        if (ph != 3) ptx::mbar_wait(&ctl->go[ph == 4 ? 3 : ph], tile_it & 1);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % kStages, round = it / kStages;
          uint8_t* dst = base + s * kStageBytes;
          if (kb >= pre) {
            ptx::mbar_wait(&ctl->empty[s], (round & 1) ^ 1);
            if (leader) {
              ptx::mbar_arrive_expect_tx(&ctl->full[s], 16384 + wbytes);
              ptx::tma_load_2d(dst + kWOff, mw, &ctl->full[s], 64 * kb, wrow);
            }
          }
          if (leader) ptx::tma_load_2d(dst, ma, &ctl->full[s], 64 * kb, srow);
        }
      }
    }
  } else if (warp < 16) {
    const int ct = tid;
    const bool mma_warp = warp == 0;
    const bool leader = ptx::elect_one();
    const int q4 = warp & 3, sub = warp >> 2;
    const int row = q4 * 32 + lane;                          // tile row = TMEM lane
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const uint32_t tm = ctl->tmem_base;
    uint32_t mma_phase = 0, ring_it = 0, tile_it = 0;
    const float sc = 0.17677669529663688110f * 1.44269504088896340736f;
    int* const flags = p.flags + group * 4;

    auto wait_mma = [&]() {
      ptx::mbar_wait(&ctl->mma_bar, mma_phase);
      mma_phase ^= 1;
      ptx::tc_fence_after();
    };
    // one streamed GEMM: D[128 x N] at TMEM column `col` = A (ring) x W (ring)^T over nkb k-blocks of 64 (last one: K tail)
    auto stream_gemm = [&](uint32_t col, int N, int nkb, int ktail_steps) {
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t s = ring_it % kStages, round = ring_it / kStages;
        ptx::mbar_wait(&ctl->full[s], round & 1);
        ptx::tc_fence_after();
        const uint32_t st = sRing + s * kStageBytes;
        mma_kblock(leader, tm + col, st, st + kWOff, idesc(128, (uint32_t)N, 0), kb == nkb - 1 ? ktail_steps : 4, kb == 0);
        if (leader) ptx::umma_commit(&ctl->empty[s]);
        ++ring_it;
      }
    };
    // group barrier k: every CTA of the group has written its slice of the phase's output to the scratch area.  Slices
    // staged in shared memory (att, y: one SWIZZLE_128B k-block tile; hid: dense rows) leave through ONE TMA tensor store
    // issued by thread 0 here (store_map != nullptr), instead of 16-byte global stores with one row per lane.
    auto group_sync = [&](int k, const CUtensorMap* store_map, int c0) {
      if (store_map) ptx::fence_proxy_async();
      compute_sync();
      if (ct == 0) {
        if (store_map) {
          ptx::tma_store_2d(store_map, sR, c0, srow);
          ptx::bulk_commit();
          ptx::bulk_wait_all();
        }
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");     // generic-proxy stores before the TMA (async proxy) loads of the partners
        atomicAdd(flags + k, 1);
        const int target = G * (int)(tile_it + 1);
        int v;
        uint32_t spins = 0;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + k) : "memory");
          if (++spins > (1u << 26)) __trap();
        } while (v < target);
        asm volatile("fence.proxy.async;" ::: "memory");
        ptx::mbar_arrive(&ctl->go[k]);
      }
      compute_sync();
    };

    for (int t = group; t < n_tiles; t += n_groups, ++tile_it) {
      GL_STAMP(0);
      const int row0 = p.tile_list[2 * t];
      int rows = p.tile_list[2 * t + 1];
      const int split = rows >> 8;
      rows &= 255;
      if (ct < 128) {                                        // row table of the whole tile (attention, scatter-back)
        const int src = tile_src(ct, rows, split);
        int pix = 0, lo = 0, hi = 0;
        if (src >= 0) {
          pix = p.row_pix[row0 + src];
          const int wv = p.row_win[row0 + src];
          const int off = (wv >> 8) - row0;
          lo = (split && off) ? 64 : off;
          hi = lo + (wv & 255);
        }
        ctl->pix[ct] = pix; ctl->lo[ct] = (uint8_t)lo; ctl->hi[ct] = (uint8_t)hi;
      }
      // ---- LN1 / LN2 of this CTA's rows of the tile -> n2 (bf16 operand rows + fp32 shortcut rows) in the scratch area ----
      {
        const int rl = ct / LPT, l = ct % LPT;
        const int trow = slice * K::RPC + rl;
        const int src = tile_src(trow, rows, split);
        float4 v[4];
        if (src >= 0) {
          const float* xp = p.x + (long long)p.row_pix[row0 + src] * C;
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(xp + (l + LPT * i) * 4));
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ln_row<C, LPT>(v, l, spv + K::PV_LN, spv + K::PV_LN + C, p.eps);
        ln_row<C, LPT>(v, l, spv + K::PV_LN + 2 * C, spv + K::PV_LN + 3 * C, p.eps);
        __nv_bfloat16* hrow = p.n2h + (size_t)(srow + trow) * C;
        float* frow = p.n2f + (size_t)(srow + trow) * C;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (src < 0) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int c = (l + LPT * i) * 4;
          *reinterpret_cast<uint2*>(hrow + c) = make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
          *reinterpret_cast<float4*>(frow + c) = v[i];
        }
      }
      GL_STAMP(1);
      group_sync(0, nullptr, 0);
      GL_STAMP(2);

      // ---- QKV slice (heads 2s, 2s+1) ----
      if (mma_warp) {
        stream_gemm(0, 192, KB, 4);
        if (leader) ptx::umma_commit(&ctl->mma_bar);
      }
      float y[16];                                           // fp32 shortcut n2[row][64 s + 16 sub ..] (written by the row's owner CTA)
      {
        const float* nf = p.n2f + (size_t)(srow + row) * C + 64 * slice + 16 * sub;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 f = __ldcg(reinterpret_cast<const float4*>(nf + 4 * j));     // L2: the scratch is rewritten every tile
          y[4 * j] = f.x; y[4 * j + 1] = f.y; y[4 * j + 2] = f.z; y[4 * j + 3] = f.w;
        }
      }
      wait_mma();
      GL_STAMP(3);
      for (int u = sub; u < 6; u += 4) {                     // 32 accumulator columns = q, k or v of one head
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tm + lane_sel + (uint32_t)(u * 32), raw);
        ptx::tmem_ld_wait();
        const int h = u / 3, which = u - 3 * h;
        const uint32_t tile = which < 2 ? sR + (uint32_t)(h * 16384 + which * 8192) : sR + (uint32_t)(2 * 16384 + h * 8192);
        const uint32_t dst = tile + (uint32_t)(row * 64);
        const uint32_t sw = (uint32_t)((row >> 1) & 3);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 b0 = lds128(sPV + (uint32_t)(K::PV_QKVB + u * 32 + c * 8) * 4);
          const float4 b1 = lds128(sPV + (uint32_t)(K::PV_QKVB + u * 32 + c * 8 + 4) * 4);
          sts128(dst + (((uint32_t)c ^ sw) << 4),
                 pack_bf16(__uint_as_float(raw[8 * c]) + b0.x, __uint_as_float(raw[8 * c + 1]) + b0.y),
                 pack_bf16(__uint_as_float(raw[8 * c + 2]) + b0.z, __uint_as_float(raw[8 * c + 3]) + b0.w),
                 pack_bf16(__uint_as_float(raw[8 * c + 4]) + b1.x, __uint_as_float(raw[8 * c + 5]) + b1.y),
                 pack_bf16(__uint_as_float(raw[8 * c + 6]) + b1.z, __uint_as_float(raw[8 * c + 7]) + b1.w));
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      compute_sync();

      GL_STAMP(4);
      // ---- attention of the two heads (same code as the one-CTA kernel, two softmax threads per row) ----
      const int lo = ctl->lo[row], hi = ctl->hi[row];
      const bool rvalid = hi > lo;
      const int hh = sub >> 1, half = sub & 1;
      {
        if (mma_warp) {
          ptx::tc_fence_after();
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t qa = sR + (uint32_t)(e * 16384);
            const uint64_t dq = desc_sw64_k(qa), dk = desc_sw64_k(qa + 8192);
            if (leader) {
              ptx::umma_f16_ss(tm + 128 * e, dq, dk, idesc(128, 128, 0), 0u);
              ptx::umma_f16_ss(tm + 128 * e, dq + 2, dk + 2, idesc(128, 128, 0), 1u);
            }
          }
          if (leader) ptx::umma_commit(&ctl->mma_bar);
        }
        wait_mma();
        float mx = -INFINITY;
        bool need[2], full[2];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = half * 64 + cc * 32;
          need[cc] = __any_sync(kFull, lo < c0 + 32 && hi > c0);
          full[cc] = __all_sync(kFull, !rvalid || (lo <= c0 && c0 + 32 <= hi));
          if (need[cc]) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tm + lane_sel + (uint32_t)(128 * hh + c0), raw);
            ptx::tmem_ld_wait();
            if (full[cc]) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(raw[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j >= lo && c0 + j < hi) mx = fmaxf(mx, __uint_as_float(raw[j]));
            }
          }
        }
        ctl->pmax[hh][half][row] = mx;
        compute_sync();
        mx = fmaxf(ctl->pmax[hh][0][row], ctl->pmax[hh][1][row]);
        const float mxs = mx * sc;
        float rsum = 0.f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = half * 64 + cc * 32;
          uint32_t pk[16];
          if (need[cc]) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tm + lane_sel + (uint32_t)(128 * hh + c0), raw);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float p0 = ex2_approx(fmaf(__uint_as_float(raw[j]), sc, -mxs));
              float p1 = ex2_approx(fmaf(__uint_as_float(raw[j + 1]), sc, -mxs));
              if (!full[cc]) {
                p0 = (c0 + j >= lo && c0 + j < hi) ? p0 : 0.f;
                p1 = (c0 + j + 1 >= lo && c0 + j + 1 < hi) ? p1 : 0.f;
              }
              const uint32_t w = pack_bf16(p0, p1);
              pk[j >> 1] = w;
              rsum += __uint_as_float(w << 16) + __uint_as_float(w & 0xFFFF0000u);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
          ptx::tmem_st_32x16(tm + lane_sel + (uint32_t)(256 + 64 * hh + (c0 >> 1)), pk);
        }
        ctl->psum[hh][half][row] = rsum;
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        compute_sync();
        if (mma_warp) {
          ptx::tc_fence_after();
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint64_t dv = desc_sw64_mn(sR + (uint32_t)(2 * 16384 + e * 8192));
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              if (leader) ptx::umma_f16_ts(tm + 384 + 32 * e, tm + 256 + 64 * e + 8 * ks, dv + (uint64_t)(ks * 64), idesc(128, 32, 1), ks ? 1u : 0u);
          }
          if (leader) ptx::umma_commit(&ctl->mma_bar);
        }
        wait_mma();
        {
          uint32_t raw[16];
          ptx::tmem_ld_32x16(tm + lane_sel + (uint32_t)(384 + 32 * hh + 16 * half), raw);
          ptx::tmem_ld_wait();
          const float il = rvalid ? __fdividef(1.0f, ctl->psum[hh][0][row] + ctl->psum[hh][1][row]) : 0.f;
          // att slice [128 x 64] bf16 as one SWIZZLE_128B k-block tile in shared memory (Q,K,V are dead: P V has completed);
          // the group barrier's TMA store moves it to the scratch area (the A operand of every slice's proj GEMM)
          const uint32_t dst = sR + (uint32_t)(row * 128);
          const uint32_t ch = (uint32_t)(4 * hh + 2 * half), r7 = (uint32_t)(row & 7);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            sts128(dst + (((ch + c) ^ r7) << 4),
                   pack_bf16(__uint_as_float(raw[8 * c]) * il, __uint_as_float(raw[8 * c + 1]) * il),
                   pack_bf16(__uint_as_float(raw[8 * c + 2]) * il, __uint_as_float(raw[8 * c + 3]) * il),
                   pack_bf16(__uint_as_float(raw[8 * c + 4]) * il, __uint_as_float(raw[8 * c + 5]) * il),
                   pack_bf16(__uint_as_float(raw[8 * c + 6]) * il, __uint_as_float(raw[8 * c + 7]) * il));
        }
        ptx::tc_fence_before();
      }
      GL_STAMP(5);
      group_sync(1, &map_att, 64 * slice);
      GL_STAMP(6);

      // ---- proj slice + LayerScale + shortcut ----
      if (mma_warp) {
        ptx::tc_fence_after();
        stream_gemm(0, 64, KB, 4);
        if (leader) ptx::umma_commit(&ctl->mma_bar);
      }
      wait_mma();
      GL_STAMP(7);
      {
        const int col0 = sub * 16;
        uint32_t raw[16];
        ptx::tmem_ld_32x16(tm + lane_sel + (uint32_t)col0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = lds128(sPV + (uint32_t)(K::PV_PROJB + col0 + j) * 4);
          const float4 g4 = lds128(sPV + (uint32_t)(K::PV_G1 + col0 + j) * 4);
          y[j] = fmaf(g4.x, __uint_as_float(raw[j]) + b4.x, y[j]);
          y[j + 1] = fmaf(g4.y, __uint_as_float(raw[j + 1]) + b4.y, y[j + 1]);
          y[j + 2] = fmaf(g4.z, __uint_as_float(raw[j + 2]) + b4.z, y[j + 2]);
          y[j + 3] = fmaf(g4.w, __uint_as_float(raw[j + 3]) + b4.w, y[j + 3]);
        }
        const uint32_t dst = sR + (uint32_t)(row * 128);
        const uint32_t ch = (uint32_t)(col0 >> 3), r7 = (uint32_t)(row & 7);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          sts128(dst + (((ch + c) ^ r7) << 4), pack_bf16(y[8 * c], y[8 * c + 1]), pack_bf16(y[8 * c + 2], y[8 * c + 3]),
                 pack_bf16(y[8 * c + 4], y[8 * c + 5]), pack_bf16(y[8 * c + 6], y[8 * c + 7]));
      }
      ptx::tc_fence_before();
      GL_STAMP(8);
      group_sync(2, &map_y, 64 * slice);
      GL_STAMP(9);

      // ---- GLU slice: 336 accumulator columns in two passes over y ----
      if (mma_warp) {
        ptx::tc_fence_after();
        stream_gemm(128, kGluA, KB, 4);
        stream_gemm(128 + kGluA, kGluB, KB, 4);
        if (leader) ptx::umma_commit(&ctl->mma_bar);
      }
      wait_mma();
      GL_STAMP(10);
      for (int u = sub; u < 2 * IS / 16; u += 4) {            // 16 accumulator columns -> 8 hid columns = 16 bytes
        uint32_t raw[16];
        ptx::tmem_ld_32x16(tm + lane_sel + (uint32_t)(128 + 16 * u), raw);
        ptx::tmem_ld_wait();
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = lds128(sPV + (uint32_t)(K::PV_B1 + 16 * u + 4 * j) * 4);
          pk[j] = pack_bf16(glu_tanh_fit(0.5f * (__uint_as_float(raw[4 * j]) + b4.x), __uint_as_float(raw[4 * j + 1]) + b4.y),
                            glu_tanh_fit(0.5f * (__uint_as_float(raw[4 * j + 2]) + b4.z), __uint_as_float(raw[4 * j + 3]) + b4.w));
        }
        sts128(sR + (uint32_t)(row * (IS * 2) + 16 * u), pk[0], pk[1], pk[2], pk[3]);     // dense [128][IS] bf16: the TMA store's box
      }
      ptx::tc_fence_before();
      GL_STAMP(11);
      group_sync(3, &map_hid_st, IS * slice);
      GL_STAMP(12);

      // ---- MLP-out slice + LayerScale + residual + scatter-back ----
      if (mma_warp) {
        ptx::tc_fence_after();
        stream_gemm(0, 64, KBI, (I - (KBI - 1) * 64) / 16);
        if (leader) ptx::umma_commit(&ctl->mma_bar);
      }
      wait_mma();
      GL_STAMP(13);
      {
        const int col0 = sub * 16;
        uint32_t raw[16];
        ptx::tmem_ld_32x16(tm + lane_sel + (uint32_t)col0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = lds128(sPV + (uint32_t)(K::PV_B2 + col0 + j) * 4);
          const float4 g4 = lds128(sPV + (uint32_t)(K::PV_G2 + col0 + j) * 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sR + (uint32_t)(row * K::OUT_PITCH + (col0 + j) * 4)),
                       "f"(fmaf(g4.x, __uint_as_float(raw[j]) + b4.x, y[j])), "f"(fmaf(g4.y, __uint_as_float(raw[j + 1]) + b4.y, y[j + 1])),
                       "f"(fmaf(g4.z, __uint_as_float(raw[j + 2]) + b4.z, y[j + 2])), "f"(fmaf(g4.w, __uint_as_float(raw[j + 3]) + b4.w, y[j + 3]))
                       : "memory");
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      compute_sync();
      {
        const int orow = warp * 8 + lane;
        if (lane < 8 && ctl->hi[orow] > ctl->lo[orow]) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.out + (long long)ctl->pix[orow] * C + 64 * slice),
                       "r"(sR + (uint32_t)(orow * K::OUT_PITCH)), "r"(64 * 4) : "memory");
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the staging area is the next tile's Q,K,V tiles
      ptx::tc_fence_before();
      compute_sync();
      GL_STAMP(14);
    }
    SAST_STAMP(trc, tid == 0, 22);

    // ---- unselected tokens keep norm1(x): this CTA's share of the map, LPT lanes per token ----
    {
      constexpr int TOK = 512 / LPT;
      const int l = ct % LPT;
      const long long step = (long long)gridDim.x * TOK;
      const bool any_unselected = (long long)p.counts[1] < p.g.P;      // dense scene: every token selected, nothing to keep
      for (long long q = (long long)blockIdx.x * TOK + ct / LPT; any_unselected && q - ct / LPT < p.g.P; q += step) {
        const bool todo = q < p.g.P && p.tok_row[q] < 0;
        if (!__any_sync(kFull, todo)) continue;
        const long long pix = todo ? token_pixel(q, p.g, p.flavor) : 0;
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          v[i] = todo ? __ldg(reinterpret_cast<const float4*>(p.x + pix * C + (l + LPT * i) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        ln_row<C, LPT>(v, l, spv + K::PV_LN, spv + K::PV_LN + C, p.eps);
        if (todo) {
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(p.out + pix * C + (l + LPT * i) * 4) = v[i];
        }
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    SAST_STAMP(trc, tid == 0, 23);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(ctl->tmem_base, 512);
  }
}

template <int C>
static int launch_group_t(const sast_layer_args& a, const Geom& g, cudaStream_t st) {
  using K = Cfg<C>;
  const sast_layer_weights& w = a.w;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  const long long max_tiles = (g.P + 63) / 64;                   // a tile holds >= 1 window; worst case one window (> 64 rows) per tile
  int n_groups = (sms > 192 ? 192 : sms) / K::G;
  if (n_groups > max_tiles) n_groups = (int)max_tiles;
  if (n_groups < 1) n_groups = 1;
  // scratch: [n_groups * 128 rows] of n2 (bf16 + fp32), att, y (bf16, C wide) and hid (bf16, I wide), then the flags
  const size_t rows = (size_t)n_groups * 128;
  char* ws = (char*)a.workspace;
  auto take = [&](size_t bytes) { char* r = ws; ws += (bytes + 1023) / 1024 * 1024; return r; };
  Params p;
  p.n2h = (__nv_bfloat16*)take(rows * C * 2);
  p.att = (__nv_bfloat16*)take(rows * C * 2);
  p.yh = (__nv_bfloat16*)take(rows * C * 2);
  p.hid = (__nv_bfloat16*)take(rows * K::I * 2);
  p.n2f = (float*)take(rows * C * 4);
  p.flags = (int*)take((size_t)n_groups * 4 * sizeof(int));
  if ((size_t)(ws - (char*)a.workspace) > a.workspace_bytes) return SAST_E_WORKSPACE;
  cudaError_t e = cudaMemsetAsync(p.flags, 0, (size_t)n_groups * 4 * sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  CUtensorMap mn, ma, my, mh, mq, mp, m1, m2;
  int rc;
  if ((rc = make_tmap_bf16_box(&mn, p.n2h, (long long)rows, C, C, 64, 128, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&ma, p.att, (long long)rows, C, C, 64, 128, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&my, p.yh, (long long)rows, C, C, 64, 128, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&mh, p.hid, (long long)rows, K::I, K::I, 64, 128, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&mq, w.qkv_w_bf16, 3 * C, C, C, 64, 192, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&mp, w.proj_w_bf16, C, C, C, 64, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&m1, w.mlp1_w_bf16, 2 * K::I, C, C, 64, kGluA, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&m2, w.mlp2_w_bf16, C, K::I, K::I, 64, 64, 128))) return rc;
  CUtensorMap mhs;                                               // hid slice store: dense [128 x IS] box, no swizzle
  if ((rc = make_tmap_bf16_box(&mhs, p.hid, (long long)rows, K::I, K::I, K::IS, 128, 0))) return rc;
  p.x = a.x; p.out = a.out;
  p.ln1_w = w.ln1_w; p.ln1_b = w.ln1_b; p.ln2_w = w.ln2_w; p.ln2_b = w.ln2_b;
  p.qkv_b = w.qkv_b; p.proj_b = w.proj_b; p.gamma1 = w.gamma1; p.gamma2 = w.gamma2; p.mlp1_b = w.mlp1_b; p.mlp2_b = w.mlp2_b;
  p.eps = w.ln_eps;
  p.counts = a.sel.counts; p.tile_list = a.sel.tile_list; p.row_pix = a.sel.row_pix; p.row_win = a.sel.row_win;
  p.tok_row = a.sel.tok_row;
  p.g = g; p.flavor = a.flavor;
  p.trace = g_trace_which == 5 ? g_trace : nullptr;
  const size_t smem = (size_t)kStages * kStageBytes + K::R_BYTES + (sizeof(Ctl) + 15) / 16 * 16 + K::PV_FLOATS * 4;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e2 = cudaFuncSetAttribute(layer_group_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e2 != cudaSuccess) return (int)e2;
  }
  sast::launch_k(layer_group_kernel<C>, (unsigned)(n_groups * K::G), 544, smem, st, mn, ma, my, mh, mq, mp, m1, m2, mhs, p);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

}  // namespace gl

// scratch bytes of the group kernel for a map of P tokens (0: not a group-kernel shape)
size_t group_layer_workspace_bytes(long long P, int C, int I) {
  if (!((C == 256 && I == gl::Cfg<256>::I) || (C == 512 && I == gl::Cfg<512>::I))) return 0;
  const int G = C / 64;
  long long n_groups = 192 / G;                                  // groups a device of up to 192 SMs launches (B200: 148 / G)
  const long long max_tiles = (P + 63) / 64;
  if (n_groups > max_tiles) n_groups = max_tiles;
  if (n_groups < 1) n_groups = 1;
  const size_t rows = (size_t)n_groups * 128;
  auto up = [](size_t b) { return (b + 1023) / 1024 * 1024; };
  return 3 * up(rows * C * 2) + up(rows * (size_t)I * 2) + up(rows * C * 4) + up((size_t)n_groups * 16) + 1024;
}

static bool group_layer_enabled() {
  static int v = -1;     // read once; A/B knob: SAST_B200_GROUP=0 keeps the multi-kernel chain for C = 256 / 512
  if (v < 0) {
    const char* e = getenv("SAST_B200_GROUP");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// true if this layer can take the group kernel (bf16 path, dim_head 32, C 256 / 512 with the mlp_ratio-4 GLU width, no CB)
bool group_layer_supported(const sast_layer_args& a) {
  if (a.precision != SAST_BF16 || a.enable_cb || !group_layer_enabled()) return false;
  if (a.w.dim_head != 0 && a.w.dim_head != 32) return false;
  if (a.g.C == 256) return a.w.I == gl::Cfg<256>::I;
  if (a.g.C == 512) return a.w.I == gl::Cfg<512>::I;
  return false;
}

int launch_layer_group(const sast_layer_args& a, const Geom& g, cudaStream_t st) {
  if (!a.sel.tile_list || !a.sel.row_win || !a.workspace) return SAST_E_NULL;
  if (g.C == 256) return gl::launch_group_t<256>(a, g, st);
  if (g.C == 512) return gl::launch_group_t<512>(a, g, st);
  return SAST_E_UNSUPPORTED;
}

}  // namespace sast
