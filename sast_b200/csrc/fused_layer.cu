// One MS-WSA layer as ONE kernel (SAST_BF16 path, C = 64 or 128): replaces MS_WSA.forward, SAST.py:199-255.
//
// A persistent CTA (one per SM, 16 compute warps) walks the dense tile list of the selection (select.cu: a tile =
// whole selected windows of one frame, <= 128 compacted rows).  Per tile every intermediate stays on the SM:
//
//   gather     rows x[row_pix] (fp32 NHWC) -> LN1 -> LN2 -> bf16 A tile (SWIZZLE_128B K-major) + fp32 shortcut
//   QKV        tcgen05.mma  A x Wqkv^T  -> TMEM -> (+bias) bf16 Q,K,V operand tiles in shared memory (SWIZZLE_64B)
//   attention  per head pair: S = Q K^T (TMEM) -> block-diagonal softmax (a row sees the keys of its own window
//              only; the compacted tile holds no padding, so SAST.py:223-226's -1e4 mask has no counterpart)
//              -> P as bf16 pairs back into TMEM (tcgen05.st) -> O = P V with P as the TMEM A operand
//              (tcgen05.mma ts-form), row sums = P x ones on the tensor core -> O / rowsum -> bf16 A tile
//   proj       A x Wp^T -> y = n2 + g1 (o + b)   (fp32 in registers; bf16 copy -> A tile)
//   GLU        A x W1^T (value/gate rows interleaved) -> val * gelu(gate) -> bf16 hid tile in shared memory
//   MLP out    hid x W2^T -> out[row_pix] = y + g2 (m + b)   (scatter-back, SAST.py:248-254)
//
// and after its tiles the CTA takes its share of the unselected tokens: out = LN1(x) (they keep norm1(x),
// SAST.py:251-254).  HBM traffic per layer = read every token once + write every token once + the selection
// indices; qkv / att / y / hid never leave the SM (the 6-kernel chain of layer.cu moved ~7x that).
//
// Weights: C = 64 keeps all four matrices (96 KB of bf16 SWIZZLE_128B tiles) resident in shared memory for the
// life of the CTA; C = 128 (368 KB) streams them per tile through a 4 x 16 KB TMA ring fed by a 17th warp.
// Threads: warp w owns TMEM lanes 32 (w % 4) .. +31 (= tile rows) and column slice w / 4 of every accumulator.
// Warp 0 additionally issues every tcgen05.mma (whole warp on uniform values, one elected lane).
#include "layer.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace sast {

int make_tmap_bf16_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows,
                       int swizzle_bytes);

namespace fl {

constexpr int kComputeThreads = 512;
constexpr int kRingStages = 4, kRingBytes = 16384;

template <int C_>
struct Cfg {
  static constexpr int C = C_;
  static constexpr int H = C / 32;                 // heads (dim_head 32)
  static constexpr int I = C * 5 / 2;              // GLU width for mlp_ratio 4: floor(4C*2/3/32)*32 = 160 / 320
  static constexpr bool kRing = C > 64;
  static constexpr int kThreads = kComputeThreads + (kRing ? 32 : 0);
  static constexpr int KB_A = C / 64;              // 64-column k-blocks of the A tile
  static constexpr int A_BYTES = KB_A * 16384;
  static constexpr int R_BYTES = 128 * 3 * C * 2;  // Q,K,V operand tiles; also shortcut staging (before) and hid (after)
  static constexpr int KB_HID = (I + 63) / 64;
  static constexpr int W_QKV = 0;                  // resident layout (C = 64)
  static constexpr int W_PROJ = W_QKV + 3 * C * 128;
  static constexpr int W_1 = W_PROJ + C * 128;
  static constexpr int W_2 = W_1 + 2 * I * 128;
  static constexpr int W_BYTES = kRing ? kRingStages * kRingBytes : W_2 + KB_HID * C * 128;
  static constexpr int CPT = C / 4;                // accumulator columns per thread in the C-wide epilogues
  static constexpr int GLU_CAP = 384;              // TMEM columns of one GLU round
  static constexpr int GLU_ROUNDS = (2 * I + GLU_CAP - 1) / GLU_CAP;
  static constexpr int RING_CHUNKS = 3 * C / 64 + C / 64 + 2 * I / 64 + KB_HID;    // per tile (ring mode)
  static_assert(C == 64 || C == 128, "fused layer kernel: C = 64 or 128");
  static_assert(R_BYTES >= 128 * C * 4 && R_BYTES >= KB_HID * 16384, "R region too small");
};

// TMEM columns (512 allocated).  QKV accumulator [0,3C); during attention (QKV drained): S of the two heads of a
// pair, their P (bf16 pairs), O and row sums; then proj [0,C), GLU [128,512), MLP out [0,C).
constexpr uint32_t TM_QKV = 0, TM_S = 0, TM_P = 256, TM_O = 384, TM_RS = 448, TM_PROJ = 0, TM_GLU = 128, TM_OUT = 0;

struct Ctl {
  uint64_t mma_bar;
  uint64_t w_full[kRingStages];
  uint64_t w_empty[kRingStages];
  uint32_t tmem_base;
  int pix[128], lo[128], hi[128];
  float pmax[2][2][128];
};

struct Params {
  const float* x;
  float* out;
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *qkv_b, *proj_b, *gamma1, *gamma2, *mlp1_b, *mlp2_b;
  float eps;
  const int *counts, *tile_list, *row_tok, *row_pix, *win_row0, *tok_row;
  Geom g;
  int flavor;
  long long* trace;      // debug stamps (sast_debug_trace which = 4), normally null; trace build only
};

__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ float group4_sum(float v) {
  v += __shfl_xor_sync(kFull, v, 1);
  v += __shfl_xor_sync(kFull, v, 2);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// K-major SWIZZLE_64B operand (rows of 32 bf16, 8-row atoms 512 bytes apart): Q and K tiles
__device__ __forceinline__ uint64_t desc_sw64_k(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}
// MN-major SWIZZLE_64B operand (one key per 64-byte row of 32 output dims): V tiles as the B operand of P V
__device__ __forceinline__ uint64_t desc_sw64_mn(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__host__ __device__ constexpr uint32_t idesc(uint32_t M, uint32_t N, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// one 64-column k-block (nks <= 4 k-steps of 16) of D[128 x N] (+)= A B^T, both operands SWIZZLE_128B K-major
__device__ __forceinline__ void mma_kblock(bool leader, uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t id,
                                           int nks, bool fresh) {
  const uint64_t da = ptx::umma_desc_sw128_kmajor(a_addr), db = ptx::umma_desc_sw128_kmajor(b_addr);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nks && leader) ptx::umma_f16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), id, (k > 0 || !fresh) ? 1u : 0u);
}

// trace build only: [CTA][32] clock64 stamps of thread 0: 0..14 phase boundaries of the CTA's SECOND tile (steady state),
// 15 kernel entry, 16 set-up done, 17 tiles done, 18 unselected pass done, 19 SM id, 20 tiles of this CTA
#define FL_STAMP(i) SAST_STAMP(trc, tid == 0 && ti == 1, (i))

template <int C>
__global__ void __launch_bounds__(Cfg<C>::kThreads, 1)
layer_fused_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_proj,
                   const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const Params p) {
  using K = Cfg<C>;
  constexpr int H = K::H, I = K::I, CPT = K::CPT;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sW = ptx::smem_u32(base);
  const uint32_t sA = sW + K::W_BYTES;
  const uint32_t sR = sA + K::A_BYTES;
  const uint32_t sOnes = sR + K::R_BYTES;
  Ctl* ctl = reinterpret_cast<Ctl*>(base + K::W_BYTES + K::A_BYTES + K::R_BYTES + 1024);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(kFull, tid >> 5, 0), lane = tid & 31;
  [[maybe_unused]] long long* const trc = p.trace ? p.trace + (size_t)blockIdx.x * 32 : nullptr;
  SAST_STAMP(trc, tid == 0, 15);

  if (tid == 0) {
    ptx::tma_prefetch_desc(&map_qkv); ptx::tma_prefetch_desc(&map_proj);
    ptx::tma_prefetch_desc(&map_w1); ptx::tma_prefetch_desc(&map_w2);
    ptx::mbar_init(&ctl->mma_bar, 1);
    for (int s = 0; s < kRingStages; ++s) { ptx::mbar_init(&ctl->w_full[s], 1); ptx::mbar_init(&ctl->w_empty[s], 1); }
    ptx::fence_barrier_init();
    if (!K::kRing) {
      // the weights never change inside a forward: load them before the PDL wait.  One box per MMA operand tile.
      ptx::mbar_arrive_expect_tx(&ctl->w_full[0], (uint32_t)K::W_BYTES);
      ptx::tma_load_2d(base + K::W_QKV, &map_qkv, &ctl->w_full[0], 0, 0);                    // [3C rows x 64]
      ptx::tma_load_2d(base + K::W_PROJ, &map_proj, &ctl->w_full[0], 0, 0);                  // [C x 64]
      ptx::tma_load_2d(base + K::W_1, &map_w1, &ctl->w_full[0], 0, 0);                       // [I x 64] rows 0..I-1
      ptx::tma_load_2d(base + K::W_1 + I * 128, &map_w1, &ctl->w_full[0], 0, I);             //          rows I..2I-1
      for (int kb = 0; kb < K::KB_HID; ++kb)                                                  // [C x 64] per k-block (zero-filled K tail)
        ptx::tma_load_2d(base + K::W_2 + kb * C * 128, &map_w2, &ctl->w_full[0], kb * 64, 0);
    }
  }
  if (warp == 0) ptx::tmem_alloc(&ctl->tmem_base, 512);
  if (tid < 256) reinterpret_cast<uint32_t*>(base + K::W_BYTES + K::A_BYTES + K::R_BYTES)[tid] = 0x3F803F80u;   // bf16 ones
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = ctl->tmem_base;

  pdl_entry();                                             // selection + input map are read from here on
  const int n_tiles = p.counts[3];
  SAST_STAMP(trc, tid == 0, 16);

  if (K::kRing && warp == 16) {
    // ---------------- weight producer (ring mode): 23 chunks of 16 KB per tile, in consumption order ----------------
    const bool leader = ptx::elect_one();
    uint32_t it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int c = 0; c < K::RING_CHUNKS; ++c, ++it) {
        const uint32_t s = it % kRingStages, round = it / kRingStages;
        ptx::mbar_wait(&ctl->w_empty[s], (round & 1) ^ 1);
        if (leader) {
          uint8_t* dst = base + s * kRingBytes;
          ptx::mbar_arrive_expect_tx(&ctl->w_full[s], kRingBytes);
          constexpr int nq = 3 * C / 64, np = C / 64, n1 = 2 * I / 64;
          if (c < nq + np + n1) {                       // 64 weight rows x K = 128: two [64 x 64] boxes
            const CUtensorMap* m = c < nq ? &map_qkv : c < nq + np ? &map_proj : &map_w1;
            const int row = 64 * (c < nq ? c : c < nq + np ? c - nq : c - nq - np);
            ptx::tma_load_2d(dst, m, &ctl->w_full[s], 0, row);
            ptx::tma_load_2d(dst + 8192, m, &ctl->w_full[s], 64, row);
          } else {                                      // W2: all C rows x one 64-column k-block
            ptx::tma_load_2d(dst, &map_w2, &ctl->w_full[s], 64 * (c - nq - np - n1), 0);
          }
        }
      }
    }
  } else if (warp < 16) {
    const bool mma_warp = warp == 0;
    const bool leader = ptx::elect_one();
    const int q4 = warp & 3, sub = warp >> 2;
    const int row = q4 * 32 + lane;                        // tile row = TMEM lane
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const uint32_t r7 = (uint32_t)(row & 7);
    uint32_t mma_phase = 0, ring_it = 0;
    const float sc = 0.17677669529663688110f * 1.44269504088896340736f;   // 32^-0.5 * log2(e)

    auto wait_mma = [&]() {
      ptx::mbar_wait(&ctl->mma_bar, mma_phase);
      mma_phase ^= 1;
      ptx::tc_fence_after();
    };
    // ring consumer helpers (warp 0 only)
    auto ring_acquire = [&]() -> uint32_t {
      const uint32_t s = ring_it % kRingStages, round = ring_it / kRingStages;
      ptx::mbar_wait(&ctl->w_full[s], round & 1);
      ptx::tc_fence_after();
      return sW + s * kRingBytes;
    };
    auto ring_release = [&]() {
      if (leader) ptx::umma_commit(&ctl->w_empty[ring_it % kRingStages]);
      ++ring_it;
    };

    if (!K::kRing && mma_warp) ptx::mbar_wait(&ctl->w_full[0], 0);     // resident weights have landed

    [[maybe_unused]] int ti = -1;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      ++ti;
      FL_STAMP(0);
      const int row0 = p.tile_list[2 * t], rows = p.tile_list[2 * t + 1];

      // ---- tile bookkeeping (consumed after later barriers) + gather / LN1 / LN2 -------------------------------
      if (tid < 128) {
        int pix = 0, lo = 0, hi = 0;
        if (tid < rows) {
          pix = p.row_pix[row0 + tid];
          const int w = p.row_tok[row0 + tid] / p.g.T;
          lo = p.win_row0[w] - row0;
          hi = p.win_row0[w + 1] - row0;
        }
        ctl->pix[tid] = pix; ctl->lo[tid] = lo; ctl->hi[tid] = hi;
      }
      {
        constexpr int NV = C / 16;                           // float4 per lane, 4 lanes per row
        const int r = tid >> 2, l = tid & 3;
        const bool valid = r < rows;
        float4 v[NV];
        {
          const long long pix = valid ? p.row_pix[row0 + r] : 0;
          const float* xp = p.x + pix * C;
#pragma unroll
          for (int i = 0; i < NV; ++i)
            v[i] = valid ? __ldg(reinterpret_cast<const float4*>(xp + (l + 4 * i) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float inv_c = 1.0f / (float)C;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {               // LN1 then LN2 (SAST.py:206, :213)
          const float* gw = pass == 0 ? p.ln1_w : p.ln2_w;
          const float* gb = pass == 0 ? p.ln1_b : p.ln2_b;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
          const float mean = group4_sum(s) * inv_c;
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
          }
          const float rstd = rsqrtf(group4_sum(ss) * inv_c + p.eps);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(gw + (l + 4 * i) * 4));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(gb + (l + 4 * i) * 4));
            v[i].x = (v[i].x - mean) * rstd * w4.x + b4.x; v[i].y = (v[i].y - mean) * rstd * w4.y + b4.y;
            v[i].z = (v[i].z - mean) * rstd * w4.z + b4.z; v[i].w = (v[i].w - mean) * rstd * w4.w + b4.w;
          }
        }
        const uint32_t rr7 = (uint32_t)(r & 7);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (!valid) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);    // rows past the tile: finite operands
          const int c = (l + 4 * i) * 4;
          const uint32_t chunk = (uint32_t)((c & 63) >> 3);
          sts64(sA + (uint32_t)((c >> 6) * 16384 + r * 128) + ((chunk ^ rr7) << 4) + (uint32_t)((c & 7) * 2),
                pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
          const uint32_t ch4 = (uint32_t)(l + 4 * i);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sR + (uint32_t)(r * C * 4) + ((ch4 ^ rr7) << 4)),
                       "f"(v[i].x), "f"(v[i].y), "f"(v[i].z), "f"(v[i].w) : "memory");
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      compute_sync();
      FL_STAMP(1);

      // ---- QKV = n2 Wqkv^T ----------------------------------------------------------------------------------------
      if (mma_warp) {
        ptx::tc_fence_after();
        if (!K::kRing) {
          mma_kblock(leader, tm + TM_QKV, sA, sW + K::W_QKV, idesc(128, 3 * C, 0), 4, true);
        } else {
          for (int j = 0; j < 3 * C / 64; ++j) {
            const uint32_t st = ring_acquire();
            for (int kb = 0; kb < K::KB_A; ++kb)
              mma_kblock(leader, tm + TM_QKV + 64 * j, sA + kb * 16384, st + kb * 8192, idesc(128, 64, 0), 4, kb == 0);
            ring_release();
          }
        }
        if (leader) ptx::umma_commit(&ctl->mma_bar);
      }
      // shortcut n2 (fp32) from the staging area into registers, in the epilogue mapping (row, column slice `sub`)
      float y[CPT];
#pragma unroll
      for (int j = 0; j < CPT / 4; ++j) {
        const uint32_t ch4 = (uint32_t)(sub * (CPT / 4) + j);
        const float4 f = lds128(sR + (uint32_t)(row * C * 4) + ((ch4 ^ r7) << 4));
        y[4 * j] = f.x; y[4 * j + 1] = f.y; y[4 * j + 2] = f.z; y[4 * j + 3] = f.w;
      }
      wait_mma();
      FL_STAMP(2);
      compute_sync();                                        // every shortcut read is done: the Q,K,V tiles may overwrite it

      // ---- QKV epilogue: + bias, bf16, operand tiles of the attention MMAs ------------------------------------
      for (int u = sub; u < 3 * H; u += 4) {                 // 32 accumulator columns = q, k or v of one head
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tm + lane_sel + TM_QKV + (uint32_t)(u * 32), raw);
        ptx::tmem_ld_wait();
        const int h = u / 3, which = u - 3 * h;
        const uint32_t tile = which < 2 ? sR + (uint32_t)(h * 16384 + which * 8192) : sR + (uint32_t)(H * 16384 + h * 8192);
        const uint32_t dst = tile + (uint32_t)(row * 64);
        const uint32_t sw = (uint32_t)((row >> 1) & 3);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (p.qkv_b) {
            b0 = __ldg(reinterpret_cast<const float4*>(p.qkv_b + u * 32 + c * 8));
            b1 = __ldg(reinterpret_cast<const float4*>(p.qkv_b + u * 32 + c * 8 + 4));
          }
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
      FL_STAMP(3);

      // ---- attention, two heads at a time -------------------------------------------------------------------------
      const int lo = ctl->lo[row], hi = ctl->hi[row];
      const bool rvalid = row < rows;
      const int hh = sub >> 1, half = sub & 1;               // softmax / O mapping: head of the pair, half of the columns
      for (int hp = 0; hp < H / 2; ++hp) {
        if (mma_warp) {
          ptx::tc_fence_after();
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t qa = sR + (uint32_t)((2 * hp + e) * 16384);
            const uint64_t dq = desc_sw64_k(qa), dk = desc_sw64_k(qa + 8192);
            if (leader) {
              ptx::umma_f16_ss(tm + TM_S + 128 * e, dq, dk, idesc(128, 128, 0), 0u);
              ptx::umma_f16_ss(tm + TM_S + 128 * e, dq + 2, dk + 2, idesc(128, 128, 0), 1u);     // dims 16..31
            }
          }
          if (leader) ptx::umma_commit(&ctl->mma_bar);
        }
        wait_mma();
        FL_STAMP(4);
        // pass 1: row maximum over the keys of the row's own window, this thread's 64 columns
        float mx = -INFINITY;
        bool need[2];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = half * 64 + cc * 32;
          need[cc] = __any_sync(kFull, lo < c0 + 32 && hi > c0);
          if (need[cc]) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tm + lane_sel + TM_S + (uint32_t)(128 * hh + c0), raw);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j >= lo && c0 + j < hi) mx = fmaxf(mx, __uint_as_float(raw[j]));
          }
        }
        ctl->pmax[hh][half][row] = mx;
        compute_sync();
        FL_STAMP(5);
        mx = fmaxf(ctl->pmax[hh][0][row], ctl->pmax[hh][1][row]);
        const float mxs = mx * sc;
        // pass 2: p = 2^((s - max) scale log2e) as bf16 pairs into TMEM -- the A operand of P V
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = half * 64 + cc * 32;
          uint32_t pk[16];
          if (need[cc]) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tm + lane_sel + TM_S + (uint32_t)(128 * hh + c0), raw);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float p0 = ex2_approx(fmaf(__uint_as_float(raw[j]), sc, -mxs));
              float p1 = ex2_approx(fmaf(__uint_as_float(raw[j + 1]), sc, -mxs));
              p0 = (c0 + j >= lo && c0 + j < hi) ? p0 : 0.f;
              p1 = (c0 + j + 1 >= lo && c0 + j + 1 < hi) ? p1 : 0.f;
              pk[j >> 1] = pack_bf16(p0, p1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
          ptx::tmem_st_32x16(tm + lane_sel + TM_P + (uint32_t)(64 * hh + (c0 >> 1)), pk);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        compute_sync();
        FL_STAMP(6);
        if (mma_warp) {
          ptx::tc_fence_after();
          const uint64_t d1 = desc_sw64_k(sOnes);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint64_t dv = desc_sw64_mn(sR + (uint32_t)(H * 16384 + (2 * hp + e) * 8192));
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              if (leader) {
                // P: one k-step = 16 keys = 8 TMEM columns; V: next 16 keys = +1 KB
                ptx::umma_f16_ts(tm + TM_O + 32 * e, tm + TM_P + 64 * e + 8 * ks, dv + (uint64_t)(ks * 64), idesc(128, 32, 1), ks ? 1u : 0u);
                ptx::umma_f16_ts(tm + TM_RS + 16 * e, tm + TM_P + 64 * e + 8 * ks, d1, idesc(128, 16, 0), ks ? 1u : 0u);
              }
            }
          }
          if (leader) ptx::umma_commit(&ctl->mma_bar);
        }
        wait_mma();
        FL_STAMP(7);
        {
          uint32_t raw[16];
          ptx::tmem_ld_32x16(tm + lane_sel + TM_O + (uint32_t)(32 * hh + 16 * half), raw);
          const uint32_t rs = ptx::tmem_ld_32x1(tm + lane_sel + TM_RS + (uint32_t)(16 * hh));
          ptx::tmem_ld_wait();
          const float il = rvalid ? __fdividef(1.0f, __uint_as_float(rs)) : 0.f;
          const int col = (2 * hp + hh) * 32 + half * 16;
          const uint32_t dst = sA + (uint32_t)((col >> 6) * 16384 + row * 128);
          const uint32_t ch = (uint32_t)((col & 63) >> 3);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            sts128(dst + (((ch + c) ^ r7) << 4),
                   pack_bf16(__uint_as_float(raw[8 * c]) * il, __uint_as_float(raw[8 * c + 1]) * il),
                   pack_bf16(__uint_as_float(raw[8 * c + 2]) * il, __uint_as_float(raw[8 * c + 3]) * il),
                   pack_bf16(__uint_as_float(raw[8 * c + 4]) * il, __uint_as_float(raw[8 * c + 5]) * il),
                   pack_bf16(__uint_as_float(raw[8 * c + 6]) * il, __uint_as_float(raw[8 * c + 7]) * il));
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        compute_sync();
        FL_STAMP(8);
      }

      // ---- proj + LayerScale + shortcut: y = n2 + g1 (o Wp^T + b) ---------------------------------------------
      if (mma_warp) {
        ptx::tc_fence_after();
        if (!K::kRing) {
          mma_kblock(leader, tm + TM_PROJ, sA, sW + K::W_PROJ, idesc(128, C, 0), 4, true);
        } else {
          for (int j = 0; j < C / 64; ++j) {
            const uint32_t st = ring_acquire();
            for (int kb = 0; kb < K::KB_A; ++kb)
              mma_kblock(leader, tm + TM_PROJ + 64 * j, sA + kb * 16384, st + kb * 8192, idesc(128, 64, 0), 4, kb == 0);
            ring_release();
          }
        }
        if (leader) ptx::umma_commit(&ctl->mma_bar);
      }
      wait_mma();
      FL_STAMP(9);
      {
        const int col0 = sub * CPT;
        uint32_t raw[CPT];
        if constexpr (CPT == 16) ptx::tmem_ld_32x16(tm + lane_sel + TM_PROJ + (uint32_t)col0, raw);
        else ptx::tmem_ld_32x32(tm + lane_sel + TM_PROJ + (uint32_t)col0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.proj_b) b4 = __ldg(reinterpret_cast<const float4*>(p.proj_b + col0 + j));
          if (p.gamma1) g4 = __ldg(reinterpret_cast<const float4*>(p.gamma1 + col0 + j));
          y[j] = fmaf(g4.x, __uint_as_float(raw[j]) + b4.x, y[j]);
          y[j + 1] = fmaf(g4.y, __uint_as_float(raw[j + 1]) + b4.y, y[j + 1]);
          y[j + 2] = fmaf(g4.z, __uint_as_float(raw[j + 2]) + b4.z, y[j + 2]);
          y[j + 3] = fmaf(g4.w, __uint_as_float(raw[j + 3]) + b4.w, y[j + 3]);
        }
        const uint32_t dst = sA + (uint32_t)((col0 >> 6) * 16384 + row * 128);
        const uint32_t ch = (uint32_t)((col0 & 63) >> 3);
#pragma unroll
        for (int c = 0; c < CPT / 8; ++c)
          sts128(dst + (((ch + c) ^ r7) << 4), pack_bf16(y[8 * c], y[8 * c + 1]), pack_bf16(y[8 * c + 2], y[8 * c + 3]),
                 pack_bf16(y[8 * c + 4], y[8 * c + 5]), pack_bf16(y[8 * c + 6], y[8 * c + 7]));
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      compute_sync();
      FL_STAMP(10);

      // ---- GLU: hid = val * gelu(gate) of y W1^T + b  (ops.py:135-137; weight rows interleaved value_j, gate_j) ----
      for (int rd = 0; rd < K::GLU_ROUNDS; ++rd) {
        const int ncols = min(2 * I - rd * K::GLU_CAP, K::GLU_CAP);     // accumulator columns of this round
        if (mma_warp) {
          ptx::tc_fence_after();
          if (!K::kRing) {
            mma_kblock(leader, tm + TM_GLU, sA, sW + K::W_1, idesc(128, I, 0), 4, true);
            mma_kblock(leader, tm + TM_GLU + I, sA, sW + K::W_1 + I * 128, idesc(128, I, 0), 4, true);
          } else {
            for (int j = 0; j < ncols / 64; ++j) {
              const uint32_t st = ring_acquire();
              for (int kb = 0; kb < K::KB_A; ++kb)
                mma_kblock(leader, tm + TM_GLU + 64 * j, sA + kb * 16384, st + kb * 8192, idesc(128, 64, 0), 4, kb == 0);
              ring_release();
            }
          }
          if (leader) ptx::umma_commit(&ctl->mma_bar);
        }
        wait_mma();
        FL_STAMP(11);
        for (int u = sub; u < ncols / 16; u += 4) {            // 16 accumulator columns -> 8 hid columns = one 16-byte chunk
          uint32_t raw[16];
          ptx::tmem_ld_32x16(tm + lane_sel + TM_GLU + (uint32_t)(16 * u), raw);
          ptx::tmem_ld_wait();
          const int ac = rd * K::GLU_CAP + 16 * u;             // global accumulator column (= interleaved bias index)
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.mlp1_b) b4 = __ldg(reinterpret_cast<const float4*>(p.mlp1_b + ac + 4 * j));
            pk[j] = pack_bf16(glu_tanh_fit(0.5f * (__uint_as_float(raw[4 * j]) + b4.x), __uint_as_float(raw[4 * j + 1]) + b4.y),
                              glu_tanh_fit(0.5f * (__uint_as_float(raw[4 * j + 2]) + b4.z), __uint_as_float(raw[4 * j + 3]) + b4.w));
          }
          const int hc = ac >> 1;
          sts128(sR + (uint32_t)((hc >> 6) * 16384 + row * 128) + ((((uint32_t)(hc & 63) >> 3) ^ r7) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        compute_sync();
        FL_STAMP(12);
      }

      // ---- MLP out + LayerScale + residual + scatter-back: out[pix] = y + g2 (hid W2^T + b) ---------------------
      if (mma_warp) {
        ptx::tc_fence_after();
        for (int kb = 0; kb < K::KB_HID; ++kb) {
          const int nks = min(4, (I - kb * 64) / 16);
          if (!K::kRing) {
            mma_kblock(leader, tm + TM_OUT, sR + kb * 16384, sW + K::W_2 + kb * C * 128, idesc(128, C, 0), nks, kb == 0);
          } else {
            const uint32_t st = ring_acquire();
            mma_kblock(leader, tm + TM_OUT, sR + kb * 16384, st, idesc(128, C, 0), nks, kb == 0);
            ring_release();
          }
        }
        if (leader) ptx::umma_commit(&ctl->mma_bar);
      }
      wait_mma();
      FL_STAMP(13);
      {
        const int col0 = sub * CPT;
        uint32_t raw[CPT];
        if constexpr (CPT == 16) ptx::tmem_ld_32x16(tm + lane_sel + TM_OUT + (uint32_t)col0, raw);
        else ptx::tmem_ld_32x32(tm + lane_sel + TM_OUT + (uint32_t)col0, raw);
        ptx::tmem_ld_wait();
        float* op = p.out + (long long)ctl->pix[row] * C + col0;
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.mlp2_b) b4 = __ldg(reinterpret_cast<const float4*>(p.mlp2_b + col0 + j));
          if (p.gamma2) g4 = __ldg(reinterpret_cast<const float4*>(p.gamma2 + col0 + j));
          const float4 o = make_float4(fmaf(g4.x, __uint_as_float(raw[j]) + b4.x, y[j]), fmaf(g4.y, __uint_as_float(raw[j + 1]) + b4.y, y[j + 1]),
                                       fmaf(g4.z, __uint_as_float(raw[j + 2]) + b4.z, y[j + 2]), fmaf(g4.w, __uint_as_float(raw[j + 3]) + b4.w, y[j + 3]));
          if (rvalid) *reinterpret_cast<float4*>(op + j) = o;
        }
      }
      ptx::tc_fence_before();
      compute_sync();                                        // TMEM, the A tile and ctl->pix/lo/hi are free for the next tile
      FL_STAMP(14);
    }
    SAST_STAMP(trc, tid == 0, 17);
#ifdef SAST_TRACE
    if (trc && tid == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); trc[19] = smid; trc[20] = ti + 1; }
#endif

    // ---- unselected tokens keep norm1(x)  (SAST.py:251-254): this CTA's share, 4 lanes per token ------------------
    {
      constexpr int NV = C / 16;
      const int l = tid & 3;
      const float inv_c = 1.0f / (float)C;
      for (long long q0 = (long long)blockIdx.x * 128; q0 < p.g.P; q0 += (long long)gridDim.x * 128) {
        const long long q = q0 + (tid >> 2);
        const bool todo = q < p.g.P && p.tok_row[q] < 0;
        const long long pix = todo ? token_pixel(q, p.g, p.flavor) : 0;
        float4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
          v[i] = todo ? __ldg(reinterpret_cast<const float4*>(p.x + pix * C + (l + 4 * i) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (!__any_sync(kFull, todo)) continue;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float mean = group4_sum(s) * inv_c;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
          ss += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(group4_sum(ss) * inv_c + p.eps);
        if (todo) {
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.ln1_w + (l + 4 * i) * 4));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln1_b + (l + 4 * i) * 4));
            *reinterpret_cast<float4*>(p.out + pix * C + (l + 4 * i) * 4) =
                make_float4((v[i].x - mean) * rstd * w4.x + b4.x, (v[i].y - mean) * rstd * w4.y + b4.y,
                            (v[i].z - mean) * rstd * w4.z + b4.z, (v[i].w - mean) * rstd * w4.w + b4.w);
          }
        }
      }
    }
    SAST_STAMP(trc, tid == 0, 18);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tm, 512);
  }
}

template <int C>
static int launch_fused_t(const sast_layer_args& a, const Geom& g, cudaStream_t st) {
  using K = Cfg<C>;
  const sast_layer_weights& w = a.w;
  CUtensorMap mq, mp, m1, m2;
  int rc;
  // box rows = the weight rows one TMA box (= one MMA B tile or ring chunk) holds
  if ((rc = make_tmap_bf16_box(&mq, w.qkv_w_bf16, 3 * C, C, C, 64, K::kRing ? 64 : 3 * C, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&mp, w.proj_w_bf16, C, C, C, 64, K::kRing ? 64 : C, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&m1, w.mlp1_w_bf16, 2 * K::I, C, C, 64, K::kRing ? 64 : K::I, 128))) return rc;
  if ((rc = make_tmap_bf16_box(&m2, w.mlp2_w_bf16, C, K::I, K::I, 64, C, 128))) return rc;
  Params p;
  p.x = a.x; p.out = a.out;
  p.ln1_w = w.ln1_w; p.ln1_b = w.ln1_b; p.ln2_w = w.ln2_w; p.ln2_b = w.ln2_b;
  p.qkv_b = w.qkv_b; p.proj_b = w.proj_b; p.gamma1 = w.gamma1; p.gamma2 = w.gamma2; p.mlp1_b = w.mlp1_b; p.mlp2_b = w.mlp2_b;
  p.eps = w.ln_eps;
  p.counts = a.sel.counts; p.tile_list = a.sel.tile_list; p.row_tok = a.sel.row_tok; p.row_pix = a.sel.row_pix;
  p.win_row0 = a.sel.win_row0; p.tok_row = a.sel.tok_row;
  p.g = g; p.flavor = a.flavor;
  p.trace = g_trace_which == 4 ? g_trace : nullptr;
  const size_t smem = 1024 + (size_t)K::W_BYTES + K::A_BYTES + K::R_BYTES + 1024 + sizeof(Ctl);
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(layer_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  const long long chunks = (g.P + 127) / 128;
  const unsigned grid = (unsigned)(chunks < sms ? chunks : sms);
  sast::launch_k(layer_fused_kernel<C>, grid, K::kThreads, smem, st, mq, mp, m1, m2, p);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

}  // namespace fl

bool fused_layer_enabled() {
  static int v = -1;     // read once; A/B knob: SAST_B200_FUSED=0 selects the 6-kernel chain
  if (v < 0) {
    const char* e = getenv("SAST_B200_FUSED");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// true if this layer can take the fused kernel (bf16 path, C 64 / 128 with the mlp_ratio-4 GLU width, no context broadcast)
bool fused_layer_supported(const sast_layer_args& a) {
  if (a.precision != SAST_BF16 || a.enable_cb || !fused_layer_enabled()) return false;
  if (a.g.C == 64) return a.w.I == fl::Cfg<64>::I;
  if (a.g.C == 128) return a.w.I == fl::Cfg<128>::I;
  return false;
}

int launch_layer_fused(const sast_layer_args& a, const Geom& g, cudaStream_t st) {
  if (!a.sel.tile_list) return SAST_E_NULL;
  if (g.C == 64) return fl::launch_fused_t<64>(a, g, st);
  if (g.C == 128) return fl::launch_fused_t<128>(a, g, st);
  return SAST_E_UNSUPPORTED;
}

}  // namespace sast
