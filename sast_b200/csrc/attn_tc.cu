// Attention over the compacted rows on the 5th-gen tensor cores (SAST_BF16 path):
// per attention tile (consecutive selected windows of one frame, <= 128 compacted rows, packed
// by select_scan_kernel) and head:  S = Q K^T (tcgen05, fp32 in TMEM) -> block-diagonal softmax
// (a row only sees the keys of its own window) -> P (bf16, shared memory) -> O = P V and the row sums
// P 1 (tcgen05) -> O / rowsum -> bf16.                                  (replaces SAST.py:219-229)
//
// The compacted buffer holds selected tokens only, so the reference's -1e4 column mask for
// padding (SAST.py:223-226) has no counterpart: the only mask is "same window".
//
// CTA = 256 threads, two threads per tile row (= TMEM lane), each on half of the key columns.  Warp 0 issues TMA
// and MMA through an elected lane and is the only one polling mbarriers; the others block in bar.sync.
// 128 TMEM columns and 44 KB of shared memory per CTA -> 4 CTAs per SM hide each other's
// load -> MMA -> softmax -> MMA latency chain.  The grid is tile-major over the per-frame tile slot lists
// (blockIdx.x = slot * B + frame), so the CTAs of unused slots are launched last and exit at once.
//   Q,K,V tiles [128 x 32] bf16: TMA boxes out of the qkv buffer ([rows, 3C], head-major
//   [h][q,k,v][32]) in SWIZZLE_64B; Q,K are K-major operands, V is the MN-major B operand of PV.
//   P [128 x 128] bf16 is written by the softmax threads in the SWIZZLE_128B K-major layout.
//   A 1 KB tile of bf16 ones is the B operand of a second product into TMEM columns 32..47: the row sums of
//   the bf16-rounded P, for free (the softmax loop is instruction bound).
#include "layer.cuh"
#include "ptx.cuh"

namespace sast {

constexpr uint32_t AT_TMEM_COLS = 128;     // S: columns [0,128); O re-uses [0,32) and the row sums [32,48) once S has been consumed
constexpr int AT_TILE = 8192;              // one 128 x 32 bf16 operand tile

struct AttnSmem {
  uint64_t bar_load, bar_s, bar_o;
  uint32_t tmem_base;
};

// K-major, SWIZZLE_64B: rows of 64 bytes (32 bf16), 8-row atoms 512 bytes apart
__device__ __forceinline__ uint64_t desc_sw64_kmajor(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;                  // SWIZZLE_64B
  return d;
}
// MN-major, SWIZZLE_64B: one K index (key) per 64-byte row of 32 MN elements, 8-key atoms 512 bytes apart
__device__ __forceinline__ uint64_t desc_sw64_mnmajor(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(512 >> 4) << 16;         // LBO: next 32-wide MN block (unused, N = 32)
  d |= (uint64_t)(512 >> 4) << 32;         // SBO: next 8 keys
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
using ptx::tmem_ld_32x16;

// 256 threads: warps w and w+4 own TMEM lanes 32*(w%4)..+31 (= tile rows); the low warp of a pair works on key
// columns [0,64), the high warp on [64,128) -- two threads per row halve the softmax latency chain and double
// the warps the SM can interleave.
__global__ void __launch_bounds__(256) attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv,
                                                           __nv_bfloat16* __restrict__ att, int C, int heads_per_cta,
                                                           const int* __restrict__ tiles, const int* __restrict__ win_row0,
                                                           int B, long long* __restrict__ trace) {
  // trace (trace build only): per CTA 16 clock64 stamps of thread 0 at the phase boundaries of its first head
#define AT_STAMP(i) SAST_STAMP(trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16, trace && threadIdx.x == 0 && hi_ == 0, (i))
#ifdef SAST_TRACE
  const long long t_entry = trace ? clock64() : 0;
#endif
  // PDL: the selection record was written >= 2 kernels ago (select -> gather -> QKV GEMM -> here), so the whole prologue
  // -- index loads, barriers, TMEM -- runs before griddepcontrol.wait and overlaps the tail of the QKV GEMM
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ float pmax[2][128];
  __shared__ int wst[130];                                 // first compacted row (tile relative) of each window of this tile
  // grid.x is tile-major (j * B + b) over the per-frame slot lists, so the CTAs of unused slots are launched last
  const int slot = (int)(blockIdx.x % (unsigned)B) * (int)(gridDim.x / (unsigned)B) + (int)(blockIdx.x / (unsigned)B);
  const int2 tile = *reinterpret_cast<const int2*>(tiles + 2 * slot);  // first window, one past the last window
  if (tile.x < 0) return;                                  // unused slot
  const int w = tile.x;
  const int row0 = win_row0[w];
  const int rows = win_row0[tile.y] - row0;
  const int nwin = min(tile.y - w, 128);
  const int tid = threadIdx.x, warp = __shfl_sync(kFull, tid >> 5, 0), lane = tid & 31;   // warp: provably uniform
  const int half = warp >> 2;                              // which 64 key columns
  const int t = (warp & 3) * 32 + lane;                    // tile row = TMEM lane

  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = base;                                      // Q, K are dead once S = Q K^T has completed:
  uint8_t* sK = base + AT_TILE;                            // P (32 KB) is written over them
  uint8_t* sP = base;
  uint8_t* sV = base + 4 * AT_TILE;
  uint8_t* sOnes = base + 5 * AT_TILE;                     // 1 KB of bf16 1.0: the (layout-agnostic) B operand of the row-sum MMA
  AttnSmem* sm = reinterpret_cast<AttnSmem*>(base + 5 * AT_TILE + 1024);

  if (tid == 0) {
    ptx::tma_prefetch_desc(&map_qkv);
    ptx::mbar_init(&sm->bar_load, 1);
    ptx::mbar_init(&sm->bar_s, 1);
    ptx::mbar_init(&sm->bar_o, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc(&sm->tmem_base, AT_TMEM_COLS);
  reinterpret_cast<uint32_t*>(sOnes)[tid] = 0x3F803F80u;   // 256 threads x 4 bytes
  if (tid <= nwin) wst[tid] = win_row0[w + tid] - row0;
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_s = sm->tmem_base;
  const uint32_t tmem_o = tmem_s;
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;

  // key range of this row: the compacted rows of its own window
  int lo = 0, hi = 0;
  if (t < rows) {
    for (int i = 0; i < nwin; ++i) {                       // 2-3 windows per tile as a rule
      const int a = wst[i], b = wst[i + 1];
      if (t >= a && t < b) { lo = a; hi = b; }
    }
  }
  const int rows16 = (rows + 15) & ~15;
  const float sc = 0.17677669529663688110f * 1.44269504088896340736f;   // 32^-0.5 * log2(e)
  // 32-column chunks of this warp's half that any of its rows needs
  const int wlo = max(__reduce_min_sync(kFull, t < rows ? lo : 128) & ~31, half * 64);
  const int whi = min(__reduce_max_sync(kFull, t < rows ? hi : 0), half * 64 + 64);
  const int r8 = t & 7;
  uint8_t* prow = sP + (t >> 3) * 1024 + r8 * 128;

  const int h_begin = blockIdx.y * heads_per_cta;
  pdl_entry();                                             // qkv (the QKV GEMM's output) is read from here on
  for (int hi_ = 0; hi_ < heads_per_cta; ++hi_) {
    const int h = h_begin + hi_;
    const uint32_t ph = (uint32_t)(hi_ & 1);
#ifdef SAST_TRACE
    if (trace && tid == 0 && hi_ == 0) {
      trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + 0] = t_entry;
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + 15] = smid;
    }
#endif
    AT_STAMP(1);
    if (warp == 0) {                                          // whole warp, uniform operands; one elected lane issues
      const bool leader = ptx::elect_one();
      if (leader) {
        ptx::mbar_arrive_expect_tx(&sm->bar_load, 3 * AT_TILE);
        ptx::tma_load_2d(sQ, &map_qkv, &sm->bar_load, h * 96, row0);
        ptx::tma_load_2d(sK, &map_qkv, &sm->bar_load, h * 96 + 32, row0);
        ptx::tma_load_2d(sV, &map_qkv, &sm->bar_load, h * 96 + 64, row0);
      }
      ptx::mbar_wait(&sm->bar_load, ph);
      AT_STAMP(2);
      ptx::tc_fence_after();
      const uint32_t id_s = idesc_bf16(128, 128, 0);
      const uint64_t dq = desc_sw64_kmajor(ptx::smem_u32(sQ)), dk = desc_sw64_kmajor(ptx::smem_u32(sK));
      if (leader) {
        ptx::umma_f16_ss(tmem_s, dq, dk, id_s, 0u);
        ptx::umma_f16_ss(tmem_s, dq + 2, dk + 2, id_s, 1u);          // +32 bytes: dims 16..31
        ptx::umma_commit(&sm->bar_s);
      }
      ptx::mbar_wait(&sm->bar_s, ph);
    }
    // only warp 0 polls the mbarriers; everybody else blocks in bar.sync
    __syncthreads();
    AT_STAMP(3);
    ptx::tc_fence_after();

    // ---- softmax over this row's window: pass 1 = max over the valid columns of this thread's half ----
    float mx = -INFINITY;
    for (int c0 = wlo; c0 < whi; c0 += 32) {
      uint32_t raw[32];
      ptx::tmem_ld_32x32(tmem_s + lane_sel + (uint32_t)c0, raw);
      ptx::tmem_ld_wait();
      if (lo <= c0 && c0 + 32 <= hi) {                      // chunk entirely inside this row's window
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(raw[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c0 + j;
          if (col >= lo && col < hi) mx = fmaxf(mx, __uint_as_float(raw[j]));
        }
      }
    }
    pmax[half][t] = mx;
    AT_STAMP(4);
    __syncthreads();
    AT_STAMP(5);
    mx = fmaxf(pmax[0][t], pmax[1][t]);
    const float mxs = mx * sc;

    // ---- pass 2: p = 2^((s - max) * scale*log2e), bf16 P into the SWIZZLE_128B operand tile.  The row sums are
    // not accumulated here (unpack + add per element on an issue-bound loop): the tensor core produces them as
    // P x ones next to P x V, from the same bf16-rounded probabilities ----
    for (int c0 = half * 64; c0 < min(half * 64 + 64, rows16); c0 += 32) {
      uint32_t pk[16];
      if (c0 >= wlo && c0 < whi) {
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tmem_s + lane_sel + (uint32_t)c0, raw);
        ptx::tmem_ld_wait();
        if (lo <= c0 && c0 + 32 <= hi) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(ex2_fast(fmaf(__uint_as_float(raw[j]), sc, -mxs)),
                                                            ex2_fast(fmaf(__uint_as_float(raw[j + 1]), sc, -mxs)));
            pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&b2);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const int col = c0 + j;
            float p0 = ex2_fast(fmaf(__uint_as_float(raw[j]), sc, -mxs));
            float p1 = ex2_fast(fmaf(__uint_as_float(raw[j + 1]), sc, -mxs));
            p0 = (col >= lo && col < hi) ? p0 : 0.f;
            p1 = (col + 1 >= lo && col + 1 < hi) ? p1 : 0.f;
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
            pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&b2);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = 0u;
      }
      // 32 keys = 4 chunks of 16 bytes, SWIZZLE_128B K-major: chunk index XOR (row % 8)
      const int kb = c0 >> 6;
      const int cbase = (c0 & 63) >> 3;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int chunk = (cbase + cc) ^ r8;
        *reinterpret_cast<uint4*>(prow + kb * 16384 + chunk * 16) = make_uint4(pk[cc * 4], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]);
      }
    }
    // V rows past the tile may be uninitialised memory (0 * NaN = NaN): zero the ones the PV product reads
    if (half == 1 && t >= rows && t < rows16) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4*>(sV + t * 64 + cc * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    AT_STAMP(6);
    ptx::fence_proxy_async();                               // generic-proxy smem writes -> visible to tcgen05
    ptx::tc_fence_before();
    __syncthreads();
    AT_STAMP(7);
    if (warp == 0) {
      AT_STAMP(14);
      ptx::tc_fence_after();
      AT_STAMP(13);
      const bool leader = ptx::elect_one();
      const uint32_t id_o = idesc_bf16(128, 32, 1), id_sum = idesc_bf16(128, 16, 0);
      const uint32_t pa = ptx::smem_u32(sP), va = ptx::smem_u32(sV);
      const uint64_t d1 = desc_sw64_kmajor(ptx::smem_u32(sOnes));
      const uint64_t dp0 = ptx::umma_desc_sw128_kmajor(pa), dv0 = desc_sw64_mnmajor(va);
      const int nks = rows16 >> 4;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {                      // unrolled: the descriptor offsets are immediates
        if (ks < nks && leader) {
          // P: k-step = +32 bytes inside the 128-byte swizzle atom, next 64 keys = +16 KB; V: next 16 keys = +1 KB  (>>4 fields)
          const uint64_t dp = dp0 + (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2);
          ptx::umma_f16_ss(tmem_o, dp, dv0 + (uint64_t)(ks * 64), id_o, ks ? 1u : 0u);
          ptx::umma_f16_ss(tmem_o + 32, dp, d1, id_sum, ks ? 1u : 0u);      // columns 32..47: row sums of P
        }
      }
      AT_STAMP(11);
      if (leader) ptx::umma_commit(&sm->bar_o);
      AT_STAMP(12);
      ptx::mbar_wait(&sm->bar_o, ph);
    }
    __syncthreads();
    AT_STAMP(8);
    ptx::tc_fence_after();
    {
      uint32_t raw[16], rs;                                   // this thread's 16 of the 32 output dims, and the row sum
      tmem_ld_32x16(tmem_o + lane_sel + (uint32_t)(half * 16), raw);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(rs) : "r"(tmem_o + lane_sel + 32u) : "memory");
      ptx::tmem_ld_wait();
      if (t < rows) {
        const float il = __fdividef(1.0f, __uint_as_float(rs));
        __nv_bfloat16* dst = att + (size_t)(row0 + t) * C + h * 32 + half * 16;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint4 o;
          __nv_bfloat162 b2;
          b2 = __floats2bfloat162_rn(__uint_as_float(raw[cc * 8 + 0]) * il, __uint_as_float(raw[cc * 8 + 1]) * il); o.x = *reinterpret_cast<uint32_t*>(&b2);
          b2 = __floats2bfloat162_rn(__uint_as_float(raw[cc * 8 + 2]) * il, __uint_as_float(raw[cc * 8 + 3]) * il); o.y = *reinterpret_cast<uint32_t*>(&b2);
          b2 = __floats2bfloat162_rn(__uint_as_float(raw[cc * 8 + 4]) * il, __uint_as_float(raw[cc * 8 + 5]) * il); o.z = *reinterpret_cast<uint32_t*>(&b2);
          b2 = __floats2bfloat162_rn(__uint_as_float(raw[cc * 8 + 6]) * il, __uint_as_float(raw[cc * 8 + 7]) * il); o.w = *reinterpret_cast<uint32_t*>(&b2);
          *reinterpret_cast<uint4*>(dst + cc * 8) = o;
        }
      }
    }
    AT_STAMP(9);
    ptx::tc_fence_before();
    __syncthreads();          // everyone is done with S, O, pmax/psum and the smem tiles before the next head reuses them
    AT_STAMP(10);
    ptx::tc_fence_after();
  }
#undef AT_STAMP
  if (warp == 0) ptx::tmem_dealloc(tmem_s, AT_TMEM_COLS);
}

// ---- CUDA-core variant (debug / A-B knob SAST_B200_ATTN=simt): one thread per query row ----
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
  pdl_entry();
  extern __shared__ __align__(16) float kv[];      // k [K][32] then v [K][32]
  const int w = blockIdx.x, h = blockIdx.y;
  const int K = win_K[w];
  if (K == 0) return;
  const int row0 = win_row0[w];
  const int ld = 3 * C;
  float* ks = kv;
  float* vs = kv + (size_t)K * 32;
  for (int i = threadIdx.x; i < K * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const uint4 u = *reinterpret_cast<const uint4*>(qkv + (size_t)(row0 + r) * ld + h * 96 + 32 + c * 8);
    bf16x8_to_f32(u, c < 4 ? ks + r * 32 + c * 8 : vs + r * 32 + (c - 4) * 8);
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= K) return;
  float q[32], o[32];
  const float scale = 0.17677669529663688110f;
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
    __nv_bfloat162 t2;
    t2 = __floats2bfloat162_rn(o[c * 8 + 0], o[c * 8 + 1]); pk.x = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __floats2bfloat162_rn(o[c * 8 + 2], o[c * 8 + 3]); pk.y = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __floats2bfloat162_rn(o[c * 8 + 4], o[c * 8 + 5]); pk.z = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __floats2bfloat162_rn(o[c * 8 + 6], o[c * 8 + 7]); pk.w = *reinterpret_cast<uint32_t*>(&t2);
    *reinterpret_cast<uint4*>(dst + c * 8) = pk;
  }
}

int make_tmap_bf16_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows,
                       int swizzle_bytes);

int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* att, int C, const sast_selection& sel, int NW, int B, int T,
                        long long max_rows, int variant, cudaStream_t st) {
  const int heads = C / 32;
  if (variant == 1) {
    const size_t smem = (size_t)T * 64 * sizeof(float);
    sast::launch_k(attention_bf16_kernel, dim3(NW, heads), 128, smem, st, qkv, att, C, sel.win_K, sel.win_row0);
    SAST_LAUNCH_CHECK();
    return SAST_OK;
  }
  CUtensorMap mq;
  int rc = make_tmap_bf16_box(&mq, qkv, max_rows, 3 * C, 3 * C, 32, 128, 64);
  if (rc) return rc;
  const size_t smem = 1024 + 5 * AT_TILE + 1024 + sizeof(AttnSmem);
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  // enough CTAs to fill the chip (4 per SM): all heads in one CTA when there are many tile slots, fewer otherwise
  int hpc = heads;
  while (hpc > 1 && (long long)NW * (heads / hpc) < 4 * 148 && hpc % 2 == 0) hpc /= 2;
  sast::launch_k(attention_tc_kernel, dim3(NW, heads / hpc), 256, smem, st, mq, att, C, hpc, sel.tiles, sel.win_row0, B, g_trace_which == 1 ? g_trace : nullptr);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

}  // namespace sast
