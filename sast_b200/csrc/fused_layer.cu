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
//              (tcgen05.mma ts-form) -> O / rowsum (of the same bf16-rounded P) -> bf16 A tile
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
#include "fused_common.cuh"
#include <cstdlib>

namespace sast {

int make_tmap_bf16_box(CUtensorMap* m, const void* ptr, long long rows, int cols, int ld, int box_cols, int box_rows,
                       int swizzle_bytes);

namespace fl {

constexpr int kRingStages = 4, kRingBytes = 16384;

template <int C_, int NCTX_>
struct Cfg {
  static constexpr int C = C_;
  static constexpr int H = C / 32;                 // heads (dim_head 32)
  static constexpr int I = C * 5 / 2;              // GLU width for mlp_ratio 4: floor(4C*2/3/32)*32 = 160 / 320
  static constexpr bool kRing = C > 64;            // weights streamed through a TMA ring (C = 128) or resident (C = 64)
  static constexpr int NCTX = NCTX_;               // tile contexts per CTA (each works on its own tile)
  static constexpr int WPC = 16 / NCTX;            // warps per context
  static constexpr int TPC = WPC * 32;             // threads per context
  static constexpr int SUBS = WPC / 4;             // column slices per tile row (warp w: TMEM lane quarter w % 4, slice w / 4)
  static constexpr int TPR = SUBS / 2;             // softmax threads per (row, head of the pair)
  static constexpr int kThreads = 512 + (kRing ? 32 : 0);
  static constexpr int KB_A = C / 64;              // 64-column k-blocks of the A tile
  static constexpr int A_BYTES = KB_A * 16384;
  static constexpr int R_BYTES = 128 * 3 * C * 2;  // Q,K,V operand tiles; also shortcut staging (before) and hid (after)
  static constexpr int CTX_BYTES = A_BYTES + R_BYTES;
  static constexpr int KB_HID = (I + 63) / 64;
  static constexpr int W_QKV = 0;                  // resident layout (C = 64)
  static constexpr int W_PROJ = W_QKV + 3 * C * 128;
  static constexpr int W_1 = W_PROJ + C * 128;
  static constexpr int W_2 = W_1 + 2 * I * 128;
  static constexpr int W_BYTES = kRing ? kRingStages * kRingBytes : W_2 + KB_HID * C * 128;
  static constexpr int CPT = C / SUBS;             // accumulator columns per thread in the C-wide epilogues (32)
  static constexpr int RPP = TPC / 4;              // rows per gather pass (4 lanes per row)
  static constexpr int TM_CTX = 512 / NCTX;        // TMEM columns of one context
  // TMEM columns relative to the context base.  QKV accumulator [0,3C); during attention (QKV drained) S of the two
  // heads e of a pair, P (bf16 pairs) and O; proj and MLP-out accumulators [0,C); GLU rounds of GLU_CAP.
  // One softmax thread per row (TPR == 1): P overwrites the columns of S the same thread has already consumed and
  // O lands in the dead upper half of S, so a context needs 256 columns only.
  static constexpr int TM_S = 128, TM_P0 = TPR == 1 ? 0 : 256, TM_PS = TPR == 1 ? 128 : 64;      // S(e) = 128 e, P(e) = P0 + PS e
  static constexpr int TM_O0 = TPR == 1 ? 64 : 384, TM_OS = TPR == 1 ? 128 : 32;
  static constexpr int TM_GLU = kRing ? 128 : 64;
  static constexpr int GLU_CAP = kRing ? 384 : NCTX == 2 ? 160 : 320;                            // accumulator columns per GLU round
  // per-channel parameter vectors staged in shared memory once per CTA (floats): the 60 KB of L1 left beside the tile
  // buffers is swept by every tile's gather, so __ldg of these tiny vectors went back to L2 in every epilogue
  static constexpr int PV_LN1W = 0, PV_LN1B = C, PV_LN2W = 2 * C, PV_LN2B = 3 * C, PV_QKVB = 4 * C, PV_PROJB = 7 * C,
                       PV_G1 = 8 * C, PV_B1 = 9 * C, PV_B2 = 9 * C + 2 * I, PV_G2 = 10 * C + 2 * I, PV_FLOATS = 11 * C + 2 * I;
  static constexpr int OUT_PITCH = C * 4 + 16;     // bytes per staged output row (padding: conflict-free 16-byte stores)
  static constexpr int GLU_ROUNDS = (2 * I + GLU_CAP - 1) / GLU_CAP;
  static constexpr int RING_CHUNKS = 3 * C / 64 + C / 64 + 2 * I / 64 + KB_HID;                  // per tile (ring mode)
  static_assert(C == 64 || C == 128, "fused layer kernel: C = 64 or 128");
  static_assert(NCTX == 1 || (NCTX == 2 && !kRing), "two contexts need resident weights");
  static_assert(R_BYTES >= 128 * (C * 4 + 16) && R_BYTES >= KB_HID * 16384, "R region too small");
  static_assert(CPT == 32 || CPT == 16, "C-wide epilogues read 16 or 32 accumulator columns per thread");
  static_assert(TM_GLU + GLU_CAP <= TM_CTX && 3 * C <= TM_CTX, "TMEM plan");
};

template <int NCTX, int TPR>
struct Ctl {
  uint64_t w_full[kRingStages];
  uint64_t w_empty[kRingStages];
  uint64_t mma_bar[NCTX];
  uint32_t tmem_base;
  int pix[NCTX][128];
  int pass_id[3];                  // unselected-token pass: chunk tickets of this CTA (ring of 3: current, next, next but one)
  uint8_t lo[NCTX][128], hi[NCTX][128];
  float pmax[TPR == 2 ? 2 : 1][TPR == 2 ? 2 : 1][TPR == 2 ? 128 : 1];     // cross-thread row maxima / row sums (two softmax threads per row only)
  float psum[TPR == 2 ? 2 : 1][TPR == 2 ? 2 : 1][TPR == 2 ? 128 : 1];
};

struct Params {
  const float* x;
  float* out;
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *qkv_b, *proj_b, *gamma1, *gamma2, *mlp1_b, *mlp2_b;
  float eps;
  const int *counts, *tile_list, *row_pix, *row_win, *tok_row;
  Geom g;
  int flavor;
  long long* trace;      // debug stamps (sast_debug_trace which = 4), normally null; trace build only
};

// trace build only: [CTA][32] clock64 stamps of thread 0: 0..14 phase boundaries of context 0's SECOND tile (steady state),
// 15 kernel entry, 16 set-up done, 17 tiles done, 18 unselected pass done, 19 SM id, 20 tiles of context 0
#define FL_STAMP(i) SAST_STAMP(trc, tid == 0 && ti == 1, (i))

template <int C, int NCTX_>
// (the 544-thread ring variant gets 96 registers: the hardware budgets any block above 512 threads as 640 -- a launch with
// __maxnreg__(100) and 544 threads fails with "too many resources requested", probed with a test kernel)
__global__ void __launch_bounds__(Cfg<C, NCTX_>::kThreads, 1)
layer_fused_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_proj,
                   const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const Params p) {
  using K = Cfg<C, NCTX_>;
  using CtlT = Ctl<K::NCTX, K::TPR>;
  constexpr int H = K::H, I = K::I, CPT = K::CPT, NCTX = K::NCTX, SUBS = K::SUBS, TPR = K::TPR, TPC = K::TPC;
  extern __shared__ __align__(1024) uint8_t base[];        // no static shared memory in this kernel: the window starts 1024-aligned
  const uint32_t sW = ptx::smem_u32(base);
  if ((sW & 1023u) != 0) __trap();
  CtlT* ctl = reinterpret_cast<CtlT*>(base + K::W_BYTES + NCTX * K::CTX_BYTES);
  float* const spv = reinterpret_cast<float*>(base + K::W_BYTES + NCTX * K::CTX_BYTES + ((sizeof(CtlT) + 15) / 16) * 16);
  const uint32_t sPV = sW + K::W_BYTES + NCTX * K::CTX_BYTES + ((sizeof(CtlT) + 15) / 16) * 16;

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(kFull, tid >> 5, 0), lane = tid & 31;
  [[maybe_unused]] long long* const trc = p.trace ? p.trace + (size_t)blockIdx.x * 32 : nullptr;
  SAST_STAMP(trc, tid == 0, 15);

  if (tid == 0) {
    ptx::tma_prefetch_desc(&map_qkv); ptx::tma_prefetch_desc(&map_proj);
    ptx::tma_prefetch_desc(&map_w1); ptx::tma_prefetch_desc(&map_w2);
    for (int c = 0; c < NCTX; ++c) ptx::mbar_init(&ctl->mma_bar[c], 1);
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
  for (int i = threadIdx.x; i < K::PV_FLOATS; i += blockDim.x) {        // parameters are constants of the forward: before the PDL wait
    float v;
    if (i < K::PV_LN1B) v = p.ln1_w[i];
    else if (i < K::PV_LN2W) v = p.ln1_b[i - K::PV_LN1B];
    else if (i < K::PV_LN2B) v = p.ln2_w[i - K::PV_LN2W];
    else if (i < K::PV_QKVB) v = p.ln2_b[i - K::PV_LN2B];
    else if (i < K::PV_PROJB) v = p.qkv_b ? p.qkv_b[i - K::PV_QKVB] : 0.f;
    else if (i < K::PV_G1) v = p.proj_b ? p.proj_b[i - K::PV_PROJB] : 0.f;
    else if (i < K::PV_B1) v = p.gamma1 ? p.gamma1[i - K::PV_G1] : 1.f;
    else if (i < K::PV_B2) v = p.mlp1_b ? p.mlp1_b[i - K::PV_B1] : 0.f;
    else if (i < K::PV_G2) v = p.mlp2_b ? p.mlp2_b[i - K::PV_B2] : 0.f;
    else v = p.gamma2 ? p.gamma2[i - K::PV_G2] : 1.f;
    spv[i] = v;
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

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
    const int ctx = warp / K::WPC;                         // which tile context this warp belongs to
    const int cw = warp - ctx * K::WPC;                    // warp index inside the context
    const int ct = tid - ctx * TPC;                        // thread index inside the context
    const bool mma_warp = cw == 0;
    const bool leader = ptx::elect_one();
    const int q4 = cw & 3, sub = cw >> 2;
    const int row = q4 * 32 + lane;                        // tile row = TMEM lane
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const uint32_t r7 = (uint32_t)(row & 7);
    const uint32_t tm = ctl->tmem_base + (uint32_t)(ctx * K::TM_CTX);
    const uint32_t sA = sW + K::W_BYTES + ctx * K::CTX_BYTES;
    const uint32_t sR = sA + K::A_BYTES;
    uint64_t* const mma_bar = &ctl->mma_bar[ctx];
    uint32_t mma_phase = 0, ring_it = 0;
    const float sc = 0.17677669529663688110f * 1.44269504088896340736f;   // 32^-0.5 * log2(e)

    auto wait_mma = [&]() {
      ptx::mbar_wait(mma_bar, mma_phase);
      mma_phase ^= 1;
      ptx::tc_fence_after();
    };
    // ring consumer helpers (MMA warp only)
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

    // ---- unselected tokens keep norm1(x)  (SAST.py:251-254): this context's share of the map, 4 lanes per token.
    // With two contexts, context 1 does its share BEFORE its tiles and context 0 after: the contexts then sit in
    // different phases of their tiles (tensor-core waits of one under the epilogue arithmetic of the other).
    // The chunks of 128 tokens are handed out DYNAMICALLY (a ticket counter in the selection record, counts[4]): at low keep
    // ratios only some CTAs own a tile (55 of 148 at keep 5 %), and with a static share of the map those finished last --
    // tile + share = 47-56 k clk against 30 k for the others.  Tickets are fetched two chunks ahead and the chunk's token
    // flags one chunk ahead, so a chunk costs no more dependent latency than with a static stride.  Every CTA draws exactly
    // two tickets past the end; the CTA that draws the very last one resets the counter for the next launch on this
    // selection (a block reuses it for its second layer call and non-first blocks reuse index lists).
    auto unselected_pass = [&]() {
      constexpr int NV = C / 16;
      constexpr int TOK = TPC / 4;                           // tokens per chunk
      static_assert(NCTX == 1, "ticket accounting assumes one context per CTA");
      if ((long long)p.counts[1] >= p.g.P) return;          // every token of the map is selected (dense scene): nothing to keep
      const int l = ct & 3;
      const float inv_c = 1.0f / (float)C;
      const int nchunks = (int)((p.g.P + TOK - 1) / TOK);
      int* const ticket = const_cast<int*>(p.counts) + 4;
      auto draw = [&](int slot) {
        if (ct == 0) {
          const int c = atomicAdd(ticket, 1);
          if (c == nchunks + 2 * (int)gridDim.x - 1) atomicExch(ticket, 0);      // the last draw of the whole grid
          ctl->pass_id[slot] = c;
        }
      };
      draw(0);
      draw(1);
      ctx_sync<TPC>(ctx);
      int c_cur = ctl->pass_id[0], c_nxt = ctl->pass_id[1];
      long long q = (long long)c_cur * TOK + (ct >> 2);
      int trow = (c_cur < nchunks && q < p.g.P) ? p.tok_row[q] : 0;
      for (int k = 0; c_cur < nchunks; ++k) {
        draw((k + 2) % 3);                                   // one draw per processed chunk: exactly two failed draws per CTA
        q = (long long)c_cur * TOK + (ct >> 2);
        const bool todo = q < p.g.P && trow < 0;
        const long long qn = (long long)c_nxt * TOK + (ct >> 2);
        trow = (c_nxt < nchunks && qn < p.g.P) ? p.tok_row[qn] : 0;      // next chunk's flags, requested before this chunk's rows
        if (__any_sync(kFull, todo)) {
          const long long pix = todo ? token_pixel(q, p.g, p.flavor) : 0;
          float4 v[NV];
#pragma unroll
          for (int i = 0; i < NV; ++i)
            v[i] = todo ? __ldg(reinterpret_cast<const float4*>(p.x + pix * C + (l + 4 * i) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
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
              const float4 w4 = lds128(sPV + (uint32_t)(K::PV_LN1W + (l + 4 * i) * 4) * 4);
              const float4 b4 = lds128(sPV + (uint32_t)(K::PV_LN1B + (l + 4 * i) * 4) * 4);
              *reinterpret_cast<float4*>(p.out + pix * C + (l + 4 * i) * 4) =
                  make_float4((v[i].x - mean) * rstd * w4.x + b4.x, (v[i].y - mean) * rstd * w4.y + b4.y,
                              (v[i].z - mean) * rstd * w4.z + b4.z, (v[i].w - mean) * rstd * w4.w + b4.w);
            }
          }
        }
        ctx_sync<TPC>(ctx);                                  // the ticket drawn at the top of this iteration is in shared memory
        c_cur = c_nxt;
        c_nxt = ctl->pass_id[(k + 2) % 3];
      }
    };
    if (ctx == 1) unselected_pass();

    // Tiles of this context: t0, t0 + stride, ...  The NEXT tile's {row0, rows}, the pixel of this thread's gather row
    // and an L2 prefetch of that row are issued one tile ahead, so that a tile starts with L2-resident rows and no
    // dependent index loads.
    const int t_stride = gridDim.x * NCTX;
    const int gr = ct >> 2, gl = ct & 3;                   // gather mapping: 4 lanes per row
    int t = blockIdx.x * NCTX + ctx;
    int row0 = 0, rows = 0, split = 0;
    int gpix[128 / K::RPP];
    int bk_pix = 0, bk_win = 0;                            // threads ct < 128: pixel of tile row ct and its window {first row << 8 | K}
    if (t < n_tiles) {
      row0 = p.tile_list[2 * t]; rows = p.tile_list[2 * t + 1];
      split = rows >> 8; rows &= 255;
#pragma unroll
      for (int ps = 0; ps < 128 / K::RPP; ++ps) {
        const int src = tile_src(ps * K::RPP + gr, rows, split);
        gpix[ps] = src >= 0 ? p.row_pix[row0 + src] : -1;
      }
      if (ct < 128) {
        const int src = tile_src(ct, rows, split);
        if (src >= 0) { bk_pix = p.row_pix[row0 + src]; bk_win = p.row_win[row0 + src]; }
      }
    }
    [[maybe_unused]] int ti = -1;
    for (; t < n_tiles; t += t_stride) {
      ++ti;
      FL_STAMP(0);
      const int tn = t + t_stride;
      int nrow0 = 0, nrows = 0;                              // nrows: rows | split << 8, decoded after the QKV epilogue
      if (tn < n_tiles) { nrow0 = ldg_pinned(p.tile_list + 2 * tn); nrows = ldg_pinned(p.tile_list + 2 * tn + 1); }

      // ---- tile bookkeeping (consumed after later barriers) + gather / LN1 / LN2 -------------------------------
      if (ct < 128) {                                        // keys of row ct: the rows of its own window (a tile holds whole windows)
        const bool on = tile_src(ct, rows, split) >= 0;
        const int off = (bk_win >> 8) - row0;                // first compacted row of the row's window, tile relative
        const int lo = !on ? 0 : (split && off) ? 64 : off;
        ctl->pix[ctx][ct] = bk_pix; ctl->lo[ctx][ct] = (uint8_t)lo; ctl->hi[ctx][ct] = (uint8_t)(on ? lo + (bk_win & 255) : 0);
      }
      constexpr int NV = C / 16;                             // float4 per lane, 4 lanes per row
      float4 xin[128 / K::RPP][NV];                          // every pass's rows requested up front (y[] is not live yet)
#pragma unroll
      for (int ps = 0; ps < 128 / K::RPP; ++ps) {
        const float* xp = p.x + (long long)(gpix[ps] >= 0 ? gpix[ps] : 0) * C;
#pragma unroll
        for (int i = 0; i < NV; ++i)
          xin[ps][i] = gpix[ps] >= 0 ? __ldg(reinterpret_cast<const float4*>(xp + (gl + 4 * i) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int ps = 0; ps < 128 / K::RPP; ++ps) {
        const int r = ps * K::RPP + gr;
        const bool valid = gpix[ps] >= 0;
        float4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = xin[ps][i];
        const float inv_c = 1.0f / (float)C;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {               // LN1 then LN2 (SAST.py:206, :213)
          const uint32_t gw = sPV + (uint32_t)(pass == 0 ? K::PV_LN1W : K::PV_LN2W) * 4;
          const uint32_t gb = sPV + (uint32_t)(pass == 0 ? K::PV_LN1B : K::PV_LN2B) * 4;
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
            const float4 w4 = lds128(gw + (uint32_t)(gl + 4 * i) * 16);
            const float4 b4 = lds128(gb + (uint32_t)(gl + 4 * i) * 16);
            v[i].x = (v[i].x - mean) * rstd * w4.x + b4.x; v[i].y = (v[i].y - mean) * rstd * w4.y + b4.y;
            v[i].z = (v[i].z - mean) * rstd * w4.z + b4.z; v[i].w = (v[i].w - mean) * rstd * w4.w + b4.w;
          }
        }
        if (ps == 0) {                                         // the previous tile's output rows have left the staging area
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          ctx_sync<TPC>(ctx);
        }
        const uint32_t rr7 = (uint32_t)(r & 7);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (!valid) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);    // rows past the tile: finite operands
          const int c = (gl + 4 * i) * 4;
          const uint32_t chunk = (uint32_t)((c & 63) >> 3);
          sts64(sA + (uint32_t)((c >> 6) * 16384 + r * 128) + ((chunk ^ rr7) << 4) + (uint32_t)((c & 7) * 2),
                pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w));
          const uint32_t ch4 = (uint32_t)(gl + 4 * i);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sR + (uint32_t)(r * C * 4) + ((ch4 ^ rr7) << 4)),
                       "f"(v[i].x), "f"(v[i].y), "f"(v[i].z), "f"(v[i].w) : "memory");
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ctx_sync<TPC>(ctx);
      FL_STAMP(1);

      // ---- QKV = n2 Wqkv^T ----------------------------------------------------------------------------------------
      if (mma_warp) {
        ptx::tc_fence_after();
        if (!K::kRing) {
          mma_kblock(leader, tm, sA, sW + K::W_QKV, idesc(128, 3 * C, 0), 4, true);
        } else {
          for (int j = 0; j < 3 * C / 64; ++j) {
            const uint32_t st = ring_acquire();
            for (int kb = 0; kb < K::KB_A; ++kb)
              mma_kblock(leader, tm + 64 * j, sA + kb * 16384, st + kb * 8192, idesc(128, 64, 0), 4, kb == 0);
            ring_release();
          }
        }
        if (leader) ptx::umma_commit(mma_bar);
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
      ctx_sync<TPC>(ctx);                                    // every shortcut read is done: the Q,K,V tiles may overwrite it

      // ---- QKV epilogue: + bias, bf16, operand tiles of the attention MMAs ------------------------------------
      for (int u = sub; u < 3 * H; u += SUBS) {              // 32 accumulator columns = q, k or v of one head
        uint32_t raw[32];
        ptx::tmem_ld_32x32(tm + lane_sel + (uint32_t)(u * 32), raw);
        ptx::tmem_ld_wait();
        const int h = u / 3, which = u - 3 * h;
        const uint32_t tile = which < 2 ? sR + (uint32_t)(h * 16384 + which * 8192) : sR + (uint32_t)(H * 16384 + h * 8192);
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
      ctx_sync<TPC>(ctx);
      FL_STAMP(3);
      // next tile: pixel of this thread's gather rows (the tile_list entry requested at the top has landed by now) and,
      // for threads ct < 128, the row table: each dependent load is issued one phase before its value is needed
      const int nsplit = nrows >> 8;
      nrows &= 255;
      int npix[128 / K::RPP];
#pragma unroll
      for (int ps = 0; ps < 128 / K::RPP; ++ps) {
        const int src = tile_src(ps * K::RPP + gr, nrows, nsplit);
        npix[ps] = src >= 0 ? ldg_pinned(p.row_pix + nrow0 + src) : -1;
      }
      int nb_pix = 0, nb_win = 0;
      if (ct < 128) {
        const int src = tile_src(ct, nrows, nsplit);
        if (src >= 0) { nb_pix = ldg_pinned(p.row_pix + nrow0 + src); nb_win = ldg_pinned(p.row_win + nrow0 + src); }
      }

      // ---- attention, two heads at a time -------------------------------------------------------------------------
      const int lo = ctl->lo[ctx][row], hi = ctl->hi[ctx][row];
      const bool rvalid = hi > lo;                           // rows past the tile / in the alignment gap hold no token
      const int hh = sub / TPR, half = sub % TPR;            // softmax / O mapping: head of the pair, part of the columns
      constexpr int SCOLS = 128 / TPR;                       // S columns per softmax thread
      for (int hp = 0; hp < H / 2; ++hp) {
        if (mma_warp) {
          ptx::tc_fence_after();
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t qa = sR + (uint32_t)((2 * hp + e) * 16384);
            const uint64_t dq = desc_sw64_k(qa), dk = desc_sw64_k(qa + 8192);
            if (leader) {
              ptx::umma_f16_ss(tm + K::TM_S * e, dq, dk, idesc(128, 128, 0), 0u);
              ptx::umma_f16_ss(tm + K::TM_S * e, dq + 2, dk + 2, idesc(128, 128, 0), 1u);     // dims 16..31
            }
          }
          if (leader) ptx::umma_commit(mma_bar);
        }
        wait_mma();
        FL_STAMP(4);
        // pass 1: row maximum over the keys of the row's own window, this thread's columns
        float mx = -INFINITY;
        bool need[SCOLS / 32], full[SCOLS / 32];
#pragma unroll
        for (int cc = 0; cc < SCOLS / 32; ++cc) {
          const int c0 = half * SCOLS + cc * 32;
          need[cc] = __any_sync(kFull, lo < c0 + 32 && hi > c0);
          // whole chunk inside the window of every row of the warp that has one: no per-element masks (rows without a
          // token compute garbage that is never stored)
          full[cc] = __all_sync(kFull, !rvalid || (lo <= c0 && c0 + 32 <= hi));
          if (need[cc]) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tm + lane_sel + (uint32_t)(K::TM_S * hh + c0), raw);
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
        if constexpr (TPR == 2) {
          ctl->pmax[hh][half][row] = mx;
          ctx_sync<TPC>(ctx);
          mx = fmaxf(ctl->pmax[hh][0][row], ctl->pmax[hh][1][row]);
        }
        FL_STAMP(5);
        const float mxs = mx * sc;
        float rsum = 0.f;                                    // row sum of the bf16-ROUNDED probabilities (what P V multiplies)
        // pass 2: p = 2^((s - max) scale log2e) as bf16 pairs into TMEM -- the A operand of P V.  (TPR == 1: P overwrites
        // S columns this thread has already read: 16 P columns per 32 S columns, always behind the read position.)
#pragma unroll
        for (int cc = 0; cc < SCOLS / 32; ++cc) {
          const int c0 = half * SCOLS + cc * 32;
          uint32_t pk[16];
          if (need[cc]) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tm + lane_sel + (uint32_t)(K::TM_S * hh + c0), raw);
            ptx::tmem_ld_wait();
            if (full[cc]) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const uint32_t w = pack_bf16(ex2_approx(fmaf(__uint_as_float(raw[j]), sc, -mxs)),
                                             ex2_approx(fmaf(__uint_as_float(raw[j + 1]), sc, -mxs)));
                pk[j >> 1] = w;
                rsum += __uint_as_float(w << 16) + __uint_as_float(w & 0xFFFF0000u);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(raw[j]), sc, -mxs));
                float p1 = ex2_approx(fmaf(__uint_as_float(raw[j + 1]), sc, -mxs));
                p0 = (c0 + j >= lo && c0 + j < hi) ? p0 : 0.f;
                p1 = (c0 + j + 1 >= lo && c0 + j + 1 < hi) ? p1 : 0.f;
                const uint32_t w = pack_bf16(p0, p1);
                pk[j >> 1] = w;
                rsum += __uint_as_float(w << 16) + __uint_as_float(w & 0xFFFF0000u);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
          ptx::tmem_st_32x16(tm + lane_sel + (uint32_t)(K::TM_P0 + K::TM_PS * hh + (c0 >> 1)), pk);
        }
        if constexpr (TPR == 2) ctl->psum[hh][half][row] = rsum;
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ctx_sync<TPC>(ctx);
        FL_STAMP(6);
        if (mma_warp) {
          ptx::tc_fence_after();
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint64_t dv = desc_sw64_mn(sR + (uint32_t)(H * 16384 + (2 * hp + e) * 8192));
            const uint32_t ta = tm + (uint32_t)(K::TM_P0 + K::TM_PS * e);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              if (leader) {
                // P: one k-step = 16 keys = 8 TMEM columns; V: next 16 keys = +1 KB
                ptx::umma_f16_ts(tm + (uint32_t)(K::TM_O0 + K::TM_OS * e), ta + 8 * ks, dv + (uint64_t)(ks * 64), idesc(128, 32, 1), ks ? 1u : 0u);
              }
            }
          }
          if (leader) ptx::umma_commit(mma_bar);
        }
        if (hp == 0) {                                       // next tile's rows towards L2 while the tensor core works
#pragma unroll
          for (int ps = 0; ps < 128 / K::RPP; ++ps)
            if (npix[ps] >= 0 && gl < C / 32) prefetch_l2(p.x + (long long)npix[ps] * C + gl * 32);
        }
        wait_mma();
        FL_STAMP(7);
        {
          constexpr int OC = 32 / TPR;                         // output dims per thread
          uint32_t raw[OC];
          if constexpr (OC == 32) ptx::tmem_ld_32x32(tm + lane_sel + (uint32_t)(K::TM_O0 + K::TM_OS * hh), raw);
          else ptx::tmem_ld_32x16(tm + lane_sel + (uint32_t)(K::TM_O0 + K::TM_OS * hh + OC * half), raw);
          ptx::tmem_ld_wait();
          if constexpr (TPR == 2) rsum = ctl->psum[hh][0][row] + ctl->psum[hh][1][row];
          const float il = rvalid ? __fdividef(1.0f, rsum) : 0.f;
          const int col = (2 * hp + hh) * 32 + half * OC;
          const uint32_t dst = sA + (uint32_t)((col >> 6) * 16384 + row * 128);
          const uint32_t ch = (uint32_t)((col & 63) >> 3);
#pragma unroll
          for (int c = 0; c < OC / 8; ++c)
            sts128(dst + (((ch + c) ^ r7) << 4),
                   pack_bf16(__uint_as_float(raw[8 * c]) * il, __uint_as_float(raw[8 * c + 1]) * il),
                   pack_bf16(__uint_as_float(raw[8 * c + 2]) * il, __uint_as_float(raw[8 * c + 3]) * il),
                   pack_bf16(__uint_as_float(raw[8 * c + 4]) * il, __uint_as_float(raw[8 * c + 5]) * il),
                   pack_bf16(__uint_as_float(raw[8 * c + 6]) * il, __uint_as_float(raw[8 * c + 7]) * il));
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        ctx_sync<TPC>(ctx);
        FL_STAMP(8);
      }

      // ---- proj + LayerScale + shortcut: y = n2 + g1 (o Wp^T + b) ---------------------------------------------
      if (mma_warp) {
        ptx::tc_fence_after();
        if (!K::kRing) {
          mma_kblock(leader, tm, sA, sW + K::W_PROJ, idesc(128, C, 0), 4, true);
        } else {
          for (int j = 0; j < C / 64; ++j) {
            const uint32_t st = ring_acquire();
            for (int kb = 0; kb < K::KB_A; ++kb)
              mma_kblock(leader, tm + 64 * j, sA + kb * 16384, st + kb * 8192, idesc(128, 64, 0), 4, kb == 0);
            ring_release();
          }
        }
        if (leader) ptx::umma_commit(mma_bar);
      }
      wait_mma();
      FL_STAMP(9);
      {
        const int col0 = sub * CPT;
        uint32_t raw[CPT];
        if constexpr (CPT == 32) ptx::tmem_ld_32x32(tm + lane_sel + (uint32_t)col0, raw);
        else ptx::tmem_ld_32x16(tm + lane_sel + (uint32_t)col0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          const float4 b4 = lds128(sPV + (uint32_t)(K::PV_PROJB + col0 + j) * 4);
          const float4 g4 = lds128(sPV + (uint32_t)(K::PV_G1 + col0 + j) * 4);
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
      ctx_sync<TPC>(ctx);
      FL_STAMP(10);

      // ---- GLU: hid = val * gelu(gate) of y W1^T + b  (ops.py:135-137; weight rows interleaved value_j, gate_j) ----
      for (int rd = 0; rd < K::GLU_ROUNDS; ++rd) {
        const int ncols = min(2 * I - rd * K::GLU_CAP, K::GLU_CAP);     // accumulator columns of this round
        if (mma_warp) {
          ptx::tc_fence_after();
          if (!K::kRing) {
            for (int j = 0; j < ncols / 160; ++j)              // B tiles of 160 weight rows (N <= 256 per instruction)
              mma_kblock(leader, tm + K::TM_GLU + 160 * j, sA, sW + K::W_1 + (rd * K::GLU_CAP + 160 * j) * 128, idesc(128, 160, 0), 4, true);
          } else {
            for (int j = 0; j < ncols / 64; ++j) {
              const uint32_t st = ring_acquire();
              for (int kb = 0; kb < K::KB_A; ++kb)
                mma_kblock(leader, tm + K::TM_GLU + 64 * j, sA + kb * 16384, st + kb * 8192, idesc(128, 64, 0), 4, kb == 0);
              ring_release();
            }
          }
          if (leader) ptx::umma_commit(mma_bar);
        }
        wait_mma();
        FL_STAMP(11);
        for (int u = sub; u < ncols / 16; u += SUBS) {          // 16 accumulator columns -> 8 hid columns = one 16-byte chunk
          uint32_t raw[16];
          ptx::tmem_ld_32x16(tm + lane_sel + (uint32_t)(K::TM_GLU + 16 * u), raw);
          ptx::tmem_ld_wait();
          const int ac = rd * K::GLU_CAP + 16 * u;             // global accumulator column (= interleaved bias index)
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = lds128(sPV + (uint32_t)(K::PV_B1 + ac + 4 * j) * 4);
            pk[j] = pack_bf16(glu_tanh_fit(0.5f * (__uint_as_float(raw[4 * j]) + b4.x), __uint_as_float(raw[4 * j + 1]) + b4.y),
                              glu_tanh_fit(0.5f * (__uint_as_float(raw[4 * j + 2]) + b4.z), __uint_as_float(raw[4 * j + 3]) + b4.w));
          }
          const int hc = ac >> 1;
          sts128(sR + (uint32_t)((hc >> 6) * 16384 + row * 128) + ((((uint32_t)(hc & 63) >> 3) ^ r7) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        ctx_sync<TPC>(ctx);
        FL_STAMP(12);
      }

      // ---- MLP out + LayerScale + residual + scatter-back: out[pix] = y + g2 (hid W2^T + b) ---------------------
      if (mma_warp) {
        ptx::tc_fence_after();
        for (int kb = 0; kb < K::KB_HID; ++kb) {
          const int nks = min(4, (I - kb * 64) / 16);
          if (!K::kRing) {
            mma_kblock(leader, tm, sR + kb * 16384, sW + K::W_2 + kb * C * 128, idesc(128, C, 0), nks, kb == 0);
          } else {
            const uint32_t st = ring_acquire();
            mma_kblock(leader, tm, sR + kb * 16384, st, idesc(128, C, 0), nks, kb == 0);
            ring_release();
          }
        }
        if (leader) ptx::umma_commit(mma_bar);
      }
      wait_mma();
      FL_STAMP(13);
      {
        const int col0 = sub * CPT;
        uint32_t raw[CPT];
        if constexpr (CPT == 32) ptx::tmem_ld_32x32(tm + lane_sel + (uint32_t)col0, raw);
        else ptx::tmem_ld_32x16(tm + lane_sel + (uint32_t)col0, raw);
        ptx::tmem_ld_wait();
        // rows -> shared memory (hid is dead: its MMA has completed; 16-byte padded pitch: conflict-free), then ONE bulk
        // async copy per row to the map (scatter-back): a thread owns CPT columns of one row here, which as direct stores
        // would be 32 half-filled sectors per instruction
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          const float4 b4 = lds128(sPV + (uint32_t)(K::PV_B2 + col0 + j) * 4);
          const float4 g4 = lds128(sPV + (uint32_t)(K::PV_G2 + col0 + j) * 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sR + (uint32_t)(row * K::OUT_PITCH + (col0 + j) * 4)),
                       "f"(fmaf(g4.x, __uint_as_float(raw[j]) + b4.x, y[j])), "f"(fmaf(g4.y, __uint_as_float(raw[j + 1]) + b4.y, y[j + 1])),
                       "f"(fmaf(g4.z, __uint_as_float(raw[j + 2]) + b4.z, y[j + 2])), "f"(fmaf(g4.w, __uint_as_float(raw[j + 3]) + b4.w, y[j + 3]))
                       : "memory");
        }
      }
      FL_STAMP(21);
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ctx_sync<TPC>(ctx);
      FL_STAMP(22);
      {   // one bulk copy per row, spread over all warps (a warp issues its lanes' copies one after the other)
        constexpr int RPW = 128 / K::WPC;                    // rows per warp
        const int orow = cw * RPW + lane;
        if (lane < RPW && ctl->hi[ctx][orow] > ctl->lo[ctx][orow]) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.out + (long long)ctl->pix[ctx][orow] * C),
                       "r"(sR + (uint32_t)(orow * K::OUT_PITCH)), "r"(C * 4) : "memory");
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      FL_STAMP(23);
      ptx::tc_fence_before();
      ctx_sync<TPC>(ctx);                                    // TMEM, the A tile and ctl->pix/lo/hi are free for the next tile
      FL_STAMP(14);
      bk_pix = nb_pix; bk_win = nb_win;
      row0 = nrow0; rows = nrows; split = nsplit;
#pragma unroll
      for (int ps = 0; ps < 128 / K::RPP; ++ps) gpix[ps] = npix[ps];
    }
    SAST_STAMP(trc, tid == 0, 17);
#ifdef SAST_TRACE
    if (trc && tid == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); trc[19] = smid; trc[20] = ti + 1; }
#endif

    if (ctx == 0) unselected_pass();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");          // the last tile's rows have reached the map
    SAST_STAMP(trc, tid == 0, 18);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(ctl->tmem_base, 512);
  }
}

template <int C, int NCTX>
static int launch_fused_t(const sast_layer_args& a, const Geom& g, cudaStream_t st) {
  using K = Cfg<C, NCTX>;
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
  p.counts = a.sel.counts; p.tile_list = a.sel.tile_list; p.row_pix = a.sel.row_pix; p.row_win = a.sel.row_win;
  p.tok_row = a.sel.tok_row;
  p.g = g; p.flavor = a.flavor;
  p.trace = g_trace_which == 4 ? g_trace : nullptr;
  const size_t smem = (size_t)K::W_BYTES + K::NCTX * K::CTX_BYTES + (sizeof(Ctl<K::NCTX, K::TPR>) + 15) / 16 * 16 + K::PV_FLOATS * 4;
  static thread_local unsigned long long attr_mask = 0;
  if (first_use_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(layer_fused_kernel<C, NCTX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  const long long chunks = (g.P + 127) / 128;
  const unsigned grid = (unsigned)(chunks < sms ? chunks : sms);
  sast::launch_k(layer_fused_kernel<C, NCTX>, grid, K::kThreads, smem, st, mq, mp, m1, m2, p);
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

// Tile contexts per CTA: the kernel template also instantiates with two independent 8-warp contexts sharing C = 64's
// resident weights (NCTX = 2).  Measured equal to one 16-warp context on the 1 Mpx B=8 stage-1 shape (19.6 k clk per tile
// against 23 k, but 3.46 tiles per context quantise to 4 rounds; profiles/r02_fused_trace_v7_2ctx.txt) and no longer
// fitting shared memory next to the staged parameter vectors, so only NCTX = 1 is built.

// true if this layer can take the fused kernel (bf16 path, C 64 / 128 with the mlp_ratio-4 GLU width, no context broadcast)
bool fused_layer_supported(const sast_layer_args& a) {
  if (a.precision != SAST_BF16 || a.enable_cb || !fused_layer_enabled()) return false;
  if (a.w.dim_head != 0 && a.w.dim_head != 32) return false;
  if (a.g.C == 64) return a.w.I == fl::Cfg<64, 1>::I;
  // C = 128 streams 368 KB of weights per tile: below ~64 tiles (Gen1 B=1 stage 2: 10) the N-split GEMM chain, which
  // spreads one weight pass over all SMs, is faster (measured: 67 us against ~50 us per layer)
  if (a.g.C == 128) return a.w.I == fl::Cfg<128, 1>::I && (a.g.B <= 0 || (long long)a.g.B * a.g.H * a.g.W >= 64 * 128);
  return false;
}

int launch_layer_fused(const sast_layer_args& a, const Geom& g, cudaStream_t st) {
  if (!a.sel.tile_list || !a.sel.row_win) return SAST_E_NULL;
  if (g.C == 64) return fl::launch_fused_t<64, 1>(a, g, st);
  if (g.C == 128) return fl::launch_fused_t<128, 1>(a, g, st);
  return SAST_E_UNSUPPORTED;
}

}  // namespace sast
