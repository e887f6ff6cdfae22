// a5/a6: scene-adaptive window + token selection and index compaction
// (replaces SAST.py:84-96 window_selection/token_selection and :258-281
//  get_score_index_2d21d / get_score_index_with_padding, plus the isin() at :122/:147).
//
// Four small launches, no host round trip, both partition flavours of a block (window layer and
// grid layer share the same per-token scores, SAST.py:141-142) batched along gridDim.z:
//   win_logit_kernel     : warp per window, mean of its T token scores.
//   select_flags_kernel  : CTA per 8 windows: frame softmax statistics over the N logits, window
//                          keep = prob >= thr_win; per kept window softmax over T token scores,
//                          keep via __ballot_sync, K = popc.  (Or thresholds on given
//                          probabilities / given flags: modes PROBS, FLAGS.)
//   select_scan_kernel   : CTA per frame: exclusive prefix sums across frames and windows ->
//                          win_rank, sel_win, win_row0, counts{M,S,Kmax}; greedy packing of
//                          consecutive windows into <=128-row attention tiles (dense per-frame slot list).
//   select_tokens_kernel : warp per window: ballot/popc prefix inside the window -> tok_row,
//                          row_tok, row_pix (compacted row <-> token <-> NHWC pixel).
// The compare is `prob >= thr` on fp32 with thr = fp32((1/N)/(1+BOUNCE)) exactly as torch
// evaluates `x >= d / (1 + b)`; ids come out ascending like torch.nonzero.
#include "common.cuh"

namespace sast {

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / 32;

struct SelectParams {
  sast_select_args a;      // primary selection (a.flavor, a.sel)
  sast_selection sel_b;    // optional second selection over the same scores
  int flavor_b;
};

__device__ __forceinline__ const sast_selection& pick_sel(const SelectParams& p, int z) { return z == 0 ? p.a.sel : p.sel_b; }
__device__ __forceinline__ int pick_flavor(const SelectParams& p, int z) { return z == 0 ? p.a.flavor : p.flavor_b; }

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
  return r;
}

__global__ void __launch_bounds__(kSelThreads) win_logit_kernel(SelectParams p) {
  pdl_entry();
  const int flavor = pick_flavor(p, blockIdx.z);
  const sast_selection& sel = pick_sel(p, blockIdx.z);
  const Geom g = make_geom(p.a.g, flavor);
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * kSelWarps + (threadIdx.x >> 5);
  if (w >= g.NW) return;
  const int b = w / g.N, n = w - b * g.N;
  const float* fs = p.a.tok_score + (size_t)b * g.H * g.W;
  float s = 0.f;
  for (int t = lane; t < g.T; t += 32) s += fs[frame_pixel(n, t, g.H, g.W, g.p0, g.p1, flavor)];
  s = warp_sum(s);
  if (lane == 0) sel.win_logit[w] = s / (float)g.T;
}

__global__ void __launch_bounds__(kSelThreads) select_flags_kernel(SelectParams p) {
  pdl_entry();
  const int flavor = pick_flavor(p, blockIdx.z);
  const sast_selection& sel = pick_sel(p, blockIdx.z);
  const sast_select_args& a = p.a;
  const Geom g = make_geom(a.g, flavor);
  __shared__ float red[kSelWarps];
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    sel.counts[3] = 0;            // tile_list fill count (select_scan adds)
    sel.counts[4] = 0;            // ticket counter of the layer kernels' unselected-token pass (self-resetting after each launch)
  }
  const int b = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = blockIdx.x * kSelWarps + wid;
  const int w = b * g.N + n;

  // ---- window keep ----------------------------------------------------------------------
  float mx = 0.f, se = 1.f;
  if (a.mode == SAST_SEL_SCORES) {   // frame softmax statistics, identical in every CTA of the frame
    const float* wl = sel.win_logit + (size_t)b * g.N;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < g.N; i += blockDim.x) m = fmaxf(m, wl[i]);
    mx = block_reduce_max(m, red);
    float e = 0.f;
    for (int i = threadIdx.x; i < g.N; i += blockDim.x) e += expf(wl[i] - mx);
    se = block_reduce_sum(e, red);
  }
  if (n >= g.N) return;               // warp-uniform, after the block-wide reductions
  bool wkeep;
  if (a.mode == SAST_SEL_SCORES) {
    const float pw = expf(sel.win_logit[w] - mx) / se;
    wkeep = pw >= a.thr_win;
    if (a.win_prob_out && blockIdx.z == 0 && lane == 0) a.win_prob_out[w] = pw;
  } else if (a.mode == SAST_SEL_PROBS) {
    wkeep = a.win_prob[w] >= a.thr_win;
  } else {
    wkeep = a.win_flag[w] != 0;
  }

  // ---- tokens ------------------------------------------------------------------------------
  const size_t q0 = (size_t)w * g.T;
  if (!wkeep) {
    for (int t = lane; t < g.T; t += 32) sel.tok_keep[q0 + t] = 0;
    if (lane == 0) { sel.win_K[w] = 0; sel.win_rank[w] = -1; }
    return;
  }
  float v[4];
  bool keep[4];
  if (a.mode == SAST_SEL_SCORES) {
    const float* fs = a.tok_score + (size_t)b * g.H * g.W;
    float tm = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = i * 32 + lane;
      v[i] = t < g.T ? fs[frame_pixel(n, t, g.H, g.W, g.p0, g.p1, flavor)] : -INFINITY;
      tm = fmaxf(tm, v[i]);
    }
    tm = warp_max(tm);
    float ts = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = (i * 32 + lane < g.T) ? expf(v[i] - tm) : 0.f;
      ts += v[i];
    }
    ts = warp_sum(ts);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = i * 32 + lane;
      const float pt = v[i] / ts;
      keep[i] = t < g.T && pt >= a.thr_tok;
      if (a.tok_prob_out && blockIdx.z == 0 && t < g.T) a.tok_prob_out[q0 + t] = pt;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = i * 32 + lane;
      keep[i] = false;
      if (t < g.T) keep[i] = a.mode == SAST_SEL_PROBS ? (a.tok_prob[q0 + t] >= a.thr_tok) : (a.tok_flag[q0 + t] != 0);
    }
  }
  int K = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = i * 32 + lane;
    K += __popc(__ballot_sync(kFull, keep[i]));
    if (t < g.T) sel.tok_keep[q0 + t] = keep[i] ? 1 : 0;
  }
  if (lane == 0) { sel.win_K[w] = K; sel.win_rank[w] = 0; }
}

// exclusive scan of two ints over the block; returns block totals through tot0/tot1
__device__ __forceinline__ void block_excl_scan2(int v0, int v1, int& e0, int& e1, int& tot0, int& tot1, int (*ws)[2]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int i0 = v0, i1 = v1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t0 = __shfl_up_sync(kFull, i0, o), t1 = __shfl_up_sync(kFull, i1, o);
    if (lane >= o) { i0 += t0; i1 += t1; }
  }
  __syncthreads();
  if (lane == 31) { ws[wid][0] = i0; ws[wid][1] = i1; }
  __syncthreads();
  int b0 = 0, b1 = 0;
  tot0 = 0; tot1 = 0;
  for (int i = 0; i < nwarp; ++i) {
    if (i < wid) { b0 += ws[i][0]; b1 += ws[i][1]; }
    tot0 += ws[i][0]; tot1 += ws[i][1];
  }
  e0 = b0 + i0 - v0;
  e1 = b1 + i1 - v1;
}

__global__ void __launch_bounds__(kSelThreads) select_scan_kernel(SelectParams p) {
  pdl_entry();
  const int flavor = pick_flavor(p, blockIdx.z);
  const sast_selection& sel = pick_sel(p, blockIdx.z);
  const Geom g = make_geom(p.a.g, flavor);
  extern __shared__ int pre[];           // [N+1] exclusive prefix of K inside this frame, then [N] next-tile pointers
  int* const nxt = pre + g.N + 1;
  __shared__ int ws[kSelWarps][2];
  __shared__ int pack_n, pack_base;
  __shared__ int red3[kSelWarps][3];
  const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  // totals of all earlier frames (int adds: order-independent)
  int m = 0, s = 0, km = 0;
  for (int w = threadIdx.x; w < b * g.N; w += blockDim.x) {
    const int K = sel.win_K[w];
    m += sel.win_rank[w] >= 0; s += K; km = max(km, K);
  }
  m = warp_sum_i(m); s = warp_sum_i(s); km = warp_max_i(km);
  if (lane == 0) { red3[wid][0] = m; red3[wid][1] = s; red3[wid][2] = km; }
  __syncthreads();
  int run_m = 0, run_s = 0, kmax = 0;
  for (int i = 0; i < kSelWarps; ++i) { run_m += red3[i][0]; run_s += red3[i][1]; kmax = max(kmax, red3[i][2]); }
  const int frame_base = run_s;            // first compacted row of this frame

  for (int n0 = 0; n0 < g.N; n0 += blockDim.x) {
    const int n = n0 + threadIdx.x;
    const int w = b * g.N + n;
    const bool in = n < g.N;
    const int kept = in ? (sel.win_rank[w] >= 0) : 0;
    const int K = in ? sel.win_K[w] : 0;
    kmax = max(kmax, K);
    int em, es, tm, ts;
    block_excl_scan2(kept, K, em, es, tm, ts, ws);
    if (in) {
      sel.win_row0[w] = run_s + es;
      pre[n] = run_s + es - frame_base;
      if (kept) {
        sel.win_rank[w] = run_m + em;
        sel.sel_win[run_m + em] = w;
      }
      sel.tiles[2 * w] = -1;           // tile slot n of this frame: unused until the packer below claims it
      sel.tiles[2 * w + 1] = 0;
    }
    run_m += tm; run_s += ts;
    __syncthreads();
  }
  if (b == g.B - 1) {
    kmax = warp_max_i(kmax);
    __syncthreads();
    if (lane == 0) red3[wid][2] = kmax;
    __syncthreads();
    if (threadIdx.x == 0) {
      int k = 0;
      for (int i = 0; i < kSelWarps; ++i) k = max(k, red3[i][2]);
      sel.counts[0] = run_m; sel.counts[1] = run_s; sel.counts[2] = k;
      sel.win_row0[g.NW] = run_s;
    }
  }
  __syncthreads();
  // Greedy packing of consecutive windows of this frame into attention tiles of <= 128 rows and <= 128 windows.  The
  // j-th tile of frame b goes to slot b*N + j: tiles[2 slot] = first window, tiles[2 slot + 1] = one past its last
  // window (first window -1: slot unused); slots are dense from j = 0, so a tile-major grid has its idle CTAs last.
  // The same tiles are appended to the dense work list tile_list as {first compacted row, rows} (counts[3] entries,
  // order arbitrary).  Parallel in three steps: (1) every window n finds by binary search on the prefix sums where a
  // tile STARTING at n would end; (2) one thread follows those pointers from window 0 -- one shared-memory load per
  // tile instead of a walk over all windows -- and claims the slots; (3) one atomicAdd reserves the frame's range of
  // tile_list, which all threads then fill.
  if (threadIdx.x == 0) pre[g.N] = run_s - frame_base;
  __syncthreads();
  for (int n = threadIdx.x; n < g.N; n += blockDim.x) {
    int lo = n + 1, hi = min(n + 128, g.N);          // largest m in [lo, hi] with pre[m] - pre[n] <= 128 (m = n + 1 always fits: K <= T <= 128)
    const int p0 = pre[n];
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (pre[mid] - p0 <= 128) lo = mid; else hi = mid - 1;
    }
    nxt[n] = lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int j = 0;
    for (int n = 0; n < g.N;) {
      const int m = nxt[n];
      if (pre[m] > pre[n]) {                          // tiles made of dropped windows only are not emitted
        const int slot = b * g.N + j;
        sel.tiles[2 * slot] = b * g.N + n;
        sel.tiles[2 * slot + 1] = b * g.N + m;
        ++j;
      }
      n = m;
    }
    pack_n = j;
    pack_base = j ? atomicAdd(&sel.counts[3], j) : 0;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < pack_n; j += blockDim.x) {
    const int slot = b * g.N + j;
    const int n = sel.tiles[2 * slot] - b * g.N, m = sel.tiles[2 * slot + 1] - b * g.N;
    // "aligned" tiles: exactly two selected windows of <= 64 tokens each (the dense 1 Mpx case: 2 x 60).  The fused layer
    // kernel then places the second window at tile row 64, so that a row's keys never straddle a 64-column block of S.
    int nsel = 0, k0 = 0, k1 = 0;
    for (int w = n; w < m; ++w) {
      const int K = pre[w + 1] - pre[w];
      if (K > 0) { if (nsel == 0) k0 = K; else k1 = K; ++nsel; }
    }
    const int split = (nsel == 2 && k0 <= 64 && k1 <= 64) ? k0 : 0;
    sel.tile_list[2 * (pack_base + j)] = frame_base + pre[n];
    sel.tile_list[2 * (pack_base + j) + 1] = (pre[m] - pre[n]) | (split << 8);
  }
}

__global__ void __launch_bounds__(kSelThreads) select_tokens_kernel(SelectParams p) {
  pdl_entry();
  const int flavor = pick_flavor(p, blockIdx.z);
  const sast_selection& sel = pick_sel(p, blockIdx.z);
  const Geom g = make_geom(p.a.g, flavor);
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * kSelWarps + (threadIdx.x >> 5);
  if (w >= g.NW) return;
  const int b = w / g.N, n = w - b * g.N;
  const size_t q0 = (size_t)w * g.T;
  const int K = sel.win_K[w];
  const int row0 = sel.win_row0[w];
  int run = 0;
  for (int t0 = 0; t0 < g.T; t0 += 32) {
    const int t = t0 + lane;
    const bool keep = K > 0 && t < g.T && sel.tok_keep[q0 + t] != 0;
    const unsigned bal = __ballot_sync(kFull, keep);
    const int pre = __popc(bal & ((1u << lane) - 1u));
    if (t < g.T) sel.tok_row[q0 + t] = keep ? row0 + run + pre : -1;
    if (keep) {
      sel.row_tok[row0 + run + pre] = (int)(q0 + t);
      sel.row_pix[row0 + run + pre] = b * g.H * g.W + frame_pixel(n, t, g.H, g.W, g.p0, g.p1, flavor);
      sel.row_win[row0 + run + pre] = (row0 << 8) | K;
    }
    run += __popc(bal);
  }
}

}  // namespace sast

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t sast_selection_bytes(int32_t B, int32_t NW, int32_t P) {
  (void)B;
  size_t n = 0;
  n += align_up(8 * 4, 16);                     // counts
  n += 4 * align_up((size_t)NW * 4, 16);        // win_K, win_rank, sel_win, win_logit
  n += align_up(((size_t)NW + 1) * 4, 16);      // win_row0
  n += 4 * align_up((size_t)P * 4, 16);         // tok_row, row_tok, row_pix, row_win
  n += align_up((size_t)P, 16);                 // tok_keep
  n += 2 * align_up((size_t)NW * 8, 16);        // tiles, tile_list
  return n;
}

extern "C" int sast_selection_bind(void* pool, int32_t B, int32_t NW, int32_t P, sast_selection* out) {
  (void)B;
  SAST_CHECK_PTR(pool); SAST_CHECK_PTR(out);
  if ((reinterpret_cast<uintptr_t>(pool) & 15) != 0) return SAST_E_SHAPE;
  char* p = (char*)pool;
  auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 16); return r; };
  out->counts = (int32_t*)take(8 * 4);
  out->win_K = (int32_t*)take((size_t)NW * 4);
  out->win_rank = (int32_t*)take((size_t)NW * 4);
  out->sel_win = (int32_t*)take((size_t)NW * 4);
  out->win_logit = (float*)take((size_t)NW * 4);
  out->win_row0 = (int32_t*)take(((size_t)NW + 1) * 4);
  out->tok_row = (int32_t*)take((size_t)P * 4);
  out->row_tok = (int32_t*)take((size_t)P * 4);
  out->row_pix = (int32_t*)take((size_t)P * 4);
  out->tok_keep = (uint8_t*)take((size_t)P);
  out->tiles = (int32_t*)take((size_t)NW * 8);
  out->tile_list = (int32_t*)take((size_t)NW * 8);
  out->row_win = (int32_t*)take((size_t)P * 4);
  return SAST_OK;
}

static int check_sel(const sast_selection& s) {
  SAST_CHECK_PTR(s.counts); SAST_CHECK_PTR(s.win_K); SAST_CHECK_PTR(s.win_rank); SAST_CHECK_PTR(s.win_row0);
  SAST_CHECK_PTR(s.sel_win); SAST_CHECK_PTR(s.tok_row); SAST_CHECK_PTR(s.row_tok); SAST_CHECK_PTR(s.row_pix);
  SAST_CHECK_PTR(s.win_logit); SAST_CHECK_PTR(s.tok_keep); SAST_CHECK_PTR(s.tiles); SAST_CHECK_PTR(s.tile_list); SAST_CHECK_PTR(s.row_win);
  return SAST_OK;
}

extern "C" int sast_select2(const sast_select_args* a, int32_t flavor_b, const sast_selection* sel_b, void* stream) {
  SAST_CHECK_PTR(a);
  const int nf = sel_b ? 2 : 1;
  for (int z = 0; z < nf; ++z) {
    const int fl = z == 0 ? a->flavor : flavor_b;
    if (fl != SAST_WINDOW && fl != SAST_GRID && fl != SAST_FLAT) return SAST_E_UNSUPPORTED;
    int rc = sast::check_geom(a->g, fl);
    if (rc) return rc;
    rc = check_sel(z == 0 ? a->sel : *sel_b);
    if (rc) return rc;
  }
  if (a->mode == SAST_SEL_SCORES) { SAST_CHECK_PTR(a->tok_score); }
  else if (a->mode == SAST_SEL_PROBS) { SAST_CHECK_PTR(a->win_prob); SAST_CHECK_PTR(a->tok_prob); }
  else if (a->mode == SAST_SEL_FLAGS) { SAST_CHECK_PTR(a->win_flag); SAST_CHECK_PTR(a->tok_flag); }
  else return SAST_E_UNSUPPORTED;
  const sast::Geom g = sast::make_geom(a->g, a->flavor);
  const size_t smem = ((size_t)g.N * 2 + 1) * 4;
  if (smem > 40 * 1024) return SAST_E_UNSUPPORTED;
  sast::SelectParams p;
  p.a = *a;
  p.sel_b = sel_b ? *sel_b : a->sel;
  p.flavor_b = sel_b ? flavor_b : a->flavor;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned wblocks = (unsigned)((g.NW + sast::kSelWarps - 1) / sast::kSelWarps);
  if (a->mode == SAST_SEL_SCORES) {
    sast::launch_k(sast::win_logit_kernel, dim3(wblocks, 1, nf), sast::kSelThreads, 0, st, p);
    SAST_LAUNCH_CHECK();
  }
  sast::launch_k(sast::select_flags_kernel, dim3((g.N + sast::kSelWarps - 1) / sast::kSelWarps, g.B, nf), sast::kSelThreads, 0, st, p);
  SAST_LAUNCH_CHECK();
  sast::launch_k(sast::select_scan_kernel, dim3(g.B, 1, nf), sast::kSelThreads, smem, st, p);
  SAST_LAUNCH_CHECK();
  sast::launch_k(sast::select_tokens_kernel, dim3(wblocks, 1, nf), sast::kSelThreads, 0, st, p);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

extern "C" int sast_select(const sast_select_args* a, void* stream) { return sast_select2(a, 0, nullptr, stream); }
