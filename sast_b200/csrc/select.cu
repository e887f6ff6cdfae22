// a5/a6: scene-adaptive window + token selection and index compaction
// (replaces SAST.py:84-96 window_selection/token_selection and :258-281
//  get_score_index_2d21d / get_score_index_with_padding, plus the isin() at :122/:147).
//
// Two launches, one CTA per frame each, no host round trip:
//   select_flags_kernel  : softmax + threshold (or thresholds on given probabilities, or
//                          given flags) -> keep flag per window and per token, K per window,
//                          per-frame totals.  Warp per window; token keep via __ballot_sync.
//   select_index_kernel  : exclusive prefix sums (ranks of kept windows, compacted row of each
//                          kept token) across frames and windows -> win_rank, sel_win,
//                          win_row0, tok_row, row_tok, counts{M,S,Kmax}.  Ballot/popc prefix
//                          inside a window, shuffle scans across windows.
// The compare is `prob >= thr` on fp32 with thr = fp32((1/N)/(1+BOUNCE)) exactly as torch
// evaluates `x >= d / (1 + b)`; ids come out ascending like torch.nonzero.
#include "common.cuh"

namespace sast {

constexpr int kSelThreads = 256;

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

__global__ void __launch_bounds__(kSelThreads) select_flags_kernel(sast_select_args a) {
  extern __shared__ float sm[];
  const Geom g = make_geom(a.g, a.flavor);
  float* wlogit = sm;                        // [N]
  int* wkeep = reinterpret_cast<int*>(sm + g.N);  // [N]
  __shared__ float red[kSelThreads / 32];
  __shared__ int tot[3];
  const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int HW = g.H * g.W;
  const float* fs = a.tok_score ? a.tok_score + (size_t)b * HW : nullptr;
  if (threadIdx.x < 3) tot[threadIdx.x] = 0;

  // ---- windows -------------------------------------------------------------------------
  if (a.mode == SAST_SEL_SCORES) {
    for (int n = wid; n < g.N; n += nwarp) {
      float s = 0.f;
      for (int t = lane; t < g.T; t += 32) s += fs[frame_pixel(n, t, g.H, g.W, g.p0, g.p1, a.flavor)];
      s = warp_sum(s);
      if (lane == 0) wlogit[n] = s / (float)g.T;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int n = threadIdx.x; n < g.N; n += blockDim.x) mx = fmaxf(mx, wlogit[n]);
    mx = block_reduce_max(mx, red);
    float se = 0.f;
    for (int n = threadIdx.x; n < g.N; n += blockDim.x) se += expf(wlogit[n] - mx);
    se = block_reduce_sum(se, red);
    for (int n = threadIdx.x; n < g.N; n += blockDim.x) {
      const float p = expf(wlogit[n] - mx) / se;
      wkeep[n] = p >= a.thr_win;
      if (a.win_prob_out) a.win_prob_out[(size_t)b * g.N + n] = p;
    }
  } else if (a.mode == SAST_SEL_PROBS) {
    for (int n = threadIdx.x; n < g.N; n += blockDim.x) wkeep[n] = a.win_prob[(size_t)b * g.N + n] >= a.thr_win;
  } else {
    for (int n = threadIdx.x; n < g.N; n += blockDim.x) wkeep[n] = a.win_flag[(size_t)b * g.N + n] != 0;
  }
  __syncthreads();

  // ---- tokens --------------------------------------------------------------------------
  int accM = 0, accS = 0, accK = 0;
  for (int n = wid; n < g.N; n += nwarp) {
    const int w = b * g.N + n;
    const size_t q0 = (size_t)w * g.T;
    if (!wkeep[n]) {
      for (int t = lane; t < g.T; t += 32) a.sel.tok_keep[q0 + t] = 0;
      if (lane == 0) { a.sel.win_K[w] = 0; a.sel.win_rank[w] = -1; }
      continue;
    }
    float v[4];
    bool keep[4];
    if (a.mode == SAST_SEL_SCORES) {
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = i * 32 + lane;
        v[i] = t < g.T ? fs[frame_pixel(n, t, g.H, g.W, g.p0, g.p1, a.flavor)] : -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
      mx = warp_max(mx);
      float se = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = (i * 32 + lane < g.T) ? expf(v[i] - mx) : 0.f;
        se += v[i];
      }
      se = warp_sum(se);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = i * 32 + lane;
        const float p = v[i] / se;
        keep[i] = t < g.T && p >= a.thr_tok;
        if (a.tok_prob_out && t < g.T) a.tok_prob_out[q0 + t] = p;
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
      if (t < g.T) a.sel.tok_keep[q0 + t] = keep[i] ? 1 : 0;
    }
    if (lane == 0) { a.sel.win_K[w] = K; a.sel.win_rank[w] = 0; }
    accM += 1; accS += K; accK = max(accK, K);
  }
  if (lane == 0) {
    atomicAdd(&tot[0], accM);
    atomicAdd(&tot[1], accS);
    atomicMax(&tot[2], accK);
  }
  __syncthreads();
  if (threadIdx.x < 3) a.sel.frame_tot[b * 4 + threadIdx.x] = tot[threadIdx.x];
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

__global__ void __launch_bounds__(kSelThreads) select_index_kernel(sast_select_args a) {
  const Geom g = make_geom(a.g, a.flavor);
  __shared__ int ws[kSelThreads / 32][2];
  __shared__ int base[2];
  const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    int m = 0, s = 0;
    for (int i = 0; i < b; ++i) { m += a.sel.frame_tot[i * 4]; s += a.sel.frame_tot[i * 4 + 1]; }
    base[0] = m; base[1] = s;
    if (b == g.B - 1) {
      int M = m + a.sel.frame_tot[b * 4], S = s + a.sel.frame_tot[b * 4 + 1], Kmax = 0;
      for (int i = 0; i < g.B; ++i) Kmax = max(Kmax, a.sel.frame_tot[i * 4 + 2]);
      a.sel.counts[0] = M; a.sel.counts[1] = S; a.sel.counts[2] = Kmax; a.sel.counts[3] = 0;
      a.sel.win_row0[g.NW] = S;
    }
  }
  __syncthreads();
  int run_m = base[0], run_s = base[1];
  for (int n0 = 0; n0 < g.N; n0 += blockDim.x) {
    const int n = n0 + threadIdx.x;
    const int w = b * g.N + n;
    const bool in = n < g.N;
    const int kept = in ? (a.sel.win_rank[w] >= 0) : 0;
    const int K = in ? a.sel.win_K[w] : 0;
    int em, es, tm, ts;
    block_excl_scan2(kept, K, em, es, tm, ts, ws);
    if (in) {
      a.sel.win_row0[w] = run_s + es;
      if (kept) {
        a.sel.win_rank[w] = run_m + em;
        a.sel.sel_win[run_m + em] = w;
      }
    }
    run_m += tm; run_s += ts;
    __syncthreads();
  }
  __syncthreads();
  for (int n = wid; n < g.N; n += nwarp) {
    const int w = b * g.N + n;
    const size_t q0 = (size_t)w * g.T;
    const int K = a.sel.win_K[w];
    const int row0 = a.sel.win_row0[w];
    int run = 0;
    for (int t0 = 0; t0 < g.T; t0 += 32) {
      const int t = t0 + lane;
      const bool keep = K > 0 && t < g.T && a.sel.tok_keep[q0 + t] != 0;
      const unsigned bal = __ballot_sync(kFull, keep);
      const int pre = __popc(bal & ((1u << lane) - 1u));
      if (t < g.T) a.sel.tok_row[q0 + t] = keep ? row0 + run + pre : -1;
      if (keep) a.sel.row_tok[row0 + run + pre] = (int)(q0 + t);
      run += __popc(bal);
    }
  }
}

}  // namespace sast

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t sast_selection_bytes(int32_t B, int32_t NW, int32_t P) {
  size_t n = 0;
  n += align_up(8 * 4, 16);                     // counts
  n += 3 * align_up((size_t)NW * 4, 16);        // win_K, win_rank, sel_win
  n += align_up(((size_t)NW + 1) * 4, 16);      // win_row0
  n += 2 * align_up((size_t)P * 4, 16);         // tok_row, row_tok
  n += align_up((size_t)B * 16, 16);            // frame_tot
  n += align_up((size_t)P, 16);                 // tok_keep
  n += align_up((size_t)NW * 16, 16);           // tiles
  return n;
}

extern "C" int sast_selection_bind(void* pool, int32_t B, int32_t NW, int32_t P, sast_selection* out) {
  SAST_CHECK_PTR(pool); SAST_CHECK_PTR(out);
  if ((reinterpret_cast<uintptr_t>(pool) & 15) != 0) return SAST_E_SHAPE;
  char* p = (char*)pool;
  auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 16); return r; };
  out->counts = (int32_t*)take(8 * 4);
  out->win_K = (int32_t*)take((size_t)NW * 4);
  out->win_rank = (int32_t*)take((size_t)NW * 4);
  out->sel_win = (int32_t*)take((size_t)NW * 4);
  out->win_row0 = (int32_t*)take(((size_t)NW + 1) * 4);
  out->tok_row = (int32_t*)take((size_t)P * 4);
  out->row_tok = (int32_t*)take((size_t)P * 4);
  out->frame_tot = (int32_t*)take((size_t)B * 16);
  out->tok_keep = (uint8_t*)take((size_t)P);
  out->tiles = (int32_t*)take((size_t)NW * 16);
  return SAST_OK;
}

extern "C" int sast_select(const sast_select_args* a, void* stream) {
  SAST_CHECK_PTR(a);
  if (a->flavor != SAST_WINDOW && a->flavor != SAST_GRID && a->flavor != SAST_FLAT) return SAST_E_UNSUPPORTED;
  int rc = sast::check_geom(a->g, a->flavor);
  if (rc) return rc;
  if (a->mode == SAST_SEL_SCORES) { SAST_CHECK_PTR(a->tok_score); }
  else if (a->mode == SAST_SEL_PROBS) { SAST_CHECK_PTR(a->win_prob); SAST_CHECK_PTR(a->tok_prob); }
  else if (a->mode == SAST_SEL_FLAGS) { SAST_CHECK_PTR(a->win_flag); SAST_CHECK_PTR(a->tok_flag); }
  else return SAST_E_UNSUPPORTED;
  const sast_selection& s = a->sel;
  SAST_CHECK_PTR(s.counts); SAST_CHECK_PTR(s.win_K); SAST_CHECK_PTR(s.win_rank); SAST_CHECK_PTR(s.win_row0);
  SAST_CHECK_PTR(s.sel_win); SAST_CHECK_PTR(s.tok_row); SAST_CHECK_PTR(s.row_tok); SAST_CHECK_PTR(s.frame_tot);
  SAST_CHECK_PTR(s.tok_keep);
  const sast::Geom g = sast::make_geom(a->g, a->flavor);
  const size_t smem = (size_t)g.N * 8;
  if (smem > 40 * 1024) return SAST_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  sast::select_flags_kernel<<<g.B, sast::kSelThreads, smem, st>>>(*a);
  SAST_LAUNCH_CHECK();
  sast::select_index_kernel<<<g.B, sast::kSelThreads, 0, st>>>(*a);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}
