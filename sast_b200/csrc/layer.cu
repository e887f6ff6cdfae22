// a8-a13: one MS-WSA layer  (replaces MS_WSA.forward, SAST.py:199-255).
//
// Pipeline (all counts device resident; every kernel is launched on a worst-case grid and
// CTAs beyond the selected-token count S exit immediately, so nothing syncs with the host):
//   gather_ln   : one pass over the map in partitioned order.  LN1 for every token;
//                 unselected tokens are written straight to `out` (they keep norm1(x),
//                 SAST.py:251-254), selected tokens get LN2 and land in the compacted
//                 [S,C] buffer (fp32 shortcut + bf16 operand copy).       (SAST.py:206-216)
//   qkv         : [S,C] x [3C,C]^T                                          (SAST.py:219)
//   attention   : per selected window, softmax(q k^T/sqrt(32)) v over the window's
//                 selected tokens only -- the compacted buffer holds no padding, so the
//                 reference's -1e4 column mask (SAST.py:223-226) has nothing to mask.
//   proj        : y = n2 + g1 * (o Wp^T + b)                                (SAST.py:230-234)
//   glu         : hid = val * gelu(gate)                                    (ops.py:135-137)
//   mlp out     : map[pixel(row)] = y + g2 * (hid W2^T + b)                 (SAST.py:237,248-254)
//   (context broadcast, SAST.py:240-246, adds a per-frame mean between the last two.)
// This file holds the orchestration and the fp32 CUDA-core kernels (SAST_FP32, validation
// grade); the tcgen05 kernels of SAST_BF16 live in gemm_tc.cu / attn_tc.cu.
#include "layer.cuh"
#include <cstdlib>

namespace sast {

int launch_attention_f32(const float* qkv, float* att, int C, int heads, int T, int NW, const sast_selection& sel, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// gather + LN1 (+ LN2)
// ------------------------------------------------------------------------------------------
template <int LPT>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPT / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// LPT lanes cooperate on one token (NV float4 each: C <= 4*LPT*NV); every lane group handles TPW
// tokens per pass with all of their loads issued before the first use (memory-level parallelism
// is what this HBM-bound kernel lives on).
template <int LPT, int NV, int TPW>
__global__ void __launch_bounds__(256) gather_ln_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                        const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ w2, const float* __restrict__ b2,
                                                        float eps, const int* __restrict__ tok_row, Geom g, int flavor,
                                                        float* __restrict__ n2f, __nv_bfloat16* __restrict__ n2h) {
  pdl_entry();
  constexpr int GROUPS = 32 / LPT;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPT, l = lane % LPT;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long q0 = (warp * GROUPS + sub) * TPW;
  const int C = g.C;
  const float inv_c = 1.0f / (float)C;

  float4 v[TPW][NV];
  long long pix[TPW];
  int row[TPW];
#pragma unroll
  for (int j = 0; j < TPW; ++j) {
    const long long q = q0 + j;
    const bool ok = q < g.P;
    pix[j] = ok ? token_pixel(q, g, flavor) : 0;
    row[j] = ok ? tok_row[q] : -1;
    const float* xp = x + pix[j] * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (l + LPT * i) * 4;
      v[j][i] = (ok && c < C) ? *reinterpret_cast<const float4*>(xp + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int j = 0; j < TPW; ++j) {
    const bool ok = q0 + j < g.P;          // uniform inside the lane group
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
    float mean = group_sum<LPT>(s) * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (l + LPT * i) * 4;
      if (c < C) {
        const float a = v[j][i].x - mean, b = v[j][i].y - mean, cc = v[j][i].z - mean, d = v[j][i].w - mean;
        ss += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    float rstd = rsqrtf(group_sum<LPT>(ss) * inv_c + eps);
    s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (l + LPT * i) * 4;
      if (c < C) {
        const float4 gw = *reinterpret_cast<const float4*>(w1 + c);
        const float4 gb = *reinterpret_cast<const float4*>(b1 + c);
        v[j][i].x = (v[j][i].x - mean) * rstd * gw.x + gb.x;
        v[j][i].y = (v[j][i].y - mean) * rstd * gw.y + gb.y;
        v[j][i].z = (v[j][i].z - mean) * rstd * gw.z + gb.z;
        v[j][i].w = (v[j][i].w - mean) * rstd * gw.w + gb.w;
        s += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
      }
    }
    // second norm only matters for selected tokens, but the shuffles must be executed by every lane
    mean = group_sum<LPT>(s) * inv_c;
    ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (l + LPT * i) * 4;
      if (c < C) {
        const float a = v[j][i].x - mean, b = v[j][i].y - mean, cc = v[j][i].z - mean, d = v[j][i].w - mean;
        ss += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    rstd = rsqrtf(group_sum<LPT>(ss) * inv_c + eps);
    if (!ok) continue;
    if (row[j] < 0) {   // unselected: keeps norm1(x)   (SAST.py:251-254)
      if (out == nullptr) continue;                       // recompute pass of the backward: only the compacted rows matter
      float* op = out + pix[j] * C;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (l + LPT * i) * 4;
        if (c < C) *reinterpret_cast<float4*>(op + c) = v[j][i];
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (l + LPT * i) * 4;
      if (c < C) {
        const float4 gw = *reinterpret_cast<const float4*>(w2 + c);
        const float4 gb = *reinterpret_cast<const float4*>(b2 + c);
        float4 o;
        o.x = (v[j][i].x - mean) * rstd * gw.x + gb.x;
        o.y = (v[j][i].y - mean) * rstd * gw.y + gb.y;
        o.z = (v[j][i].z - mean) * rstd * gw.z + gb.z;
        o.w = (v[j][i].w - mean) * rstd * gw.w + gb.w;
        *reinterpret_cast<float4*>(n2f + (size_t)row[j] * C + c) = o;
        if (n2h) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(n2h + (size_t)row[j] * C + c) = pk;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// fp32 CUDA-core GEMM  D[M,N] = A[M,K] W[N,K]^T  with fused epilogues (SAST_FP32 path)
// ------------------------------------------------------------------------------------------
constexpr int GM = 64, GN = 64, GK = 16, GPAD = 4;

template <int EPI>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                       const float* __restrict__ bias, int N, int K,
                                                       const int* __restrict__ counts, EpiParams ep) {
  pdl_entry();
  __shared__ __align__(16) float As[GK][GM + GPAD];
  __shared__ __align__(16) float Bs[GK][GN + GPAD];
  const int M = counts[1];
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  if (m0 >= M) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), w = a;
    if (m0 + lrow < M) a = *reinterpret_cast<const float4*>(A + (size_t)(m0 + lrow) * lda + k0 + lk);
    if (n0 + lrow < N) w = *reinterpret_cast<const float4*>(W + (size_t)(n0 + lrow) * K + k0 + lk);
    As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
    Bs[lk + 0][lrow] = w.x; Bs[lk + 1][lrow] = w.y; Bs[lk + 2][lrow] = w.z; Bs[lk + 3][lrow] = w.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float am[4] = {av.x, av.y, av.z, av.w};
      const float bn[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= N) return;
  float bb[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) { const float4 t = *reinterpret_cast<const float4*>(bias + n); bb[0] = t.x; bb[1] = t.y; bb[2] = t.z; bb[3] = t.w; }
  float gg[4] = {1.f, 1.f, 1.f, 1.f};
  if ((EPI == EPI_RESID || EPI == EPI_SCATTER) && ep.gamma) {
    const float4 t = *reinterpret_cast<const float4*>(ep.gamma + n); gg[0] = t.x; gg[1] = t.y; gg[2] = t.z; gg[3] = t.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bb[j];
    if (EPI == EPI_STORE) {
      *reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * ep.ldo + n) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (EPI == EPI_GLU) {
      *reinterpret_cast<float2*>(ep.out_f32 + (size_t)row * ep.ldo + n / 2) =
          make_float2(v[0] * gelu_erf(v[1]), v[2] * gelu_erf(v[3]));
    } else {
      const float4 r = *reinterpret_cast<const float4*>(ep.resid + (size_t)row * ep.ldr + n);
      const float4 o = make_float4(r.x + gg[0] * v[0], r.y + gg[1] * v[1], r.z + gg[2] * v[2], r.w + gg[3] * v[3]);
      if (EPI == EPI_RESID) {
        *reinterpret_cast<float4*>(ep.out_f32 + (size_t)row * ep.ldo + n) = o;
      } else {
        const long long pix = ep.row_pix[row];
        *reinterpret_cast<float4*>(ep.out_f32 + pix * ep.C + n) = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// fp32 attention over the compacted rows of one window and one head (SAST_FP32 path)
// ------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(128) attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ att, int C,
                                                            const int* __restrict__ win_K, const int* __restrict__ win_row0) {
  pdl_entry();
  extern __shared__ __align__(16) float kv[];      // k [K][DH] then v [K][DH]
  const int w = blockIdx.x, h = blockIdx.y;
  const int K = win_K[w];
  if (K == 0) return;
  const int row0 = win_row0[w];
  const int ld = 3 * C;
  constexpr int Q4 = DH / 4;                       // float4 per row of q, k or v (head-major qkv rows: [h][q,k,v][DH])
  float* ks = kv;
  float* vs = kv + (size_t)K * DH;
  for (int i = threadIdx.x; i < K * 2 * Q4; i += blockDim.x) {
    const int r = i / (2 * Q4), c = i % (2 * Q4);
    const float4 t = *reinterpret_cast<const float4*>(qkv + (size_t)(row0 + r) * ld + h * 3 * DH + DH + c * 4);
    *reinterpret_cast<float4*>((c < Q4 ? ks + r * DH + c * 4 : vs + r * DH + (c - Q4) * 4)) = t;
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= K) return;
  float q[DH], o[DH];
  const float scale = rsqrtf((float)DH);           // dim_head ** -0.5 (SAST.py:176)
#pragma unroll
  for (int c = 0; c < Q4; ++c) {
    const float4 t = *reinterpret_cast<const float4*>(qkv + (size_t)(row0 + i) * ld + h * 3 * DH + c * 4);
    q[c * 4] = t.x; q[c * 4 + 1] = t.y; q[c * 4 + 2] = t.z; q[c * 4 + 3] = t.w;
  }
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
  float mx = -INFINITY, l = 0.f;
  for (int j = 0; j < K; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) s = fmaf(q[d], ks[j * DH + d], s);
    s *= scale;
    const float mn = fmaxf(mx, s);
    const float corr = expf(mx - mn), p = expf(s - mn);
    l = l * corr + p;
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = fmaf(p, vs[j * DH + d], o[d] * corr);
    mx = mn;
  }
  const float il = 1.0f / l;
#pragma unroll
  for (int c = 0; c < Q4; ++c)
    *reinterpret_cast<float4*>(att + (size_t)(row0 + i) * C + h * DH + c * 4) =
        make_float4(o[c * 4] * il, o[c * 4 + 1] * il, o[c * 4 + 2] * il, o[c * 4 + 3] * il);
}

// ------------------------------------------------------------------------------------------
// context broadcast (SAST.py:240-246): m <- 0.5 m + 0.5 * (sum of the frame's selected m) / (N*T)
// ------------------------------------------------------------------------------------------
__global__ void cb_mean_kernel(const float* __restrict__ m, int C, const int* __restrict__ win_row0, int N, float inv_nt,
                               float* __restrict__ mean) {
  pdl_entry();
  __shared__ float red[8][32];
  const int b = blockIdx.x, c = blockIdx.y * 32 + threadIdx.x;
  const int r0 = win_row0[b * N], r1 = win_row0[(b + 1) * N];
  float s = 0.f;
  if (c < C)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += m[(size_t)r * C + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    mean[(size_t)b * C + c] = t * inv_nt;
  }
}

__global__ void cb_scatter_kernel(const float* __restrict__ m, const float* __restrict__ y, const float* __restrict__ mean,
                                  const float* __restrict__ gamma, const int* __restrict__ counts,
                                  const int* __restrict__ row_tok, const int* __restrict__ row_pix, Geom g,
                                  float* __restrict__ out) {
  pdl_entry();
  const int S = counts[1];
  const int c4 = g.C / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)S * c4; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / c4), c = (int)(i % c4) * 4;
    const int q = row_tok[row];
    const int b = (q / g.T) / g.N;
    const float4 mv = *reinterpret_cast<const float4*>(m + (size_t)row * g.C + c);
    const float4 yv = *reinterpret_cast<const float4*>(y + (size_t)row * g.C + c);
    const float4 mu = *reinterpret_cast<const float4*>(mean + (size_t)b * g.C + c);
    float4 gm = make_float4(1.f, 1.f, 1.f, 1.f);
    if (gamma) gm = *reinterpret_cast<const float4*>(gamma + c);
    const long long pix = row_pix[row];
    *reinterpret_cast<float4*>(out + pix * g.C + c) =
        make_float4(yv.x + gm.x * (0.5f * mv.x + 0.5f * mu.x), yv.y + gm.y * (0.5f * mv.y + 0.5f * mu.y),
                    yv.z + gm.z * (0.5f * mv.z + 0.5f * mu.z), yv.w + gm.w * (0.5f * mv.w + 0.5f * mu.w));
  }
}

// ------------------------------------------------------------------------------------------
// standalone gather / scatter of selected rows (tests + HBM microbenchmark)
// ------------------------------------------------------------------------------------------
template <bool GATHER, int LPT, int NV>
__global__ void __launch_bounds__(256) rows_copy_kernel(float* __restrict__ map, float* __restrict__ rows,
                                                        const int* __restrict__ counts, const int* __restrict__ row_pix, int C) {
  pdl_entry();
  constexpr int GROUPS = 32 / LPT, RPG = 4;              // rows per lane group per pass
  const int S = counts[1];
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPT, l = lane % LPT;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r0 = (warp * GROUPS + sub) * RPG; r0 < S; r0 += nwarps * GROUPS * RPG) {
    long long pix[RPG];
#pragma unroll
    for (int j = 0; j < RPG; ++j) pix[j] = r0 + j < S ? row_pix[r0 + j] : -1;
    float4 v[RPG][NV];
#pragma unroll
    for (int j = 0; j < RPG; ++j) {
      const float* src = GATHER ? map + pix[j] * C : rows + (r0 + j) * C;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (l + LPT * i) * 4;
        if (pix[j] >= 0 && c < C) v[j][i] = *reinterpret_cast<const float4*>(src + c);
      }
    }
#pragma unroll
    for (int j = 0; j < RPG; ++j) {
      float* dst = GATHER ? rows + (r0 + j) * C : map + pix[j] * C;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = (l + LPT * i) * 4;
        if (pix[j] >= 0 && c < C) *reinterpret_cast<float4*>(dst + c) = v[j][i];
      }
    }
  }
}

template <bool GATHER>
static int launch_rows_copy(float* map, float* rows, const sast_selection* sel, int C, cudaStream_t st) {
  // a row is read by consecutive lanes (whole 128-byte lines per request), 4 rows per lane group in flight
  // (measured: wider per-lane strips -- 4 lanes x 64 B per row -- lose ~7 % of bandwidth to partial lines)
  const dim3 grid(148 * 8), block(256);
  if (C <= 32) sast::launch_k(rows_copy_kernel<GATHER, 8, 1>, grid, block, 0, st, map, rows, sel->counts, sel->row_pix, C);
  else if (C <= 64) sast::launch_k(rows_copy_kernel<GATHER, 16, 1>, grid, block, 0, st, map, rows, sel->counts, sel->row_pix, C);
  else if (C <= 96) sast::launch_k(rows_copy_kernel<GATHER, 32, 1>, grid, block, 0, st, map, rows, sel->counts, sel->row_pix, C);
  else if (C <= 128) sast::launch_k(rows_copy_kernel<GATHER, 32, 1>, grid, block, 0, st, map, rows, sel->counts, sel->row_pix, C);
  else if (C <= 256) sast::launch_k(rows_copy_kernel<GATHER, 32, 2>, grid, block, 0, st, map, rows, sel->counts, sel->row_pix, C);
  else if (C <= 512) sast::launch_k(rows_copy_kernel<GATHER, 32, 4>, grid, block, 0, st, map, rows, sel->counts, sel->row_pix, C);
  else if (C <= 1024) sast::launch_k(rows_copy_kernel<GATHER, 32, 8>, grid, block, 0, st, map, rows, sel->counts, sel->row_pix, C);
  else return SAST_E_UNSUPPORTED;
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

// ------------------------------------------------------------------------------------------
size_t layer_workspace_layout(long long P, int C, int I, int B, int precision, void* base, LayerWorkspace* ws) {
  const size_t e = precision == SAST_BF16 ? 2 : 4;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return p ? (void*)(p + o) : nullptr; };
  void* n2f = take((size_t)P * C * 4);
  void* n2h = take(precision == SAST_BF16 ? (size_t)P * C * 2 : 0);
  void* qkv = take((size_t)P * 3 * C * e);
  void* att = take((size_t)P * C * e);
  void* yf = take((size_t)P * C * 4);
  void* yh = take(precision == SAST_BF16 ? (size_t)P * C * 2 : 0);
  void* hid = take((size_t)P * I * e);
  void* mtmp = take((size_t)P * C * 4);
  void* cbm = take((size_t)B * C * 4);
  if (ws) {
    ws->n2f = (float*)n2f; ws->n2h = precision == SAST_BF16 ? (__nv_bfloat16*)n2h : nullptr;
    ws->qkv = qkv; ws->att = att; ws->yf = (float*)yf; ws->yh = precision == SAST_BF16 ? (__nv_bfloat16*)yh : nullptr;
    ws->hid = hid; ws->mtmp = (float*)mtmp; ws->cbmean = (float*)cbm;
  }
  return off;
}

static int attention_variant() {
  static int v = -1;     // read once; debug / A-B knob only
  if (v < 0) {
    const char* e = getenv("SAST_B200_ATTN");
    v = (e && e[0] == 's') ? 1 : 0;
  }
  return v;
}

template <int EPI>
static int launch_gemm_f32(const float* A, int lda, const float* W, const float* bias, int N, int K, const int* counts,
                           long long max_rows, const EpiParams& ep, cudaStream_t st) {
  const dim3 grid((unsigned)((max_rows + GM - 1) / GM), (unsigned)((N + GN - 1) / GN));
  sast::launch_k(gemm_f32_kernel<EPI>, grid, 256, 0, st, A, lda, W, bias, N, K, counts, ep);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

template <int LPT, int NV, int TPW>
static int launch_gather_ln_t(const sast_layer_args& a, const Geom& g, const LayerWorkspace& ws, cudaStream_t st) {
  const long long tok_per_cta = 8ll * (32 / LPT) * TPW;
  const unsigned grid = (unsigned)((g.P + tok_per_cta - 1) / tok_per_cta);
  sast::launch_k(gather_ln_kernel<LPT, NV, TPW>, grid, 256, 0, st, a.x, a.out, a.w.ln1_w, a.w.ln1_b, a.w.ln2_w, a.w.ln2_b, a.w.ln_eps,
                                                        a.sel.tok_row, g, a.flavor, ws.n2f, ws.n2h);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

static int launch_gather_ln(const sast_layer_args& a, const Geom& g, const LayerWorkspace& ws, cudaStream_t st) {
  // few lanes per token (each holding up to 4 float4): short shuffle reductions, and the per-token index
  // arithmetic (token -> pixel) is amortised over 8-16 tokens per warp
  const int C = g.C;
  if (C <= 32) return launch_gather_ln_t<2, 4, 2>(a, g, ws, st);
  if (C <= 64) return launch_gather_ln_t<4, 4, 2>(a, g, ws, st);
  if (C <= 96) return launch_gather_ln_t<8, 3, 2>(a, g, ws, st);
  if (C <= 128) return launch_gather_ln_t<8, 4, 2>(a, g, ws, st);
  if (C <= 256) return launch_gather_ln_t<16, 4, 2>(a, g, ws, st);
  if (C <= 512) return launch_gather_ln_t<32, 4, 2>(a, g, ws, st);
  if (C <= 1024) return launch_gather_ln_t<32, 8, 1>(a, g, ws, st);
  return SAST_E_UNSUPPORTED;
}

// pieces of the fp32 forward reused by the recompute pass of sast_layer_bwd (layer_bwd.cu)
int launch_gather_ln_f32(const float* x, float* out_unselected, const sast_layer_weights& w, const sast_selection& sel, const Geom& g,
                         int flavor, float* n2f, cudaStream_t st) {
  sast_layer_args a{};
  a.x = x; a.out = out_unselected; a.w = w; a.sel = sel; a.flavor = flavor;
  LayerWorkspace ws{};
  ws.n2f = n2f; ws.n2h = nullptr;
  return launch_gather_ln(a, g, ws, st);
}
int launch_attention_f32(const float* qkv, float* att, int C, int heads, int T, int NW, const sast_selection& sel, cudaStream_t st) {
  const int dh = C / heads;
  const size_t smem = (size_t)T * 2 * dh * sizeof(float);
  const dim3 grid(NW, heads);
  switch (dh) {
    case 32: sast::launch_k(attention_f32_kernel<32>, grid, 128, smem, st, qkv, att, C, sel.win_K, sel.win_row0); break;
    case 24: sast::launch_k(attention_f32_kernel<24>, grid, 128, smem, st, qkv, att, C, sel.win_K, sel.win_row0); break;
    case 16: sast::launch_k(attention_f32_kernel<16>, grid, 128, smem, st, qkv, att, C, sel.win_K, sel.win_row0); break;
    case 8: sast::launch_k(attention_f32_kernel<8>, grid, 128, smem, st, qkv, att, C, sel.win_K, sel.win_row0); break;
    default: return SAST_E_UNSUPPORTED;
  }
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}
int launch_rows_gather(const float* map, float* rows, const sast_selection* sel, int C, cudaStream_t st) {
  return launch_rows_copy<true>(const_cast<float*>(map), rows, sel, C, st);
}

}  // namespace sast

extern "C" size_t sast_layer_workspace_bytes(int64_t P, int32_t C, int32_t I, int32_t B, int32_t precision) {
  if (precision == SAST_BF16_CHAIN) precision = SAST_BF16;
  const size_t chain = sast::layer_workspace_layout(P, C, I, B, precision, nullptr, nullptr);
  const size_t grp = precision == SAST_BF16 ? sast::group_layer_workspace_bytes(P, C, I) : 0;
  return chain > grp ? chain : grp;
}

extern "C" int32_t sast_layer_is_fused(int64_t P, int32_t C, int32_t I, int32_t precision, int32_t enable_cb) {
  sast_layer_args a{};
  a.g.B = 1; a.g.H = 1; a.g.W = (int32_t)(P > 0x7fffffff ? 0x7fffffff : P);
  a.g.C = C; a.w.I = I; a.precision = precision; a.enable_cb = enable_cb;
  return sast::fused_layer_supported(a) ? 1 : 0;
}

extern "C" int sast_layer_fwd(const sast_layer_args* ap, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(ap);
  sast_layer_args a_ = *ap;
  const bool chain = a_.precision == SAST_BF16_CHAIN;
  if (chain) a_.precision = SAST_BF16;
  const sast_layer_args& a = a_;
  SAST_CHECK_PTR(a.x); SAST_CHECK_PTR(a.out);
  if (a.x == a.out) return SAST_E_UNSUPPORTED;
  int rc = check_geom(a.g, a.flavor);
  if (rc) return rc;
  const sast_layer_weights& w = a.w;
  SAST_CHECK_PTR(w.ln1_w); SAST_CHECK_PTR(w.ln1_b); SAST_CHECK_PTR(w.ln2_w); SAST_CHECK_PTR(w.ln2_b);
  SAST_CHECK_PTR(w.qkv_w); SAST_CHECK_PTR(w.proj_w); SAST_CHECK_PTR(w.mlp1_w); SAST_CHECK_PTR(w.mlp2_w);
  SAST_CHECK_PTR(a.sel.counts); SAST_CHECK_PTR(a.sel.tok_row); SAST_CHECK_PTR(a.sel.row_tok); SAST_CHECK_PTR(a.sel.row_pix);
  SAST_CHECK_PTR(a.sel.win_K); SAST_CHECK_PTR(a.sel.win_row0); SAST_CHECK_PTR(a.sel.tiles);
  const Geom g = make_geom(a.g, a.flavor);
  const int C = g.C, I = w.I;
  const int dh = w.dim_head > 0 ? w.dim_head : 32;
  if (C % 8 != 0 || I % 8 != 0 || I <= 0 || C > 1024 || C % dh != 0) return SAST_E_SHAPE;
  if (a.precision != SAST_FP32 && a.precision != SAST_BF16) return SAST_E_UNSUPPORTED;
  if (dh != 32 && (a.precision != SAST_FP32 || (dh != 8 && dh != 16 && dh != 24))) return SAST_E_UNSUPPORTED;   // tcgen05 kernels: dim_head 32
  if (a.precision == SAST_BF16 && (C % 32 != 0 || I % 32 != 0)) return SAST_E_SHAPE;
  if (a.precision == SAST_BF16) {
    SAST_CHECK_PTR(w.qkv_w_bf16); SAST_CHECK_PTR(w.proj_w_bf16); SAST_CHECK_PTR(w.mlp1_w_bf16); SAST_CHECK_PTR(w.mlp2_w_bf16);
  }
  if (!chain && fused_layer_supported(a)) return launch_layer_fused(a, g, (cudaStream_t)stream);      // one kernel, no workspace
  SAST_CHECK_PTR(a.workspace);
  LayerWorkspace ws;
  const size_t need = layer_workspace_layout(g.P, C, I, g.B, a.precision, a.workspace, &ws);
  if (need > a.workspace_bytes) return SAST_E_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(a.workspace) & 255) != 0) return SAST_E_SHAPE;
  if (!chain && group_layer_supported(a)) return launch_layer_group(a, g, (cudaStream_t)stream);         // one kernel, CTA groups
  cudaStream_t st = (cudaStream_t)stream;
  const int heads = C / dh;

  rc = launch_gather_ln(a, g, ws, st);
  if (rc) return rc;

  EpiParams ep{};
  ep.C = C; ep.row_pix = a.sel.row_pix;

  if (a.precision == SAST_FP32) {
    // qkv
    ep.out_f32 = (float*)ws.qkv; ep.ldo = 3 * C;
    rc = launch_gemm_f32<EPI_STORE>(ws.n2f, C, w.qkv_w, w.qkv_b, 3 * C, C, a.sel.counts, g.P, ep, st);
    if (rc) return rc;
    // attention
    rc = launch_attention_f32((const float*)ws.qkv, (float*)ws.att, C, heads, g.T, g.NW, a.sel, st);
    if (rc) return rc;
    // proj + LayerScale + shortcut
    ep.out_f32 = ws.yf; ep.ldo = C; ep.resid = ws.n2f; ep.ldr = C; ep.gamma = w.gamma1;
    rc = launch_gemm_f32<EPI_RESID>((const float*)ws.att, C, w.proj_w, w.proj_b, C, C, a.sel.counts, g.P, ep, st);
    if (rc) return rc;
    // GLU
    ep.out_f32 = (float*)ws.hid; ep.ldo = I; ep.resid = nullptr; ep.gamma = nullptr;
    rc = launch_gemm_f32<EPI_GLU>(ws.yf, C, w.mlp1_w, w.mlp1_b, 2 * I, C, a.sel.counts, g.P, ep, st);
    if (rc) return rc;
    // MLP out (+ residual + scatter-back, or context broadcast)
    if (!a.enable_cb) {
      ep.out_f32 = a.out; ep.resid = ws.yf; ep.ldr = C; ep.gamma = w.gamma2;
      rc = launch_gemm_f32<EPI_SCATTER>((const float*)ws.hid, I, w.mlp2_w, w.mlp2_b, C, I, a.sel.counts, g.P, ep, st);
      if (rc) return rc;
    } else {
      ep.out_f32 = ws.mtmp; ep.ldo = C;
      rc = launch_gemm_f32<EPI_STORE>((const float*)ws.hid, I, w.mlp2_w, w.mlp2_b, C, I, a.sel.counts, g.P, ep, st);
      if (rc) return rc;
    }
  } else {
    ep.out_bf16 = (__nv_bfloat16*)ws.qkv; ep.ldo = 3 * C;
    rc = launch_gemm_tc(ws.n2h, C, (const __nv_bfloat16*)w.qkv_w_bf16, w.qkv_b, 3 * C, C, a.sel.counts, g.P, EPI_STORE, ep, st);
    if (rc) return rc;
    rc = launch_attention_tc((const __nv_bfloat16*)ws.qkv, (__nv_bfloat16*)ws.att, C, a.sel, g.NW, g.B, g.T, g.P, attention_variant(), st);
    if (rc) return rc;
    ep.out_f32 = ws.yf; ep.out_bf16 = ws.yh; ep.ldo = C; ep.resid = ws.n2f; ep.ldr = C; ep.gamma = w.gamma1;
    rc = launch_gemm_tc((const __nv_bfloat16*)ws.att, C, (const __nv_bfloat16*)w.proj_w_bf16, w.proj_b, C, C, a.sel.counts, g.P,
                        EPI_RESID, ep, st);
    if (rc) return rc;
    ep.out_f32 = nullptr; ep.out_bf16 = (__nv_bfloat16*)ws.hid; ep.ldo = I; ep.resid = nullptr; ep.gamma = nullptr;
    rc = launch_gemm_tc(ws.yh, C, (const __nv_bfloat16*)w.mlp1_w_bf16, w.mlp1_b, 2 * I, C, a.sel.counts, g.P, EPI_GLU, ep, st);
    if (rc) return rc;
    ep.out_bf16 = nullptr; ep.resid = ws.yf; ep.ldr = C;
    if (!a.enable_cb) {
      ep.out_f32 = a.out; ep.gamma = w.gamma2;
      rc = launch_gemm_tc((const __nv_bfloat16*)ws.hid, I, (const __nv_bfloat16*)w.mlp2_w_bf16, w.mlp2_b, C, I, a.sel.counts, g.P,
                          EPI_SCATTER, ep, st);
    } else {
      ep.out_f32 = ws.mtmp; ep.ldo = C; ep.resid = nullptr;
      rc = launch_gemm_tc((const __nv_bfloat16*)ws.hid, I, (const __nv_bfloat16*)w.mlp2_w_bf16, w.mlp2_b, C, I, a.sel.counts, g.P,
                          EPI_STORE, ep, st);
    }
    if (rc) return rc;
  }
  if (a.enable_cb) {
    sast::launch_k(cb_mean_kernel, dim3(g.B, (C + 31) / 32), dim3(32, 8), 0, st, ws.mtmp, C, a.sel.win_row0, g.N, 1.0f / (float)(g.N * g.T), ws.cbmean);
    SAST_LAUNCH_CHECK();
    sast::launch_k(cb_scatter_kernel, 148 * 8, 256, 0, st, ws.mtmp, ws.yf, ws.cbmean, w.gamma2, a.sel.counts, a.sel.row_tok, a.sel.row_pix, g, a.out);
    SAST_LAUNCH_CHECK();
  }
  return SAST_OK;
}

extern "C" int sast_gather(const sast_geom* gp, int32_t flavor, const float* x, const sast_selection* sel, float* rows, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(gp); SAST_CHECK_PTR(x); SAST_CHECK_PTR(sel); SAST_CHECK_PTR(rows);
  SAST_CHECK_PTR(sel->counts); SAST_CHECK_PTR(sel->row_pix);
  int rc = check_geom(*gp, flavor);
  if (rc) return rc;
  if (gp->C % 4 != 0) return SAST_E_SHAPE;
  return launch_rows_copy<true>(const_cast<float*>(x), rows, sel, gp->C, (cudaStream_t)stream);
}

extern "C" int sast_scatter(const sast_geom* gp, int32_t flavor, const float* rows, const sast_selection* sel, float* x, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(gp); SAST_CHECK_PTR(x); SAST_CHECK_PTR(sel); SAST_CHECK_PTR(rows);
  SAST_CHECK_PTR(sel->counts); SAST_CHECK_PTR(sel->row_pix);
  int rc = check_geom(*gp, flavor);
  if (rc) return rc;
  if (gp->C % 4 != 0) return SAST_E_SHAPE;
  return launch_rows_copy<false>(x, const_cast<float*>(rows), sel, gp->C, (cudaStream_t)stream);
}
