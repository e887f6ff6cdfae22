// Backward of the hot path for the training configuration (BASELINE.json config 4; the reference trains through
// stock autograd, train.py / modules/detection.py:113-221):
//   sast_layer_bwd : gradient of one MS-WSA layer (SAST.py:199-255) w.r.t. its input map and every parameter;
//   sast_score_bwd : gradient of the scoring module's STP weighting (SAST.py:105-114) w.r.t. x, to_scores, to_controls.
// Selection is not differentiable in the reference either (the index tensors carry no gradient; to_scores / to_controls
// learn only through the STP weight), so both take the forward's selection as a constant.
//
// Recompute-based: nothing is saved by the forward; the backward re-runs the layer in fp32 on the compacted rows
// (CUDA-core kernels of layer.cu) keeping the intermediates it needs, then walks the chain backwards:
//   d_out -> (LayerScale 2, MLP out) -> GLU -> (MLP in) -> (LayerScale 1, proj) -> attention -> QKV -> LN2 -> LN1.
// fp32 throughout; weight gradients are reductions over the S selected rows (split over CTAs, fp32 atomics into
// zero-initialised buffers: summation order varies at the 1e-7 level from run to run).
#include "layer.cuh"
#include <cstdlib>

namespace sast {

size_t layer_workspace_layout(long long P, int C, int I, int B, int precision, void* base, LayerWorkspace* ws);
// forward pieces of layer.cu reused for the recompute
int launch_gather_ln_f32(const float* x, float* out_unselected, const sast_layer_weights& w, const sast_selection& sel, const Geom& g,
                         int flavor, float* n2f, cudaStream_t st);
int launch_attention_f32(const float* qkv, float* att, int C, int heads, int T, int NW, const sast_selection& sel, cudaStream_t st);
int launch_rows_gather(const float* map, float* rows, const sast_selection* sel, int C, cudaStream_t st);

namespace bw {

// ------------------------------------------------------------------------------------------------------------------
// generic fp32 GEMM  C[i,j] (+)= sum_k A(i,k) B(k,j)   with arbitrary element strides, 64x64x16 tiles, 4x4 per thread.
// dyn = 1: the row count Mi is read from counts[1] (compacted rows); dyn = 2: the reduction length Kk is (weight
// gradients: reduction over the compacted rows), split over gridDim.z with atomicAdd.
// ------------------------------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A; long long sai, sak;
  const float* B; long long sbk, sbj;
  float* C; long long ldc;
  const float* bias;       // [Nj] added when accumulate == 0 and blockIdx.z == 0
  int Mi, Nj, Kk;
  int dyn;                 // 0 static, 1 Mi = counts[1], 2 Kk = counts[1]
  int accumulate;          // 1: C += (non-atomic unless split), 0: C =
  const int* counts;
};

__global__ void __launch_bounds__(256) gemm_gen_kernel(GemmArgs a) {
  pdl_entry();
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  int Mi = a.Mi, Kk = a.Kk;
  if (a.dyn == 1) Mi = a.counts[1];
  if (a.dyn == 2) Kk = a.counts[1];
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  if (i0 >= Mi) return;
  // reduction range of this split
  const int nz = gridDim.z;
  const int chunk = ((Kk + nz - 1) / nz + 15) / 16 * 16;
  const int k_begin = blockIdx.z * chunk, k_end = min(Kk, k_begin + chunk);
  if (k_begin >= k_end) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_icontig = a.sai == 1, b_jcontig = a.sbj == 1;
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + 256 * q;
      {
        const int i = a_icontig ? (e & 63) : (e >> 4), k = a_icontig ? (e >> 6) : (e & 15);
        const int gi = i0 + i, gk = k0 + k;
        As[k][i] = (gi < Mi && gk < k_end) ? a.A[gi * a.sai + gk * a.sak] : 0.f;
      }
      {
        const int j = b_jcontig ? (e & 63) : (e >> 4), k = b_jcontig ? (e >> 6) : (e & 15);
        const int gj = j0 + j, gk = k0 + k;
        Bs[k][j] = (gj < a.Nj && gk < k_end) ? a.B[gk * a.sbk + gj * a.sbj] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float am[4], bn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) am[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bn[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bn[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = i0 + ty * 4 + i;
    if (gi >= Mi) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = j0 + tx * 4 + j;
      if (gj >= a.Nj) continue;
      float v = acc[i][j];
      float* dst = a.C + gi * a.ldc + gj;
      if (nz > 1) atomicAdd(dst, v);
      else if (a.accumulate) *dst += v;
      else *dst = v + (a.bias ? a.bias[gj] : 0.f);
    }
  }
}

static int launch_gemm(const GemmArgs& a, long long max_rows, int splits, cudaStream_t st) {
  const long long mi = a.dyn == 1 ? max_rows : a.Mi;
  const dim3 grid((unsigned)((mi + 63) / 64), (unsigned)((a.Nj + 63) / 64), (unsigned)splits);
  sast::launch_k(gemm_gen_kernel, grid, 256, 0, st, a);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}
// D[rows, N] = X[rows, K] W[N, K]^T + bias  (forward, rows from counts)
static int gemm_nt(const float* X, const float* W, const float* bias, float* D, int N, int K, const int* counts, long long max_rows, cudaStream_t st) {
  GemmArgs a{X, K, 1, W, 1, K, D, N, bias, 0, N, K, 1, 0, counts};
  return launch_gemm(a, max_rows, 1, st);
}
// D[rows, K] (+)= G[rows, N] W[N, K]   (input gradient)
static int gemm_nn(const float* G, const float* W, float* D, int N, int K, int accumulate, const int* counts, long long max_rows, cudaStream_t st) {
  GemmArgs a{G, N, 1, W, K, 1, D, K, nullptr, 0, K, N, 1, accumulate, counts};
  return launch_gemm(a, max_rows, 1, st);
}
// dW[N, K] += G[rows, N]^T X[rows, K]   (weight gradient: reduction over the rows, split + atomics; dW zero on entry)
static int gemm_tn(const float* G, const float* X, float* dW, int N, int K, const int* counts, long long max_rows, cudaStream_t st) {
  GemmArgs a{G, 1, N, X, K, 1, dW, K, nullptr, N, K, 0, 2, 1, counts};
  int splits = (int)((max_rows + 1023) / 1024);
  splits = splits < 1 ? 1 : splits > 64 ? 64 : splits;
  if (splits == 1) splits = 2;                  // always the atomic path (dW may already hold another layer call's sum)
  return launch_gemm(a, max_rows, splits, st);
}

// ------------------------------------------------------------------------------------------------------------------
// column reductions over the compacted rows:  dgamma[c] += sum_r G V,  out = G * gamma (or G),  dbias[c] += sum_r out
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colreduce_kernel(const float* __restrict__ G, const float* __restrict__ V,
                                                        const float* __restrict__ gamma, float* __restrict__ out,
                                                        float* __restrict__ dgamma, float* __restrict__ dbias, int N,
                                                        const int* __restrict__ counts, int rows_static) {
  pdl_entry();
  __shared__ float red[2][8][32];
  const int rows = counts ? counts[1] : rows_static;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), wy = threadIdx.x >> 5;
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float sg = 0.f, sb = 0.f;
  if (c < N) {
    const float gm = gamma ? gamma[c] : 1.f;
    for (int r = r0 + wy; r < r1; r += 8) {
      const float g = G[(size_t)r * N + c];
      const float o = g * gm;
      if (V) sg += g * V[(size_t)r * N + c];
      sb += o;
      if (out) out[(size_t)r * N + c] = o;
    }
  }
  red[0][wy][threadIdx.x & 31] = sg; red[1][wy][threadIdx.x & 31] = sb;
  __syncthreads();
  if (wy == 0 && c < N) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 8; ++i) { a += red[0][i][threadIdx.x]; b += red[1][i][threadIdx.x]; }
    if (dgamma && V) atomicAdd(dgamma + c, a);
    if (dbias) atomicAdd(dbias + c, b);
  }
}
static int colreduce(const float* G, const float* V, const float* gamma, float* out, float* dgamma, float* dbias, int N,
                     const int* counts, long long max_rows, cudaStream_t st) {
  int ry = (int)((max_rows + 511) / 512);
  ry = ry < 1 ? 1 : ry > 128 ? 128 : ry;
  sast::launch_k(colreduce_kernel, dim3((N + 31) / 32, ry), 256, 0, st, G, V, gamma, out, dgamma, dbias, N, counts, (int)max_rows);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

// y = a + gamma * b   (rows from counts)
__global__ void resid_scale_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gamma, float* __restrict__ y,
                                   int N, const int* __restrict__ counts) {
  pdl_entry();
  const long long total = (long long)counts[1] * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    y[i] = a[i] + (gamma ? gamma[i % N] : 1.f) * b[i];
}

// GLU (ops.py:135-137), u interleaved value_j, gate_j:   forward  hid = val * gelu(gate)
__global__ void glu_fwd_kernel(const float* __restrict__ u, float* __restrict__ hid, int I, const int* __restrict__ counts) {
  pdl_entry();
  const long long total = (long long)counts[1] * I;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float2 p = reinterpret_cast<const float2*>(u)[i];
    hid[i] = p.x * gelu_erf(p.y);
  }
}
// backward, in place: u <- (dhid * gelu(gate), dhid * val * gelu'(gate))
__global__ void glu_bwd_kernel(float* __restrict__ u, const float* __restrict__ dhid, int I, const int* __restrict__ counts) {
  pdl_entry();
  const long long total = (long long)counts[1] * I;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float2 p = reinterpret_cast<float2*>(u)[i];
    const float d = dhid[i], x = p.y;
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    reinterpret_cast<float2*>(u)[i] = make_float2(d * x * cdf, d * p.x * (cdf + x * pdf));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// attention backward over the compacted rows of one window and one head (fp32):  qkv, d(att) -> d(qkv)
// phase 1 (thread = query i): m_i, l_i, D_i = do_i . o_i, dq_i;   phase 2 (thread = key j): dk_j, dv_j.  No atomics.
// ------------------------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(128) attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ datt,
                                                            float* __restrict__ dqkv, int C, const int* __restrict__ win_K,
                                                            const int* __restrict__ win_row0) {
  pdl_entry();
  extern __shared__ __align__(16) float sm[];       // q, k, v, do: [K][DH] each; then m, l, D: [K] each
  const int w = blockIdx.x, h = blockIdx.y;
  const int K = win_K[w];
  if (K == 0) return;
  const int row0 = win_row0[w];
  const int ld = 3 * C;
  float *qs = sm, *ks = sm + K * DH, *vs = sm + 2 * K * DH, *ds = sm + 3 * K * DH;
  float *ms = sm + 4 * K * DH, *ls = ms + K, *Ds = ls + K;
  for (int i = threadIdx.x; i < K * DH; i += blockDim.x) {
    const int r = i / DH, c = i % DH;
    const float* src = qkv + (size_t)(row0 + r) * ld + h * 3 * DH + c;
    qs[i] = src[0]; ks[i] = src[DH]; vs[i] = src[2 * DH];
    ds[i] = datt[(size_t)(row0 + r) * C + h * DH + c];
  }
  __syncthreads();
  const float scale = rsqrtf((float)DH);
  const int i = threadIdx.x;
  if (i < K) {
    float q[DH], dq[DH], o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { q[d] = qs[i * DH + d]; dq[d] = 0.f; o[d] = 0.f; }
    float mx = -INFINITY;
    for (int j = 0; j < K; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s = fmaf(q[d], ks[j * DH + d], s);
      mx = fmaxf(mx, s * scale);
    }
    float l = 0.f;
    for (int j = 0; j < K; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s = fmaf(q[d], ks[j * DH + d], s);
      const float p = expf(s * scale - mx);
      l += p;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(p, vs[j * DH + d], o[d]);
    }
    const float il = 1.0f / l;
    float D = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) D = fmaf(ds[i * DH + d], o[d] * il, D);
    for (int j = 0; j < K; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { s = fmaf(q[d], ks[j * DH + d], s); dp = fmaf(ds[i * DH + d], vs[j * DH + d], dp); }
      const float p = expf(s * scale - mx) * il;
      const float dsij = p * (dp - D) * scale;
#pragma unroll
      for (int d = 0; d < DH; ++d) dq[d] = fmaf(dsij, ks[j * DH + d], dq[d]);
    }
    ms[i] = mx; ls[i] = il; Ds[i] = D;
    float* dst = dqkv + (size_t)(row0 + i) * ld + h * 3 * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) dst[d] = dq[d];
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j < K) {
    float kk[DH], vv[DH], dk[DH], dv[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { kk[d] = ks[j * DH + d]; vv[d] = vs[j * DH + d]; dk[d] = 0.f; dv[d] = 0.f; }
    for (int r = 0; r < K; ++r) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { s = fmaf(qs[r * DH + d], kk[d], s); dp = fmaf(ds[r * DH + d], vv[d], dp); }
      const float p = expf(s * scale - ms[r]) * ls[r];
      const float dsij = p * (dp - Ds[r]) * scale;
#pragma unroll
      for (int d = 0; d < DH; ++d) { dk[d] = fmaf(dsij, qs[r * DH + d], dk[d]); dv[d] = fmaf(p, ds[r * DH + d], dv[d]); }
    }
    float* dst = dqkv + (size_t)(row0 + j) * ld + h * 3 * DH;
#pragma unroll
    for (int d = 0; d < DH; ++d) { dst[DH + d] = dk[d]; dst[2 * DH + d] = dv[d]; }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LN2 + LN1 backward over ALL tokens (partitioned order): selected tokens take dn2 from the compacted rows,
// unselected tokens pass d(out) straight to LN1 (they are norm1(x), SAST.py:251-254).  One warp per token, C <= 1024.
// ------------------------------------------------------------------------------------------------------------------
template <int NV>        // float per lane = C / 32 rounded up
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, const float* __restrict__ dn2,
                                                     const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                                     float eps, const int* __restrict__ tok_row, Geom g, int flavor, float* __restrict__ dx,
                                                     float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int C = g.C;
  const float inv_c = 1.0f / (float)C;
  float aw1[NV], ab1[NV], aw2[NV], ab2[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { aw1[i] = ab1[i] = aw2[i] = ab2[i] = 0.f; }
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < g.P; q += nwarp) {
    const long long pix = token_pixel(q, g, flavor);
    const int row = tok_row[q];
    float z[NV], zh1[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) { const int c = lane + 32 * i; z[i] = c < C ? x[pix * C + c] : 0.f; s += z[i]; }
    const float mu1 = warp_sum(s) * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) { const int c = lane + 32 * i; const float d = c < C ? z[i] - mu1 : 0.f; ss += d * d; }
    const float r1 = rsqrtf(warp_sum(ss) * inv_c + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) { const int c = lane + 32 * i; zh1[i] = c < C ? (z[i] - mu1) * r1 : 0.f; }
    float dn1[NV];
    if (row >= 0) {
      float n1[NV], zh2[NV];
      s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) { const int c = lane + 32 * i; n1[i] = c < C ? zh1[i] * w1[c] + b1[c] : 0.f; s += n1[i]; }
      const float mu2 = warp_sum(s) * inv_c;
      ss = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) { const int c = lane + 32 * i; const float d = c < C ? n1[i] - mu2 : 0.f; ss += d * d; }
      const float r2 = rsqrtf(warp_sum(ss) * inv_c + eps);
      float s1 = 0.f, s2 = 0.f, dz[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        zh2[i] = c < C ? (n1[i] - mu2) * r2 : 0.f;
        const float gq = c < C ? dn2[(size_t)row * C + c] : 0.f;
        aw2[i] += gq * zh2[i]; ab2[i] += gq;
        dz[i] = c < C ? gq * w2[c] : 0.f;
        s1 += dz[i]; s2 += dz[i] * zh2[i];
      }
      s1 = warp_sum(s1) * inv_c; s2 = warp_sum(s2) * inv_c;
#pragma unroll
      for (int i = 0; i < NV; ++i) dn1[i] = r2 * (dz[i] - s1 - zh2[i] * s2);
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) { const int c = lane + 32 * i; dn1[i] = c < C ? dout[pix * C + c] : 0.f; }
    }
    float s1 = 0.f, s2 = 0.f, dz[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      aw1[i] += dn1[i] * zh1[i]; ab1[i] += dn1[i];
      dz[i] = c < C ? dn1[i] * w1[c] : 0.f;
      s1 += dz[i]; s2 += dz[i] * zh1[i];
    }
    s1 = warp_sum(s1) * inv_c; s2 = warp_sum(s2) * inv_c;
#pragma unroll
    for (int i = 0; i < NV; ++i) { const int c = lane + 32 * i; if (c < C) dx[pix * C + c] = r1 * (dz[i] - s1 - zh1[i] * s2); }
  }
  // block-level sums in shared memory first (8 warps -> one set of global atomics per CTA)
  __shared__ float blk[4][NV * 32];
  for (int i = threadIdx.x; i < 4 * NV * 32; i += blockDim.x) (&blk[0][0])[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    atomicAdd(&blk[0][c], aw1[i]); atomicAdd(&blk[1][c], ab1[i]); atomicAdd(&blk[2][c], aw2[i]); atomicAdd(&blk[3][c], ab2[i]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(dw1 + c, blk[0][c]); atomicAdd(db1 + c, blk[1][c]); atomicAdd(dw2 + c, blk[2][c]); atomicAdd(db2 + c, blk[3][c]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// scoring module backward, element-wise part:   xw = sig(ctrl_b) sig(s) x0,  s = relu(s_pre)
//   dx0 = dxw a g;  ds = dxw a x0 g (1 - g) [s_pre > 0]  (written over s_pre);  dctrl[b,c] += sum_tokens dxw g x0 a (1 - a)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_bwd_elem_kernel(const float* __restrict__ dxw, const float* __restrict__ x0, float* __restrict__ s_pre,
                                                             const float* __restrict__ sig, int HW, int C, float* __restrict__ dx0,
                                                             float* __restrict__ dctrl) {
  pdl_entry();
  __shared__ float red[8][32];
  const int b = blockIdx.z;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), wy = threadIdx.x >> 5;
  const int per = (HW + gridDim.y - 1) / gridDim.y;
  const int t0 = blockIdx.y * per, t1 = min(HW, t0 + per);
  float acc = 0.f;
  if (c < C) {
    const float a = sig[(size_t)b * C + c];
    for (int t = t0 + wy; t < t1; t += 8) {
      const size_t i = ((size_t)b * HW + t) * C + c;
      const float sp = s_pre[i], d = dxw[i], xv = x0[i];
      const float s = fmaxf(sp, 0.f);
      const float gq = sigmoidf_acc(s);
      dx0[i] = d * a * gq;
      s_pre[i] = sp > 0.f ? d * a * xv * gq * (1.f - gq) : 0.f;
      acc += d * gq * xv;
    }
    acc *= a * (1.f - a);
  }
  red[wy][threadIdx.x & 31] = acc;
  __syncthreads();
  if (wy == 0 && c < C) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(dctrl + (size_t)b * C + c, t);
  }
}
// dWc[c,k] = sum_b dctrl[b,c] exp(Wc[c,k]) (r[b,k] + 1e-6)      (ctrl = sum_k exp(Wc) (r + 1e-6), SAST.py:325-328)
__global__ void controls_bwd_kernel(const float* __restrict__ dctrl, const float* __restrict__ ctrl_w, const float* __restrict__ r, int B,
                                    int C, int n_bins, float* __restrict__ dctrl_w) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * n_bins) return;
  const int c = i / n_bins, k = i - c * n_bins;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += dctrl[(size_t)b * C + c] * (r[(size_t)b * n_bins + k] + 1e-6f);
  dctrl_w[i] = s * expf(ctrl_w[i]);
}
__global__ void controls_sig_kernel(const float* __restrict__ r, const float* __restrict__ ctrl_w, int n_bins, int C, float* __restrict__ sig) {
  pdl_entry();
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < n_bins; ++j) acc += expf(ctrl_w[c * n_bins + j]) * (r[b * n_bins + j] + 1e-6f);
    sig[b * C + c] = sigmoidf_acc(acc);
  }
}
__global__ void add_pos_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ pos, long long pos_bstride4, long long HWC4,
                                   long long total4, float4* __restrict__ out) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = x[i];
    const float4 p = pos[(i / HWC4) * pos_bstride4 + (i % HWC4)];
    out[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}

}  // namespace bw
}  // namespace sast

// floats of workspace sast_layer_bwd needs: 7 [P,C] buffers + [P,3C] + [P,2I] + [P,I]
extern "C" size_t sast_layer_bwd_workspace_bytes(int64_t P, int32_t C, int32_t I) {
  return ((size_t)P * (10 * (size_t)C + 3 * (size_t)I) + 1024) * sizeof(float);
}

extern "C" int sast_layer_bwd(const sast_layer_args* ap, const float* d_out, float* dx, const sast_layer_grads* gp, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(ap); SAST_CHECK_PTR(d_out); SAST_CHECK_PTR(dx); SAST_CHECK_PTR(gp);
  const sast_layer_args& a = *ap;
  const sast_layer_grads& gr = *gp;
  SAST_CHECK_PTR(a.x); SAST_CHECK_PTR(a.workspace);
  if (a.enable_cb) return SAST_E_UNSUPPORTED;
  int rc = check_geom(a.g, a.flavor);
  if (rc) return rc;
  const sast_layer_weights& w = a.w;
  SAST_CHECK_PTR(w.ln1_w); SAST_CHECK_PTR(w.ln1_b); SAST_CHECK_PTR(w.ln2_w); SAST_CHECK_PTR(w.ln2_b);
  SAST_CHECK_PTR(w.qkv_w); SAST_CHECK_PTR(w.proj_w); SAST_CHECK_PTR(w.mlp1_w); SAST_CHECK_PTR(w.mlp2_w);
  SAST_CHECK_PTR(gr.ln1_w); SAST_CHECK_PTR(gr.ln1_b); SAST_CHECK_PTR(gr.ln2_w); SAST_CHECK_PTR(gr.ln2_b);
  SAST_CHECK_PTR(gr.qkv_w); SAST_CHECK_PTR(gr.proj_w); SAST_CHECK_PTR(gr.mlp1_w); SAST_CHECK_PTR(gr.mlp2_w);
  const Geom g = make_geom(a.g, a.flavor);
  const int C = g.C, I = w.I;
  const int dh = w.dim_head > 0 ? w.dim_head : 32;
  if (C % 8 != 0 || I % 8 != 0 || I <= 0 || C > 1024 || g.T > 128 || C % dh != 0) return SAST_E_SHAPE;
  if (dh != 8 && dh != 16 && dh != 24 && dh != 32) return SAST_E_UNSUPPORTED;
  if (2 * I < 3 * C) return SAST_E_UNSUPPORTED;          // d(qkv) re-uses the GLU pre-activation buffer
  if (a.workspace_bytes < sast_layer_bwd_workspace_bytes(g.P, C, I)) return SAST_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long P = g.P;
  const int* counts = a.sel.counts;
  float* p = (float*)a.workspace;
  auto take = [&](size_t n) { float* r = p; p += (n + 63) / 64 * 64; return r; };
  float* n2 = take((size_t)P * C);
  float* qkv = take((size_t)P * 3 * C);
  float* att = take((size_t)P * C);
  float* po = take((size_t)P * C);
  float* y = take((size_t)P * C);
  float* u = take((size_t)P * 2 * I);
  float* hid = take((size_t)P * I);
  float* m = take((size_t)P * C);
  float* gy = take((size_t)P * C);
  float* t1 = take((size_t)P * C);
  const int heads = C / dh;
  const unsigned eb = 148 * 8;

  // ---- recompute the forward on the compacted rows (fp32) ----
  if ((rc = launch_gather_ln_f32(a.x, nullptr, w, a.sel, g, a.flavor, n2, st))) return rc;
  if ((rc = bw::gemm_nt(n2, w.qkv_w, w.qkv_b, qkv, 3 * C, C, counts, P, st))) return rc;
  if ((rc = launch_attention_f32(qkv, att, C, heads, g.T, g.NW, a.sel, st))) return rc;
  if ((rc = bw::gemm_nt(att, w.proj_w, w.proj_b, po, C, C, counts, P, st))) return rc;
  sast::launch_k(bw::resid_scale_kernel, eb, 256, 0, st, (const float*)n2, (const float*)po, w.gamma1, y, C, counts);
  SAST_LAUNCH_CHECK();
  if ((rc = bw::gemm_nt(y, w.mlp1_w, w.mlp1_b, u, 2 * I, C, counts, P, st))) return rc;
  sast::launch_k(bw::glu_fwd_kernel, eb, 256, 0, st, (const float*)u, hid, I, counts);
  SAST_LAUNCH_CHECK();
  if ((rc = bw::gemm_nt(hid, w.mlp2_w, w.mlp2_b, m, C, I, counts, P, st))) return rc;

  // ---- backward ----
  if ((rc = launch_rows_gather(d_out, gy, &a.sel, C, st))) return rc;                              // d(out) of the selected rows = dy (residual)
  if ((rc = bw::colreduce(gy, m, w.gamma2, t1, gr.gamma2, gr.mlp2_b, C, counts, P, st))) return rc;  // dm = g2 d_out; dgamma2; db2
  if ((rc = bw::gemm_tn(t1, hid, gr.mlp2_w, C, I, counts, P, st))) return rc;                      // dW2 = dm^T hid
  if ((rc = bw::gemm_nn(t1, w.mlp2_w, hid, C, I, 0, counts, P, st))) return rc;                    // dhid = dm W2      (over hid)
  sast::launch_k(bw::glu_bwd_kernel, eb, 256, 0, st, u, (const float*)hid, I, counts);           // du                (over u)
  SAST_LAUNCH_CHECK();
  if ((rc = bw::colreduce(u, nullptr, nullptr, nullptr, nullptr, gr.mlp1_b, 2 * I, counts, P, st))) return rc;
  if ((rc = bw::gemm_tn(u, y, gr.mlp1_w, 2 * I, C, counts, P, st))) return rc;                     // dW1 = du^T y
  if ((rc = bw::gemm_nn(u, w.mlp1_w, gy, 2 * I, C, 1, counts, P, st))) return rc;                  // dy += du W1
  if ((rc = bw::colreduce(gy, po, w.gamma1, t1, gr.gamma1, gr.proj_b, C, counts, P, st))) return rc; // dpo = g1 dy; dgamma1; dbp
  if ((rc = bw::gemm_tn(t1, att, gr.proj_w, C, C, counts, P, st))) return rc;                      // dWp = dpo^T att
  if ((rc = bw::gemm_nn(t1, w.proj_w, m, C, C, 0, counts, P, st))) return rc;                      // datt = dpo Wp     (over m)
  {
    const size_t smem = ((size_t)4 * g.T * dh + 3 * g.T) * sizeof(float);
    static thread_local unsigned long long attr_mask = 0;
    const bool first = first_use_on_device(attr_mask);
#define SAST_ATT_BWD(DH)                                                                                                      \
    {                                                                                                                         \
      if (first) {                                                                                                            \
        cudaError_t e = cudaFuncSetAttribute(bw::attention_bwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (4 * 128 * DH + 3 * 128) * 4); \
        if (e != cudaSuccess) return (int)e;                                                                                  \
      }                                                                                                                       \
      sast::launch_k(bw::attention_bwd_kernel<DH>, dim3(g.NW, heads), 128, smem, st, (const float*)qkv, (const float*)m, u, C, a.sel.win_K, \
                     a.sel.win_row0);                                                              /* dqkv (over u) */       \
    }
    if (dh == 32) SAST_ATT_BWD(32) else if (dh == 24) SAST_ATT_BWD(24) else if (dh == 16) SAST_ATT_BWD(16) else SAST_ATT_BWD(8)
#undef SAST_ATT_BWD
    SAST_LAUNCH_CHECK();
  }
  if ((rc = bw::colreduce(u, nullptr, nullptr, nullptr, nullptr, gr.qkv_b, 3 * C, counts, P, st))) return rc;
  if ((rc = bw::gemm_tn(u, n2, gr.qkv_w, 3 * C, C, counts, P, st))) return rc;                     // dWqkv = dqkv^T n2
  if ((rc = bw::gemm_nn(u, w.qkv_w, gy, 3 * C, C, 1, counts, P, st))) return rc;                   // dn2 = dy + dqkv Wqkv
  {
    const unsigned grid = 148 * 4;
    const int nv = (C + 31) / 32;
#define SAST_LN_BWD(NV) sast::launch_k(bw::ln_bwd_kernel<NV>, grid, 256, 0, st, a.x, d_out, (const float*)gy, w.ln1_w, w.ln1_b, w.ln2_w, w.ln_eps, \
                                        a.sel.tok_row, g, a.flavor, dx, gr.ln1_w, gr.ln1_b, gr.ln2_w, gr.ln2_b)
    if (nv <= 1) SAST_LN_BWD(1); else if (nv <= 2) SAST_LN_BWD(2); else if (nv <= 4) SAST_LN_BWD(4); else if (nv <= 8) SAST_LN_BWD(8);
    else if (nv <= 16) SAST_LN_BWD(16); else SAST_LN_BWD(32);
#undef SAST_LN_BWD
    SAST_LAUNCH_CHECK();
  }
  return SAST_OK;
}

// Scoring-module backward.  workspace: 3 [P,C] float buffers + [B,C] x 2.
extern "C" size_t sast_score_bwd_workspace_bytes(int64_t P, int32_t C, int32_t B) {
  return ((size_t)P * C * 3 + (size_t)B * C * 2 + 256) * sizeof(float);
}

extern "C" int sast_score_bwd(const sast_score_args* ap, const float* d_xw, float* dx, float* d_score_w, float* d_score_b,
                              float* d_ctrl_w, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(ap); SAST_CHECK_PTR(d_xw); SAST_CHECK_PTR(dx);
  const sast_score_args& a = *ap;
  SAST_CHECK_PTR(a.x); SAST_CHECK_PTR(a.pos);
  const sast_geom& g = a.g;
  if (g.B <= 0 || g.H <= 0 || g.W <= 0 || g.C <= 0 || g.C % 4 != 0) return SAST_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long P = (long long)g.B * g.H * g.W;
  const int HW = g.H * g.W, C = g.C;
  if (a.score_w == nullptr) {        // non-first block: xw = x + pos  ->  dx = d_xw
    return (int)cudaMemcpyAsync(dx, d_xw, (size_t)P * C * sizeof(float), cudaMemcpyDeviceToDevice, st);
  }
  SAST_CHECK_PTR(a.r); SAST_CHECK_PTR(a.ctrl_w); SAST_CHECK_PTR(a.score_b); SAST_CHECK_PTR(workspace);
  SAST_CHECK_PTR(d_score_w); SAST_CHECK_PTR(d_score_b); SAST_CHECK_PTR(d_ctrl_w);
  if (workspace_bytes < sast_score_bwd_workspace_bytes(P, C, g.B)) return SAST_E_WORKSPACE;
  float* x0 = (float*)workspace;
  float* sp = x0 + (size_t)P * C;
  float* sig = sp + (size_t)P * C;
  float* dctrl = sig + (size_t)g.B * C;
  const long long total4 = P * C / 4;
  sast::launch_k(bw::add_pos_bwd_kernel, 148 * 8, 256, 0, st, (const float4*)a.x, (const float4*)a.pos, a.pos_batch_stride / 4,
                 (long long)HW * C / 4, total4, (float4*)x0);
  SAST_LAUNCH_CHECK();
  sast::launch_k(bw::controls_sig_kernel, g.B, 128, 0, st, a.r, a.ctrl_w, a.n_bins, C, sig);
  SAST_LAUNCH_CHECK();
  cudaError_t e = cudaMemsetAsync(dctrl, 0, (size_t)g.B * C * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  {   // s_pre = x0 Ws^T + bs
    bw::GemmArgs ga{x0, C, 1, a.score_w, 1, C, sp, C, a.score_b, (int)P, C, C, 0, 0, nullptr};
    int rc = bw::launch_gemm(ga, P, 1, st);
    if (rc) return rc;
  }
  int ry = (HW + 511) / 512;
  ry = ry < 1 ? 1 : ry > 64 ? 64 : ry;
  sast::launch_k(bw::score_bwd_elem_kernel, dim3((C + 31) / 32, ry, g.B), 256, 0, st, d_xw, (const float*)x0, sp, (const float*)sig, HW, C, dx, dctrl);
  SAST_LAUNCH_CHECK();
  int rc;
  if ((rc = bw::colreduce(sp, nullptr, nullptr, nullptr, nullptr, d_score_b, C, nullptr, P, st))) return rc;       // dbs = sum ds
  {   // dWs[N=C, K=C] += ds^T x0   (reduction over all P tokens)
    bw::GemmArgs ga{sp, 1, C, x0, C, 1, d_score_w, C, nullptr, C, C, (int)P, 0, 1, nullptr};
    int splits = (int)((P + 1023) / 1024);
    splits = splits < 2 ? 2 : splits > 64 ? 64 : splits;
    if ((rc = bw::launch_gemm(ga, P, splits, st))) return rc;
  }
  {   // dx0 += ds Ws
    bw::GemmArgs ga{sp, C, 1, a.score_w, C, 1, dx, C, nullptr, (int)P, C, C, 0, 1, nullptr};
    if ((rc = bw::launch_gemm(ga, P, 1, st))) return rc;
  }
  sast::launch_k(bw::controls_bwd_kernel, (C * a.n_bins + 127) / 128, 128, 0, st, (const float*)dctrl, a.ctrl_w, a.r, g.B, C, a.n_bins, d_ctrl_w);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}
