// a4: scoring module + STP weighting  (replaces SAST.py:105-119 and PositiveLinear :305-328).
//
// One pass over the NHWC map: x0 = x + pos is formed on the fly, the to_scores GEMM runs in
// fp32 FMA (selection is a threshold on a softmax of these numbers -- SURVEY.md "hard part 1"
// -- so the scores stay fp32-accurate), and the epilogue emits the STP-weighted map and ONE
// float per token, sum_c |amp/ctrl_c * s_c|, which is all that both layers' selection needs.
// The reference's [B,N,T,C] `scores` tensor and its two re-partitions are never materialised.
#include "common.cuh"

namespace sast {

// ctrl[b,c] = sum_j exp(Wc[c,j]) * (r[b,j] + 1e-6);  sig = sigmoid(ctrl);  inv = amp/ctrl (inf -> 0)
__global__ void controls_kernel(const float* __restrict__ r, const float* __restrict__ ctrl_w, int n_bins, int C,
                                float amp, float* __restrict__ sig, float* __restrict__ inv) {
  pdl_entry();
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < n_bins; ++j) acc += expf(ctrl_w[c * n_bins + j]) * (r[b * n_bins + j] + 1e-6f);
    sig[b * C + c] = sigmoidf_acc(acc);
    float iv = amp / acc;
    if (isinf(iv)) iv = 0.f;
    inv[b * C + c] = iv;
  }
}

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__global__ void __launch_bounds__(256) score_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                    long long pos_bstride, const float* __restrict__ Ws,
                                                    const float* __restrict__ bs, const float* __restrict__ sig,
                                                    const float* __restrict__ inv, int HW, int C, long long P,
                                                    float* __restrict__ xw, float* __restrict__ l1_out) {
  pdl_entry();
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  // loader roles
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const long long ltok = m0 + lrow;
  const bool ltok_ok = ltok < P;
  const long long lpos = ltok_ok ? ((ltok / HW) * pos_bstride + (ltok % HW) * (long long)C) : 0;

  float l1[4] = {0.f, 0.f, 0.f, 0.f};
  {
    const int n0 = blockIdx.y * BN;      // one 64-wide slice of output channels per CTA
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < C; k0 += BK) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ltok_ok) {
        const float4 xv = *reinterpret_cast<const float4*>(x + ltok * C + k0 + lk);
        const float4 pv = *reinterpret_cast<const float4*>(pos + lpos + k0 + lk);
        a = make_float4(xv.x + pv.x, xv.y + pv.y, xv.z + pv.z, xv.w + pv.w);
      }
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + lrow < C) w = *reinterpret_cast<const float4*>(Ws + (size_t)(n0 + lrow) * C + k0 + lk);
      As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
      Bs[lk + 0][lrow] = w.x; Bs[lk + 1][lrow] = w.y; Bs[lk + 2][lrow] = w.z; Bs[lk + 3][lrow] = w.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
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
    // epilogue for columns n0 + tx*4 .. +3 of rows m0 + ty*4 .. +3
    const int n = n0 + tx * 4;
    if (n < C) {
      const float4 bias = *reinterpret_cast<const float4*>(bs + n);
      const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long tok = m0 + ty * 4 + i;
        if (tok >= P) continue;
        const int b = (int)(tok / HW);
        const float4 xv = *reinterpret_cast<const float4*>(x + tok * C + n);
        const float4 pv = *reinterpret_cast<const float4*>(pos + (tok / HW) * pos_bstride + (tok % HW) * (long long)C + n);
        const float4 sg = *reinterpret_cast<const float4*>(sig + (size_t)b * C + n);
        const float4 iv = *reinterpret_cast<const float4*>(inv + (size_t)b * C + n);
        const float x0[4] = {xv.x + pv.x, xv.y + pv.y, xv.z + pv.z, xv.w + pv.w};
        const float sgv[4] = {sg.x, sg.y, sg.z, sg.w};
        const float ivv[4] = {iv.x, iv.y, iv.z, iv.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float s = fmaxf(acc[i][j] + bb[j], 0.f);
          o[j] = (sgv[j] * sigmoidf_acc(s)) * x0[j];
          l1[i] += fabsf(ivv[j] * s);
        }
        *reinterpret_cast<float4*>(xw + tok * C + n) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  // row L1: reduce over the 16 threads (tx) that share a row group
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = l1[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    const long long tok = m0 + ty * 4 + i;
    if (tx == 0 && tok < P) l1_out[(long long)blockIdx.y * P + tok] = v;   // partial over this channel slice
  }
}

// tok_score[p] = sum over channel slices, in slice order (deterministic)
__global__ void score_reduce_kernel(const float* __restrict__ part, int ny, long long P, float* __restrict__ tok_score) {
  pdl_entry();
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= P) return;
  float s = 0.f;
  for (int y = 0; y < ny; ++y) s += part[(long long)y * P + p];
  tok_score[p] = s;
}

__global__ void add_pos_kernel(const float4* __restrict__ x, const float4* __restrict__ pos, long long pos_bstride4,
                               long long HWC4, long long total4, float4* __restrict__ out) {
  pdl_entry();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = x[i];
    const float4 p = pos[(i / HWC4) * pos_bstride4 + (i % HWC4)];
    out[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}

int launch_score_tc(const sast_score_args* a, float* l1_part, int* n_slices, cudaStream_t st);

}  // namespace sast

extern "C" int sast_score_fwd(const sast_score_args* a, void* stream) {
  SAST_CHECK_PTR(a); SAST_CHECK_PTR(a->x); SAST_CHECK_PTR(a->pos); SAST_CHECK_PTR(a->xw);
  const sast_geom& g = a->g;
  if (g.B <= 0 || g.H <= 0 || g.W <= 0 || g.C <= 0 || g.C % 16 != 0) return SAST_E_SHAPE;
  if (a->score_w_hi && a->score_w_lo && g.C % 32 != 0) return SAST_E_SHAPE;        // tcgen05 scoring: C % 32
  if (a->pos_batch_stride % 4 != 0) return SAST_E_SHAPE;
  if (a->xw == a->x) return SAST_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const long long P = (long long)g.B * g.H * g.W;
  const int HW = g.H * g.W;
  if (a->score_w == nullptr) {   // non-first block: only x + pos (SAST.py:105, :124-128)
    const long long total4 = P * g.C / 4;
    const int blocks = (int)((total4 + 255) / 256 < 148 * 16 ? (total4 + 255) / 256 : 148 * 16);
    sast::launch_k(sast::add_pos_kernel, blocks, 256, 0, st, (const float4*)a->x, (const float4*)a->pos, a->pos_batch_stride / 4,
                                                 (long long)HW * g.C / 4, total4, (float4*)a->xw);
    SAST_LAUNCH_CHECK();
    return SAST_OK;
  }
  SAST_CHECK_PTR(a->r); SAST_CHECK_PTR(a->ctrl_w); SAST_CHECK_PTR(a->score_b); SAST_CHECK_PTR(a->tok_score);
  SAST_CHECK_PTR(a->ctrl_scratch);
  float* sig = a->ctrl_scratch;
  float* inv = a->ctrl_scratch + (size_t)g.B * g.C;
  float* part_buf = a->ctrl_scratch + 2 * (size_t)g.B * g.C;
  int ny;
  if (a->score_w_hi && a->score_w_lo) {      // 3xTF32 on tcgen05 (the kernel computes the control table itself)
    const int bn = g.C % 128 == 0 ? 128 : (g.C % 64 == 0 ? 64 : 32);
    float* part = g.C / bn == 1 ? a->tok_score : part_buf;
    int rc = sast::launch_score_tc(a, part, &ny, st);
    if (rc == SAST_OK) {
      if (ny > 1) {
        sast::launch_k(sast::score_reduce_kernel, (unsigned)((P + 255) / 256), 256, 0, st, part, ny, P, a->tok_score);
        SAST_LAUNCH_CHECK();
      }
      return SAST_OK;
    }
    if (rc != SAST_E_UNSUPPORTED) return rc;       // (control table of a very large batch does not fit: CUDA-core path below)
  }
  sast::launch_k(sast::controls_kernel, g.B, 128, 0, st, a->r, a->ctrl_w, a->n_bins, g.C, a->amp, sig, inv);
  SAST_LAUNCH_CHECK();
  ny = (g.C + sast::BN - 1) / sast::BN;
  const dim3 grid((unsigned)((P + sast::BM - 1) / sast::BM), ny);
  float* part = ny == 1 ? a->tok_score : part_buf;
  sast::launch_k(sast::score_kernel, grid, 256, 0, st, a->x, a->pos, a->pos_batch_stride, a->score_w, a->score_b, sig, inv, HW, g.C, P,
                                           a->xw, part);
  SAST_LAUNCH_CHECK();
  if (ny > 1) {
    sast::launch_k(sast::score_reduce_kernel, (unsigned)((P + 255) / 256), 256, 0, st, part, ny, P, a->tok_score);
    SAST_LAUNCH_CHECK();
  }
  return SAST_OK;
}
