// The dense callers on either side of the SAST block (SURVEY.md section 8f "next" rows 1-2):
// memory-bound glue that PyTorch runs as 10-20 separate kernels per stage.
//   sast_pad_input   : stem input: [B,Cin,H,W] u8/i32/f32 NCHW -> fp32 NHWC with replicate
//                      padding, in one pass (replaces x.float() + F.pad(mode='replicate') + the
//                      NCHW->NHWC conversion in front of the strided conv, ops.py:77-89).
//   sast_pad_nhwc    : the same for an fp32 NHWC map (downsample convs of stages 2-4).
//   sast_layernorm   : LayerNorm over C of an NHWC map (ops.py:85,90).
//   sast_lstm_gates  : the point-wise half of DWSConvLSTM2d (models/layers/rnn.py:58-69):
//                      sigmoid/tanh gates, cell update, hidden state, from the 1x1-conv output.
#include "common.cuh"

namespace sast {

// CTA = one padded output row segment of PX pixels: the Cin input planes are read with coalesced
// row loads into shared memory ([Cin][PX] in the input dtype), then written out pixel-major
// (Cin contiguous floats per pixel: dense 16-byte stores across the warp).
constexpr int PAD_PX = 256;
template <typename T>
__global__ void __launch_bounds__(256) pad_input_kernel(const T* __restrict__ x, int B, int Cin, int H, int W, int pad,
                                                        int cgroup, float* __restrict__ out) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t tile_raw[];
  T* tile = reinterpret_cast<T*>(tile_raw);            // [Cin][PAD_PX + 4]
  constexpr int LD = PAD_PX + 4;
  const int Ho = H + 2 * pad, Wo = W + 2 * pad;
  const int segs = (Wo + PAD_PX - 1) / PAD_PX;
  const int seg = blockIdx.x % segs;
  const int yo = (blockIdx.x / segs) % Ho;
  const int b = blockIdx.x / (segs * Ho);
  const int ys = min(max(yo - pad, 0), H - 1);
  const int x0 = seg * PAD_PX;
  const int npx = min(PAD_PX, Wo - x0);
  float* dst = out + ((size_t)(b * Ho + yo) * Wo + x0) * Cin;
  for (int cg0 = 0; cg0 < Cin; cg0 += cgroup) {            // channel groups sized to the shared-memory tile
    const int cg = min(cgroup, Cin - cg0);
    // PAD_PX == blockDim.x: thread t owns pixel column t of every plane; loads are issued four planes at a
    // time before any of them is consumed (one byte load in flight per thread is pure latency)
    const int dx = threadIdx.x;
    const int xs = min(max(x0 + dx - pad, 0), W - 1);
    const T* src = x + ((size_t)(b * Cin + cg0) * H + ys) * W + xs;
    const size_t plane = (size_t)H * W;
    int c = 0;
    for (; c + 4 <= cg; c += 4) {
      const T v0 = src[(size_t)c * plane], v1 = src[(size_t)(c + 1) * plane], v2 = src[(size_t)(c + 2) * plane],
              v3 = src[(size_t)(c + 3) * plane];
      if (dx < npx) { tile[c * LD + dx] = v0; tile[(c + 1) * LD + dx] = v1; tile[(c + 2) * LD + dx] = v2; tile[(c + 3) * LD + dx] = v3; }
    }
    for (; c < cg; ++c) {
      const T v0 = src[(size_t)c * plane];
      if (dx < npx) tile[c * LD + dx] = v0;
    }
    __syncthreads();
    if (cg % 4 == 0 && Cin % 4 == 0) {
      const int c4n = cg / 4;
      for (int i = threadIdx.x; i < npx * c4n; i += blockDim.x) {
        const int dx = i / c4n, c = (i - dx * c4n) * 4;
        *reinterpret_cast<float4*>(dst + (size_t)dx * Cin + cg0 + c) =
            make_float4((float)tile[c * LD + dx], (float)tile[(c + 1) * LD + dx], (float)tile[(c + 2) * LD + dx],
                        (float)tile[(c + 3) * LD + dx]);
      }
    } else {
      for (int i = threadIdx.x; i < npx * cg; i += blockDim.x) {
        const int dx = i / cg, c = i - dx * cg;
        dst[(size_t)dx * Cin + cg0 + c] = (float)tile[c * LD + dx];
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) pad_nhwc_kernel(const float4* __restrict__ x, int B, int H, int W, int C4, int pad,
                                                       long long sb, long long sy, long long sx,   // input strides in float4 units
                                                       float4* __restrict__ out) {
  pdl_entry();
  const int Ho = H + 2 * pad, Wo = W + 2 * pad;
  const long long total = (long long)B * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long p = i / C4;
    const int xo = (int)(p % Wo); p /= Wo;
    const int yo = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const int xs = min(max(xo - pad, 0), W - 1), ys = min(max(yo - pad, 0), H - 1);
    out[i] = x[b * sb + ys * sy + xs * sx + c];
  }
}

// the same, written as bf16: the operand map of the fused downsample kernel (downsample_tc.cu)
__global__ void __launch_bounds__(256) pad_nhwc_bf16_kernel(const float4* __restrict__ x, int B, int H, int W, int C4, int pad,
                                                            long long sb, long long sy, long long sx, uint2* __restrict__ out) {
  pdl_entry();
  const int Ho = H + 2 * pad, Wo = W + 2 * pad;
  const long long total = (long long)B * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long p = i / C4;
    const int xo = (int)(p % Wo); p /= Wo;
    const int yo = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const int xs = min(max(xo - pad, 0), W - 1), ys = min(max(yo - pad, 0), H - 1);
    const float4 v = x[b * sb + ys * sy + xs * sx + c];
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    out[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
  }
}

template <int LPT, int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                        const float* __restrict__ w, const float* __restrict__ b, float eps,
                                                        long long P, int C) {
  pdl_entry();
  constexpr int GROUPS = 32 / LPT, TPW = 4;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPT, l = lane % LPT;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long q0 = (warp * GROUPS + sub) * TPW;
  const float inv_c = 1.0f / (float)C;
  float4 v[TPW][NV];
#pragma unroll
  for (int j = 0; j < TPW; ++j) {
    const bool ok = q0 + j < P;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (l + LPT * i) * 4;
      v[j][i] = (ok && c < C) ? *reinterpret_cast<const float4*>(x + (q0 + j) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int j = 0; j < TPW; ++j) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    const float mean = s * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (l + LPT * i) * 4;
      if (c < C) {
        const float a = v[j][i].x - mean, bb = v[j][i].y - mean, cc = v[j][i].z - mean, d = v[j][i].w - mean;
        ss += (a * a + bb * bb) + (cc * cc + d * d);
      }
    }
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
    const float rstd = rsqrtf(ss * inv_c + eps);
    if (q0 + j >= P) continue;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (l + LPT * i) * 4;
      if (c < C) {
        float4 gw = make_float4(1.f, 1.f, 1.f, 1.f), gb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w) gw = *reinterpret_cast<const float4*>(w + c);
        if (b) gb = *reinterpret_cast<const float4*>(b + c);
        float4 o;
        o.x = (v[j][i].x - mean) * rstd * gw.x + gb.x;
        o.y = (v[j][i].y - mean) * rstd * gw.y + gb.y;
        o.z = (v[j][i].z - mean) * rstd * gw.z + gb.z;
        o.w = (v[j][i].w - mean) * rstd * gw.w + gb.w;
        *reinterpret_cast<float4*>(out + (q0 + j) * C + c) = o;
      }
    }
  }
}

// mix [P,4C] = conv1x1([x, h_prev]) (+bias): channels [0,3C) -> sigmoid -> (forget, input, output); [3C,4C) -> tanh -> g
__global__ void __launch_bounds__(256) lstm_gates_kernel(const float* __restrict__ mix, const float* __restrict__ bias,
                                                         const float* __restrict__ c_prev, long long P, int C,
                                                         float* __restrict__ h_out, float* __restrict__ c_out) {
  pdl_entry();
  const int c4n = C / 4;
  const long long total = P * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / c4n;
    const int c = (int)(i % c4n) * 4;
    const float* m = mix + p * 4 * C + c;
    float4 f4 = *reinterpret_cast<const float4*>(m);
    float4 i4 = *reinterpret_cast<const float4*>(m + C);
    float4 o4 = *reinterpret_cast<const float4*>(m + 2 * C);
    float4 g4 = *reinterpret_cast<const float4*>(m + 3 * C);
    if (bias) {   // the 1x1 conv's bias, folded in here instead of a separate pass over mix
      const float4 bf = *reinterpret_cast<const float4*>(bias + c), bi = *reinterpret_cast<const float4*>(bias + C + c);
      const float4 bo = *reinterpret_cast<const float4*>(bias + 2 * C + c), bg = *reinterpret_cast<const float4*>(bias + 3 * C + c);
      f4.x += bf.x; f4.y += bf.y; f4.z += bf.z; f4.w += bf.w;
      i4.x += bi.x; i4.y += bi.y; i4.z += bi.z; i4.w += bi.w;
      o4.x += bo.x; o4.y += bo.y; o4.z += bo.z; o4.w += bo.w;
      g4.x += bg.x; g4.y += bg.y; g4.z += bg.z; g4.w += bg.w;
    }
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c_prev) c0 = *reinterpret_cast<const float4*>(c_prev + p * C + c);
    const float f[4] = {f4.x, f4.y, f4.z, f4.w}, ii[4] = {i4.x, i4.y, i4.z, i4.w}, o[4] = {o4.x, o4.y, o4.z, o4.w},
                g[4] = {g4.x, g4.y, g4.z, g4.w}, cp[4] = {c0.x, c0.y, c0.z, c0.w};
    float cn[4], hn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cn[k] = sigmoidf_acc(f[k]) * cp[k] + sigmoidf_acc(ii[k]) * tanhf(g[k]);
      hn[k] = sigmoidf_acc(o[k]) * tanhf(cn[k]);
    }
    *reinterpret_cast<float4*>(c_out + p * C + c) = make_float4(cn[0], cn[1], cn[2], cn[3]);
    *reinterpret_cast<float4*>(h_out + p * C + c) = make_float4(hn[0], hn[1], hn[2], hn[3]);
  }
}

static inline unsigned grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = 148ll * 16;
  return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace sast

extern "C" int sast_pad_input(const void* x, int32_t dtype, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t pad,
                              float* out, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(x); SAST_CHECK_PTR(out);
  if (B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || pad < 0) return SAST_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((long long)B * (H + 2 * pad) * ((W + 2 * pad + PAD_PX - 1) / PAD_PX));
  const size_t esz = dtype == SAST_U8 ? 1 : 4;
  int cgroup = (int)((40 * 1024) / ((PAD_PX + 4) * esz));
  if (cgroup >= Cin) cgroup = Cin; else cgroup &= ~3;
  const size_t smem = (size_t)cgroup * (PAD_PX + 4) * esz;
  switch (dtype) {
    case SAST_U8: sast::launch_k(pad_input_kernel<uint8_t>, grid, 256, smem, st, (const uint8_t*)x, B, Cin, H, W, pad, cgroup, out); break;
    case SAST_I32: sast::launch_k(pad_input_kernel<int32_t>, grid, 256, smem, st, (const int32_t*)x, B, Cin, H, W, pad, cgroup, out); break;
    case SAST_F32: sast::launch_k(pad_input_kernel<float>, grid, 256, smem, st, (const float*)x, B, Cin, H, W, pad, cgroup, out); break;
    default: return SAST_E_UNSUPPORTED;
  }
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

extern "C" int sast_pad_nhwc(const float* x, int32_t B, int32_t H, int32_t W, int32_t C, int32_t pad, int64_t stride_b,
                             int64_t stride_y, int64_t stride_x, float* out, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(x); SAST_CHECK_PTR(out);
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 4 != 0 || pad < 0) return SAST_E_SHAPE;
  if (stride_b % 4 || stride_y % 4 || stride_x % 4) return SAST_E_SHAPE;
  const long long total = (long long)B * (H + 2 * pad) * (W + 2 * pad) * (C / 4);
  sast::launch_k(pad_nhwc_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const float4*)x, B, H, W, C / 4, pad, stride_b / 4,
                                                                         stride_y / 4, stride_x / 4, (float4*)out);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

extern "C" int sast_pad_nhwc_bf16(const float* x, int32_t B, int32_t H, int32_t W, int32_t C, int32_t pad, int64_t stride_b,
                                  int64_t stride_y, int64_t stride_x, uint16_t* out, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(x); SAST_CHECK_PTR(out);
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 4 != 0 || pad < 0) return SAST_E_SHAPE;
  if (stride_b % 4 || stride_y % 4 || stride_x % 4) return SAST_E_SHAPE;
  const long long total = (long long)B * (H + 2 * pad) * (W + 2 * pad) * (C / 4);
  sast::launch_k(pad_nhwc_bf16_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const float4*)x, B, H, W, C / 4, pad,
                 stride_b / 4, stride_y / 4, stride_x / 4, (uint2*)out);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

extern "C" int sast_layernorm(const float* x, const float* weight, const float* bias, float eps, int64_t P, int32_t C,
                              float* out, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(x); SAST_CHECK_PTR(out);
  if (P <= 0 || C <= 0 || C % 4 != 0 || C > 1024) return SAST_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
#define SAST_LN(LPT, NV)                                                                              \
  do {                                                                                                \
    const long long per_cta = 8ll * (32 / LPT) * 4;                                                   \
    sast::launch_k(layernorm_kernel<LPT, NV>, (unsigned)((P + per_cta - 1) / per_cta), 256, 0, st, x, out, weight, bias, eps, P, C); \
  } while (0)
  if (C <= 32) SAST_LN(8, 1);
  else if (C <= 64) SAST_LN(16, 1);
  else if (C <= 128) SAST_LN(32, 1);
  else if (C <= 256) SAST_LN(32, 2);
  else if (C <= 512) SAST_LN(32, 4);
  else SAST_LN(32, 8);
#undef SAST_LN
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}

extern "C" int sast_lstm_gates(const float* mix, const float* bias, const float* c_prev, int64_t P, int32_t C, float* h_out,
                               float* c_out, void* stream) {
  using namespace sast;
  SAST_CHECK_PTR(mix); SAST_CHECK_PTR(h_out); SAST_CHECK_PTR(c_out);
  if (P <= 0 || C <= 0 || C % 4 != 0) return SAST_E_SHAPE;
  sast::launch_k(lstm_gates_kernel, grid_for(P * (C / 4), 256), 256, 0, (cudaStream_t)stream, mix, bias, c_prev, P, C, h_out, c_out);
  SAST_LAUNCH_CHECK();
  return SAST_OK;
}
