// Shared device helpers for libsast_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/sast_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsast_b200 is written for sm_100a (NVIDIA B200) only"
#endif

#define SAST_CHECK_PTR(p) do { if ((p) == nullptr) return SAST_E_NULL; } while (0)
// every kernel launch goes through this: error check + statistics counter (see sast_launch_count)
extern "C" void sast_count_launch_(void);
#define SAST_LAUNCH_CHECK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; sast_count_launch_(); } while (0)

namespace sast {

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// Every kernel of this library is launched with programmaticStreamSerialization allowed and starts with
// pdl_entry(): it blocks until the preceding kernel in the stream has completed and its writes are
// visible (griddepcontrol.wait), then lets the NEXT kernel of the stream begin launching
// (griddepcontrol.launch_dependents).  The ~90 small launches of a forward thereby overlap their launch
// latency with the tail of their predecessor; data dependencies are exactly those of stream order.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

bool pdl_enabled();     // api.cu: SAST_B200_PDL=0 turns the launch attribute off (A/B knob)

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);     // errors surface through SAST_LAUNCH_CHECK
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE attribute: a launcher keeps one
// `static thread_local uint64_t` mask per kernel instantiation and sets the attribute the first time each device is
// used by each host thread (idempotent; no cross-thread state).
inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

struct Geom {
  int B, H, W, C, p0, p1, T, N, NW;
  long long P;
};

__host__ __device__ inline Geom make_geom(const sast_geom& g, int flavor) {
  Geom o;
  o.B = g.B; o.H = g.H; o.W = g.W; o.C = g.C; o.p0 = g.p0; o.p1 = g.p1;
  o.T = g.p0 * g.p1;
  o.N = (g.H * g.W) / o.T;
  o.NW = g.B * o.N;
  o.P = (long long)g.B * g.H * g.W;
  (void)flavor;
  return o;
}

inline int check_geom(const sast_geom& g, int flavor) {
  if (g.B <= 0 || g.H <= 0 || g.W <= 0 || g.C <= 0 || g.p0 <= 0 || g.p1 <= 0) return SAST_E_SHAPE;
  if (flavor != SAST_FLAT && (g.H % g.p0 != 0 || g.W % g.p1 != 0)) return SAST_E_SHAPE;
  if (flavor == SAST_FLAT && ((long long)g.H * g.W) % (g.p0 * g.p1) != 0) return SAST_E_SHAPE;
  if (g.p0 * g.p1 > 128) return SAST_E_UNSUPPORTED;   // a window's tokens must fit one 128-row tile
  if ((long long)g.B * g.H * g.W >= (1ll << 23)) return SAST_E_UNSUPPORTED;     // row_win packs a compacted row into 23 bits
  return SAST_OK;
}

// (window n of a frame, token t) -> pixel index inside the frame (y*W + x).
// WINDOW: ops.py:189-195, GRID: ops.py:206-212, FLAT: identity on n*T+t.
__device__ __forceinline__ int frame_pixel(int n, int t, int H, int W, int p0, int p1, int flavor) {
  if (flavor == SAST_WINDOW) {
    const int wj = W / p1;
    const int i = n / wj, j = n - i * wj;
    const int u = t / p1, v = t - u * p1;
    return (i * p0 + u) * W + j * p1 + v;
  } else if (flavor == SAST_GRID) {
    const int wb = W / p1, ha = H / p0;
    const int a = n / wb, b = n - a * wb;
    const int gi = t / p1, gj = t - gi * p1;
    return (gi * ha + a) * W + gj * wb + b;
  }
  return n * (p0 * p1) + t;
}

// token q = w*T + t in partitioned order -> global pixel index (b*H*W + y*W + x)
__device__ __forceinline__ long long token_pixel(long long q, const Geom& g, int flavor) {
  if (flavor == SAST_FLAT) return q;
  const int w = (int)(q / g.T), t = (int)(q - (long long)w * g.T);
  const int b = w / g.N, n = w - b * g.N;
  return (long long)b * g.H * g.W + frame_pixel(n, t, g.H, g.W, g.p0, g.p1, flavor);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// single-MUFU transcendentals of the tensor-core epilogues (relative error 2^-11, below the bf16 / TF32 noise there)
__device__ __forceinline__ float tanh_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
// sigmoid through the same single MUFU: 1/(1+e^-x) = 0.5 + 0.5 tanh(x/2)   (absolute error <= 2.5e-4)
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
// GELU of the tensor-core GLU epilogue, which is instruction-issue bound: erf(x/sqrt2) ~= tanh(x (a + b x^2)),
// a minimax fit over all x (|gelu error| <= 2.7e-4 from the fit + |x| 2.4e-4 from tanh.approx's 2^-11 relative
// error; the bf16 rounding of the result is 2e-3 relative).  Both coefficients are positive, so the argument is
// monotone and needs no clamp.  glu_tanh_fit returns value * gelu(x) in 6 instructions + ONE MUFU, with the 0.5 of
// the GELU folded into the caller's bias add of the value branch (pass half_value = 0.5 * value).
__device__ __forceinline__ float glu_tanh_fit(float half_value, float x) {
  const float p = fmaf(3.470089e-2f, x * x, 8.0015708e-1f);
  const float t = tanh_fast(x * p);
  return half_value * fmaf(x, t, x);                      // 0.5 v * x (1 + tanh)
}

extern long long* g_trace;   // api.cu; null unless sast_debug_trace armed it
extern int g_trace_which;    // 1: attention kernels stamp, 2: GEMM kernels, 3: scoring kernel

// Phase stamps for tools/attn_trace.py / tools/gemm_trace.py.  Compiled in only with -DSAST_TRACE (`make trace` ->
// libsast_b200_trace.so): even predicated-off stamps cost the GLU GEMM ~10 % through register allocation.
#ifdef SAST_TRACE
#define SAST_STAMP(ptr, cond, slot) do { if ((ptr) && (cond)) (ptr)[(slot)] = clock64(); } while (0)
#else
#define SAST_STAMP(ptr, cond, slot) ((void)0)
#endif

}  // namespace sast
