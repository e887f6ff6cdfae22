// Micro-benchmark (debug aid): clocks per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, SS and TS form,
// dependent (one accumulator) and independent (alternating accumulators) chains.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../sast_b200/csrc/ptx.cuh"
using namespace sast;

template <int N, int TS, int CHAINS>
__global__ void __launch_bounds__(128, 1) mma_bench(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc(&tbase, 512);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 0) {
    const bool leader = ptx::elect_one();
    const uint32_t idesc = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(smem)), db = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(smem) + 16384);
    uint32_t phase = 0;
    for (int round = 0; round < 3; ++round) {
      const long long t0 = clock64();
      for (int r = 0; r < reps / 8; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t d = tbase + (uint32_t)((i % CHAINS) * N);
          if (leader) {
            if (TS) ptx::umma_f16_ts(d, tbase + 448u + (uint32_t)((i & 3) * 8), db + (uint64_t)((i & 3) * 2), idesc, 1u);
            else ptx::umma_f16_ss(d, da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), idesc, 1u);
          }
        }
      }
      if (leader) ptx::umma_commit(&bar);
      __syncwarp();
      const long long t1 = clock64();
      ptx::mbar_wait(&bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[2 * round] = t1 - t0; out[2 * round + 1] = t2 - t0; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tbase, 512); }
}

template <int N, int TS, int CHAINS>
static void run(long long* d) {
  const int reps = 64;
  cudaFuncSetAttribute(mma_bench<N, TS, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  mma_bench<N, TS, CHAINS><<<1, 128, 16384 + 32768 + 1024>>>(reps, d);
  long long h[6];
  cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("%s N=%3d chains=%d: issue %6.1f clk/mma, complete %6.1f clk/mma (floor %d)%s\n", TS ? "TS" : "SS", N, CHAINS,
         (double)h[4] / reps, (double)h[5] / reps, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<32, 0, 1>(d); run<64, 0, 1>(d); run<64, 0, 2>(d); run<128, 0, 1>(d); run<128, 0, 2>(d); run<192, 0, 1>(d); run<256, 0, 1>(d);
  run<32, 1, 1>(d); run<64, 1, 1>(d); run<64, 1, 2>(d); run<128, 1, 1>(d); run<128, 1, 2>(d); run<192, 1, 1>(d); run<256, 1, 1>(d);
  return 0;
}
