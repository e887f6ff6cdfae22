// Device helpers shared by the one-kernel MS-WSA layers (fused_layer.cu: one CTA per tile; group_layer.cu: a group of
// CTAs per tile): bf16 packing, shared-memory vector accesses, UMMA descriptors for the SWIZZLE_64B attention operand
// tiles, k-block MMA issue, pinned index loads, the tile-row mapping of "aligned" tiles.
#pragma once
#include "layer.cuh"
#include "ptx.cuh"

namespace sast {
namespace fl {

// tile row -> offset of its compacted row inside the tile's row range, or -1.  split > 0 ("aligned" tile: two windows of
// <= 64 tokens): the second window starts at tile row 64, so every row's keys lie inside one 64-column block of S.
__device__ __forceinline__ int tile_src(int r, int rows, int split) {
  if (split == 0) return r < rows ? r : -1;
  if (r < 64) return r < split ? r : -1;
  const int s = split + (r - 64);
  return s < rows ? s : -1;
}

template <int TPC>
__device__ __forceinline__ void ctx_sync(int ctx) {
  asm volatile("bar.sync %0, %1;" ::"r"(ctx + 1), "n"(TPC) : "memory");
}

__device__ __forceinline__ float group4_sum(float v) {
  v += __shfl_xor_sync(kFull, v, 1);
  v += __shfl_xor_sync(kFull, v, 2);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// index load that stays where it is written: the compiler may neither sink it to its first use nor hoist it
__device__ __forceinline__ int ldg_pinned(const int* p) {
  int v;
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// K-major SWIZZLE_64B operand (rows of 32 bf16, 8-row atoms 512 bytes apart): Q and K tiles
__device__ __forceinline__ uint64_t desc_sw64_k(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}
// MN-major SWIZZLE_64B operand (one key per 64-byte row of 32 output dims): V tiles as the B operand of P V
__device__ __forceinline__ uint64_t desc_sw64_mn(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__host__ __device__ constexpr uint32_t idesc(uint32_t M, uint32_t N, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// one 64-column k-block (nks <= 4 k-steps of 16) of D[128 x N] (+)= A B^T, both operands SWIZZLE_128B K-major
__device__ __forceinline__ void mma_kblock(bool leader, uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t id,
                                           int nks, bool fresh) {
  const uint64_t da = ptx::umma_desc_sw128_kmajor(a_addr), db = ptx::umma_desc_sw128_kmajor(b_addr);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nks && leader) ptx::umma_f16_ss(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), id, (k > 0 || !fresh) ? 1u : 0u);
}

}  // namespace fl
}  // namespace sast
