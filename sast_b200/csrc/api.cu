// Library info entry points.
#include "common.cuh"
#include <atomic>
#include <cstdlib>

static std::atomic<unsigned long long> g_launches{0};
extern "C" void sast_count_launch_(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" uint64_t sast_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int sast_abi_version(void) { return SAST_ABI_VERSION; }

extern "C" const char* sast_build_info(void) {
  return "libsast_b200 abi " "2" " sm_100a nvcc " __DATE__;
}

// sizeof() of the ABI structs as this library was compiled, so that a binding can verify its mirror.
extern "C" size_t sast_struct_size(int32_t which) {
  switch (which) {
    case 0: return sizeof(sast_geom);
    case 1: return sizeof(sast_selection);
    case 2: return sizeof(sast_score_args);
    case 3: return sizeof(sast_select_args);
    case 4: return sizeof(sast_layer_weights);
    case 5: return sizeof(sast_layer_args);
    case 6: return sizeof(sast_layer_grads);
  }
  return 0;
}

namespace sast {
long long* g_trace = nullptr;   // debug: see sast_debug_trace
int g_trace_which = 0;
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SAST_B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
}  // namespace sast

// Debug aid (tools/attn_trace.py, tools/gemm_trace.py): while buf is non-null the tensor-core attention and GEMM
// kernels (which: 1 = attention, 2 = GEMM) of the trace build write clock64 stamps of their phase boundaries into it
// (layout: see the kernels).  Null = off (default).  The regular build ignores it.
extern "C" void sast_debug_trace(long long* buf, int32_t which) {
  sast::g_trace = buf;
  sast::g_trace_which = buf ? which : 0;
}
