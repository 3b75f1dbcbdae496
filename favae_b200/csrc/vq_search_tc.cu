// Tensor-core (tcgen05 / TMEM / TMA) nearest-code search -- placeholder until the UMMA kernel
// lands: reports "not available" through a zero workspace size so that callers take the exact
// CUDA-core search.
#include "common.cuh"

extern "C" {
size_t favae_vq_search_tc_workspace_bytes(int64_t, int64_t, int) { return 0; }
int favae_vq_search_tc(const void*, const void*, const float*, const float*, int64_t, int64_t, int, void*,
                       size_t, uint64_t*, int64_t*, void*) {
  return favae::fail(-38, "favae_b200: %s", "vq_search_tc is not built in this library");
}
}
