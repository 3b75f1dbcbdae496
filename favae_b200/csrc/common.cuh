// Shared host-side helpers of the C ABI: error reporting and launch accounting.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/favae_b200.h"

namespace favae {

char* last_error_buf();                 // thread-local, 512 bytes
void count_launch(int n = 1);

inline int fail(int code, const char* fmt, const char* a = "") {
  snprintf(last_error_buf(), 512, fmt, a);
  return code;
}

inline int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(last_error_buf(), 512, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define FAVAE_CUDA_OK(expr)                                                          \
  do {                                                                               \
    cudaError_t e__ = (expr);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      snprintf(favae::last_error_buf(), 512, "%s: %s", #expr, cudaGetErrorString(e__)); \
      return (int)e__;                                                               \
    }                                                                                \
  } while (0)

#define FAVAE_REQUIRE(cond, msg)                                  \
  do {                                                            \
    if (!(cond)) return favae::fail(-22, "favae_b200: %s", msg);  \
  } while (0)

// Lazily computed launch state (cudaFuncSetAttribute done, resident cluster count, SM count) is
// per DEVICE: a process may drive several GPUs, so every such cache is an array indexed by the
// current device ordinal (one writer per device: the ABI is called from one thread per device).
constexpr int MAX_DEVICES = 64;
inline int device_ordinal() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0;
  return dev;
}
template <class T> struct PerDevice {
  T v[MAX_DEVICES] = {};
  T& here() { return v[device_ordinal()]; }
};

inline int num_sms() {
  static PerDevice<int> cache;
  int& sms = cache.here();
  if (!sms) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_ordinal());
    if (sms <= 0) sms = 148;
  }
  return sms;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace favae
