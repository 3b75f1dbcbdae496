// ABI bookkeeping: version, thread-local error string, launch counter, small utilities.
#include "common.cuh"

#include <atomic>

namespace favae {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
char* last_error_buf() { return g_err; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// out = scale * sum(v) in fp64, one block, fixed reduction tree
__global__ void __launch_bounds__(1024) sum_scaled_kernel(const float* __restrict__ v, long long n,
                                                          double scale, float* __restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) acc += (double)v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(sh[0] * scale);
}

__global__ void __launch_bounds__(256) scale_inplace_kernel(float* __restrict__ a, float* __restrict__ b,
                                                            long long n4, long long n,
                                                            const float* __restrict__ s) {
  const float f = s[0];
  if (f == 1.0f) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = reinterpret_cast<float4*>(a)[i];
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    reinterpret_cast<float4*>(a)[i] = v;
    if (b) {
      float4 w = reinterpret_cast<float4*>(b)[i];
      w.x *= f; w.y *= f; w.z *= f; w.w *= f;
      reinterpret_cast<float4*>(b)[i] = w;
    }
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    a[i] *= f;
    if (b) b[i] *= f;
  }
}
}  // namespace favae

extern "C" {
int favae_abi_version(void) { return FAVAE_B200_ABI_VERSION; }
const char* favae_last_error(void) { return favae::last_error_buf(); }
long long favae_launch_count(void) { return favae::g_launches.load(); }

int favae_sum_scaled(const float* v, int64_t n, double scale, float* out, void* stream) {
  FAVAE_REQUIRE(v && out && n >= 0, "sum_scaled: bad arguments");
  favae::sum_scaled_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(v, n, scale, out);
  return favae::check_launch("sum_scaled");
}

int favae_scale_inplace(float* a, float* b, int64_t n, const float* s, void* stream) {
  FAVAE_REQUIRE(a && s && n >= 0, "scale_inplace: bad arguments");
  FAVAE_REQUIRE(((uintptr_t)a % 16 == 0) && (!b || (uintptr_t)b % 16 == 0), "scale_inplace: unaligned");
  if (n == 0) return 0;
  long long n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > favae::num_sms() * 8) blocks = favae::num_sms() * 8;
  if (blocks < 1) blocks = 1;
  favae::scale_inplace_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, b, n4, n, s);
  return favae::check_launch("scale_inplace");
}
}
