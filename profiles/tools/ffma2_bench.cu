// FFMA vs FFMA2 (fma.rn.f32x2) vs FADD2 issue throughput on sm_100a: per-SM lane-FMAs per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/tools/ffma2_bench profiles/tools/ffma2_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <int MODE>   // 0: FFMA x16 chains, 1: FFMA2 x8 chains (16 floats), 2: FADD2 x8, 3: mix FFMA2 + LDS
__global__ void bench(float* out, int iters, long long* cyc) {
  __shared__ float sh[1024];
  sh[threadIdx.x % 1024] = threadIdx.x;
  __syncthreads();
  float a[16];
  unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
  for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
  const float m = 1.0001f, c = 0.5f;
  const unsigned long long mm = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
  const unsigned long long cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma1(a[i], m, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], mm, cc);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = add2(p[i], cc);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], mm, cc); a[i] += sh[(threadIdx.x + i * 33 + it) & 1023]; }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE> void run(const char* name, int threads, float* out, long long* cyc) {
  const int iters = 4096;
  bench<MODE><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  bench<MODE><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double lane_ops = (double)threads * iters * 16;   // scalar fp32 results per CTA (= per SM)
  printf("%-28s threads/SM=%4d  %.1f fp32 results/clk/SM  (%.2f warp-instr/clk/SM)\n", name, threads, lane_ops / c,
         lane_ops / c / 32 / (MODE == 0 ? 1 : 2));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int th : {128, 256, 512, 1024}) {
    run<0>("FFMA", th, out, cyc);
    run<1>("FFMA2", th, out, cyc);
    run<2>("FADD2", th, out, cyc);
    run<3>("FFMA2 + LDS + FADD (1:1:1)", th, out, cyc);
  }
  return 0;
}
