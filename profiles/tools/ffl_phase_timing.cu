// Per-phase cycle breakdown of ffl_kernel<256> (warp 0 of every CTA stamps clock64 at the phase
// boundaries of ffl_driver.cuh).  Diagnostic binary, not part of the library:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -DFAVAE_FFL_TIMING -o profiles/tools/ffl_phase_timing profiles/tools/ffl_phase_timing.cu
#include "../../favae_b200/csrc/capi_common.cu"
#include "../../favae_b200/csrc/ffl_kernels.cu"

#include <vector>

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 8;
  const bool diff = argc > 2 && atoi(argv[2]) != 0;      // single-input form: d -> G (fused DSL level)
  const long long maps = (long long)B * 128, E = maps * 256 * 256;
  float *p, *t, *gp, *gt, *ml;
  cudaMalloc(&p, E * 4); cudaMalloc(&t, E * 4); cudaMalloc(&gp, E * 4); cudaMalloc(&gt, E * 4);
  cudaMalloc(&ml, maps * 4);
  std::vector<float> h(E);
  unsigned s = 12345u;
  for (long long i = 0; i < E; ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }
  cudaMemcpy(p, h.data(), E * 4, cudaMemcpyHostToDevice);
  for (long long i = 0; i < E; ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }
  cudaMemcpy(t, h.data(), E * 4, cudaMemcpyHostToDevice);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 3; ++it) {
    std::vector<long long> zero(16 * 1024, 0);
    cudaMemcpyToSymbol(favae::favae_ffl_phase_cycles, zero.data(), zero.size() * 8);
    cudaEventRecord(a);
    int rc = diff ? favae_ffl_forward(p, nullptr, maps, 256, 256, 1.0f, 0, 1e-3f, ml, gp, nullptr, nullptr, nullptr, nullptr)
                  : favae_ffl_forward(p, t, maps, 256, 256, 1.0f, 0, 1e-3f, ml, gp, gt, nullptr, nullptr, nullptr);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("rc=%d  %.3f ms  %.1f GB/s algorithmic\n", rc, ms, (diff ? 8.0 : 16.0) * E / ms / 1e6);
  }
  {  // checksums, to compare build variants
    std::vector<float> hl(maps), hg(65536);
    cudaMemcpy(hl.data(), ml, maps * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hg.data(), gp + (maps - 1) * 65536, 65536 * 4, cudaMemcpyDeviceToHost);
    double sl = 0, sg = 0;
    for (float v : hl) sl += v;
    for (int i = 0; i < 65536; ++i) sg += (double)hg[i] * ((i * 2654435761u >> 16) & 1023);
    printf("checksum loss %.9e grad %.9e\n", sl, sg);
  }
  std::vector<long long> acc(16 * 1024);
  cudaMemcpyFromSymbol(acc.data(), favae::favae_ffl_phase_cycles, acc.size() * 8);
  const char* names[11] = {"P1 load+stage", "P1 rowFFT+S scatter", "sync_cluster", "P2 colFFT+stats", "P3 reduce+syncs",
                           "P4 packed cols", "P5 weight+icolFFT", "sync_cluster", "P6 S gather", "P6 irowFFT+stage", "P6 store"};
  double tot[16] = {0}; double all = 0; int ctas = 0;
  for (int c = 0; c < 1024; ++c) {
    long long sum = 0; for (int k = 0; k < 11; ++k) sum += acc[c * 16 + k];
    if (!sum) continue;
    ++ctas;
    for (int k = 0; k < 11; ++k) { tot[k] += acc[c * 16 + k]; all += acc[c * 16 + k]; }
  }
  printf("%d CTAs; mean cycles per CTA %.0f\n", ctas, all / ctas);
  for (int k = 0; k < 11; ++k) printf("  %-22s %6.2f %%\n", names[k], 100.0 * tot[k] / all);
  return 0;
}
