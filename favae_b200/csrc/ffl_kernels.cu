// CUDA instantiation of the fused spectrum-loss phases (ffl_driver.cuh) for sm_100a.
// One CTA (N <= 128) or one cluster exchanging columns through distributed shared memory
// (N = 256: 2 CTAs, N = 512: 8 CTAs) owns a map from the first load to the gradient store: pred and target
// are read once, both gradients written once, nothing else touches HBM.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "ffl_configs.cuh"
#include "ffl_driver.cuh"

namespace cg = cooperative_groups;

namespace favae {

template <class Cfg> struct DevEnv {
  ThreadRegs<Cfg> r;
  float2* s_;
  float2* stg_;
  float* fb_;
  unsigned int* tab_;
  int rank_;

  template <class F> __device__ __forceinline__ void for_threads(F f) { f(rank_, (int)threadIdx.x); }
  __device__ __forceinline__ void sync_warp() { __syncwarp(); }
  __device__ __forceinline__ void sync_cta() { __syncthreads(); }
  __device__ __forceinline__ void sync_cluster() {
    if constexpr (Cfg::C == 1) __syncthreads();
    else {
#ifdef FAVAE_FFL_LIGHT_SYNC   // timing experiment only: CTA-scope fence + relaxed arrive (not a valid hand-over)
      asm volatile("fence.acq_rel.cta;" ::: "memory");
      asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
#else
      cg::this_cluster().sync();
#endif
    }
  }
  // split barrier: arrive = "my reads of S are done", wait = "everybody's are" (C == 1: a CTA barrier)
  __device__ __forceinline__ void cluster_arrive() {
    if constexpr (Cfg::C > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  }
  // arrive without a fence: only valid when every access the peers wait for has already completed
  __device__ __forceinline__ void cluster_arrive_relaxed() {
#ifdef FAVAE_FFL_NO_RELAXED
    cluster_arrive();
#else
    if constexpr (Cfg::C > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#endif
  }
  __device__ __forceinline__ void cluster_wait() {
    if constexpr (Cfg::C == 1) __syncthreads();
    else asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  // phase stamps for profiles/ffl_phase_timing.cu (compiled out of the library)
  __device__ __forceinline__ void mark(int k) {
#ifdef FAVAE_FFL_TIMING
    if (threadIdx.x == 0) {
      const long long t = clock64();
      if (k >= 0) phase_acc[k] += t - phase_last;
      phase_last = t;
    }
#else
    (void)k;
#endif
  }
#ifdef FAVAE_FFL_TIMING
  long long* phase_acc;
  long long phase_last;
#endif
  __device__ __forceinline__ ThreadRegs<Cfg>& regs(int, int) { return r; }
  __device__ __forceinline__ float2* S(int, int owner) {
    if constexpr (Cfg::C == 1) return s_;
    else return (owner == rank_) ? s_ : cg::this_cluster().map_shared_rank(s_, owner);
  }
  // S addresses for the row-FFT scatter / gather: float2 index (one CTA) or the 32-bit
  // shared::cluster byte address of the column inside its owner CTA (clusters)
  __device__ __forceinline__ unsigned int s_entry(int, int owner, int off) {
    if constexpr (Cfg::C == 1) return (unsigned int)off;
    else {
      unsigned int local = (unsigned int)__cvta_generic_to_shared(s_ + off), remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(owner));
      return remote;
    }
  }
  __device__ __forceinline__ void s_put(int, unsigned int e, int at, float2 v) {
    if constexpr (Cfg::C == 1) s_[e + at] = v;
    else asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(e + 8u * (unsigned int)at), "f"(v.x), "f"(v.y) : "memory");
  }
  __device__ __forceinline__ float2 s_get(int, unsigned int e, int at) {
    if constexpr (Cfg::C == 1) return s_[e + at];
    else {
      float2 v;
      asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(e + 8u * (unsigned int)at) : "memory");
      return v;
    }
  }
  __device__ __forceinline__ unsigned int s_entry_add(unsigned int e, int d) {
    if constexpr (Cfg::C == 1) return e + (unsigned int)d;
    else return e + 8u * (unsigned int)d;
  }
  template <int D> __device__ __forceinline__ void s_put_d(int, unsigned int e, int at, float2 v) {
    if constexpr (Cfg::C == 1) s_[e + at + D] = v;
    else asm volatile("st.shared::cluster.v2.f32 [%0+%1], {%2, %3};" ::"r"(e + 8u * (unsigned int)at), "n"(8 * D), "f"(v.x), "f"(v.y) : "memory");
  }
  template <int D> __device__ __forceinline__ float2 s_get_d(int, unsigned int e, int at) {
    if constexpr (Cfg::C == 1) return s_[e + at + D];
    else {
      float2 v;
      asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(e + 8u * (unsigned int)at), "n"(8 * D) : "memory");
      return v;
    }
  }
  __device__ __forceinline__ float2* stg(int) { return stg_; }
  __device__ __forceinline__ float* fbuf(int) { return fb_; }
  __device__ __forceinline__ unsigned int* tab(int) { return tab_; }
  __device__ __forceinline__ float* cl(int, int owner) {
    float* base = fb_ + 4 * Cfg::THREADS + 8 * Cfg::MPC;
    if constexpr (Cfg::C == 1) return base;
    else return (owner == rank_) ? base : cg::this_cluster().map_shared_rank(base, owner);
  }
  __device__ __forceinline__ void prefetch_l2(const void* ptr, size_t bytes) {
#ifndef FAVAE_FFL_NO_L2PF
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"((unsigned)bytes) : "memory");
#endif
  }
  __device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  __device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, size_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"((unsigned)bytes) : "memory");
  }
  __device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
  __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
  __device__ __forceinline__ float2 twiddle(int j, int n) {
    float s, c;
    sincospif(-2.0f * (float)j / (float)n, &s, &c);
    return make_float2(c, s);
  }
};

#ifdef FAVAE_FFL_TIMING
__device__ long long favae_ffl_phase_cycles[16 * 1024];
#endif

template <class Cfg, bool FAST, bool DIFF = false>
__global__ void __launch_bounds__(Cfg::THREADS, (Cfg::THREADS <= 256 && Cfg::SMEM_BYTES < 110 * 1024) ? 2 : 1)
ffl_kernel(const FflParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DevEnv<Cfg> env;
  env.s_ = reinterpret_cast<float2*>(smem_raw);
  env.stg_ = env.s_ + Cfg::S_FLOAT2;
  env.fb_ = reinterpret_cast<float*>(env.stg_ + Cfg::STG_FLOAT2);
  env.tab_ = reinterpret_cast<unsigned int*>(env.fb_ + 4 * Cfg::THREADS + 8 * Cfg::MPC + 8 * Cfg::C);
  if constexpr (Cfg::C == 1) env.rank_ = 0;
  else env.rank_ = (int)cg::this_cluster().block_rank();
#ifdef FAVAE_FFL_TIMING
  __shared__ long long phase_acc_s[16];
  if (threadIdx.x < 16) phase_acc_s[threadIdx.x] = 0;
  env.phase_acc = phase_acc_s;
  env.phase_last = 0;
#endif
  ffl_init_thread<Cfg>(env);
  env.cluster_arrive();                          // opens the split barrier the first batch waits on
  const long long batches = (p.maps + Cfg::MPC - 1) / Cfg::MPC;
  const long long stride = gridDim.x / Cfg::C;
  const long long first = blockIdx.x / Cfg::C;
  if (FflPipe<Cfg, DIFF>::value && first < batches) ffl_issue_loads<Cfg, DIFF>(env, p, first, 0);
  for (long long b = first; b < batches; b += stride)
    ffl_map_batch<Cfg, FAST, DIFF>(env, p, b, b + stride < batches ? b + stride : -1);
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staged gradient rows have left shared memory
  env.cluster_wait();                            // nobody leaves while a peer may still read its S
#ifdef FAVAE_FFL_TIMING
  __syncthreads();
  if (threadIdx.x < 16) favae_ffl_phase_cycles[blockIdx.x * 16 + threadIdx.x] = phase_acc_s[threadIdx.x];
#endif
}

template <class Cfg, bool FAST, bool DIFF> static int launch_ffl_impl(const FflParams& p, cudaStream_t stream) {
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.here();
  auto kern = ffl_kernel<Cfg, FAST, DIFF>;
  if (!configured) {
    FAVAE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Cfg::SMEM_BYTES));
    configured = true;
  }
  const long long batches = (p.maps + Cfg::MPC - 1) / Cfg::MPC;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = Cfg::C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent grid: exactly the clusters that can be co-resident
  static PerDevice<int> resident_dev;
  int& resident = resident_dev.here();
  if (!resident) {
    int per_sm = (int)(232448 / (Cfg::SMEM_BYTES + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm * Cfg::THREADS > 2048) per_sm = 2048 / Cfg::THREADS;
    int guess = num_sms() * per_sm / Cfg::C;
    if (Cfg::C > 1) {
      cfg.gridDim = dim3((unsigned)(guess * Cfg::C));
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) guess = n;
      else (void)cudaGetLastError();
    }
    resident = guess;
  }
  long long clusters = resident;
  if (clusters > batches) clusters = batches;
  cfg.gridDim = dim3((unsigned)(clusters * Cfg::C));
  FAVAE_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  return check_launch("ffl_kernel");
}

// alpha == 1 without log weighting (every call site of the reference) takes the lean statistics path
template <class Cfg> static int launch_ffl(const FflParams& p, cudaStream_t stream) {
  const bool fast = p.alpha == 1.0f && !p.log_matrix && p.grad_scale >= 0.0f;
  if (p.target == nullptr)       // pred holds the difference map (fused DSL level)
    return fast ? launch_ffl_impl<Cfg, true, true>(p, stream) : launch_ffl_impl<Cfg, false, true>(p, stream);
  return fast ? launch_ffl_impl<Cfg, true, false>(p, stream) : launch_ffl_impl<Cfg, false, false>(p, stream);
}

}  // namespace favae

extern "C" {

int favae_ffl_supported(int h, int w) {
  return (h == w) && (h == 8 || h == 16 || h == 32 || h == 64 || h == 128 || h == 256 || h == 512);
}

int favae_ffl_forward(const float* pred, const float* target, int64_t maps, int h, int w,
                      float alpha, int log_matrix, float grad_scale, float* map_loss,
                      float* grad_pred, float* grad_target, float* map_max,
                      const float* fmax_override, void* stream) {
  using namespace favae;
  FAVAE_REQUIRE(pred && map_loss, "ffl_forward: null pointer");
  FAVAE_REQUIRE(target || !grad_target, "ffl_forward: a gradient for a target that was not given");
  FAVAE_REQUIRE(favae_ffl_supported(h, w), "ffl_forward: maps must be square, side a power of two in [8,512]");
  FAVAE_REQUIRE(maps >= 0, "ffl_forward: negative map count");
  FAVAE_REQUIRE((((uintptr_t)pred | (uintptr_t)target | (uintptr_t)grad_pred | (uintptr_t)grad_target) & 15) == 0,
                "ffl_forward: tensors must be 16-byte aligned");
  if (maps == 0) return 0;
  FflParams p;
  p.pred = pred; p.target = target; p.grad_pred = grad_pred; p.grad_target = grad_target;
  p.map_loss = map_loss; p.maps = maps; p.alpha = alpha; p.log_matrix = log_matrix;
  p.map_max = map_max; p.fmax_override = fmax_override;
  p.grad_scale = grad_scale / (float)(h * w);
  cudaStream_t s = (cudaStream_t)stream;
  switch (h) {
    case 8: return launch_ffl<FflCfg8>(p, s);
    case 16: return launch_ffl<FflCfg16>(p, s);
    case 32: return launch_ffl<FflCfg32>(p, s);
    case 64: return launch_ffl<FflCfg64>(p, s);
    case 128: return launch_ffl<FflCfg128>(p, s);
    case 512: return launch_ffl<FflCfg512>(p, s);   // 8-CTA cluster: 18 maps in flight on 144 SMs
    default: {
      // default: 2-CTA cluster, one map per SM pair.  FAVAE_FFL256=c4 selects the 4-CTA-cluster
      // variant (two maps in flight per SM), measured 14 % slower in round 1 (profiles/)
      static const bool c4 = [] { const char* e = getenv("FAVAE_FFL256"); return e && e[0] == 'c' && e[1] == '4'; }();
      return c4 ? launch_ffl<FflCfg256c4>(p, s) : launch_ffl<FflCfg256>(p, s);
    }
  }
}
}
