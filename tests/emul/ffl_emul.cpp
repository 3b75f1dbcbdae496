// Host emulation of favae_b200/csrc/ffl_driver.cuh: the CUDA phases executed as plain
// loops over (cta, tid).  TEST INFRASTRUCTURE -- validates the FFT index math, packing
// and reductions on a machine without a GPU (tests/test_ffl_emulation.py).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../favae_b200/csrc/ffl_configs.cuh"
#include "../../favae_b200/csrc/ffl_driver.cuh"

using namespace favae;

template <class Cfg> struct HostEnv {
  std::vector<ThreadRegs<Cfg>> regs_;
  std::vector<float2> S_, stg_;
  std::vector<float> fb_;
  std::vector<unsigned int> tab_;
  static constexpr int FB = 4 * Cfg::THREADS + 8 * Cfg::MPC + 8 * Cfg::C;
  HostEnv()
      : regs_(Cfg::C * Cfg::THREADS), S_((size_t)Cfg::C * Cfg::S_FLOAT2),
        stg_((size_t)Cfg::C * (Cfg::STG_FLOAT2 + 2)), fb_((size_t)Cfg::C * FB), tab_((size_t)Cfg::C * Cfg::N) {
    // poison so that reads of never-written slots show up
    for (auto& z : S_) { z.x = NAN; z.y = NAN; }
    for (auto& z : stg_) { z.x = NAN; z.y = NAN; }
  }
  template <class F> void for_threads(F f) {
    for (int cta = 0; cta < Cfg::C; ++cta)
      for (int tid = 0; tid < Cfg::THREADS; ++tid) f(cta, tid);
  }
  void sync_warp() {}
  void sync_cta() {}
  void sync_cluster() {}
  void cluster_arrive() {}
  void cluster_arrive_relaxed() {}
  void cluster_wait() {}
  void mark(int) {}
  unsigned int s_entry(int, int owner, int off) { return ((unsigned int)owner << 24) | (unsigned int)off; }
  void s_put(int cta, unsigned int e, int at, float2 v) { S(cta, (int)(e >> 24))[(int)(e & 0xFFFFFFu) + at] = v; }
  float2 s_get(int cta, unsigned int e, int at) { return S(cta, (int)(e >> 24))[(int)(e & 0xFFFFFFu) + at]; }
  unsigned int s_entry_add(unsigned int e, int d) { return e + (unsigned int)d; }
  template <int D> void s_put_d(int cta, unsigned int e, int at, float2 v) { s_put(cta, e, at + D, v); }
  template <int D> float2 s_get_d(int cta, unsigned int e, int at) { return s_get(cta, e, at + D); }
  void prefetch_l2(const void*, size_t) {}
  void fence_async() {}
  void bulk_store(void* gdst, const void* ssrc, size_t bytes) { std::memcpy(gdst, ssrc, bytes); }
  void bulk_commit() {}
  void bulk_wait_read() {}
  ThreadRegs<Cfg>& regs(int cta, int tid) { return regs_[cta * Cfg::THREADS + tid]; }
  float2* S(int, int owner) { return S_.data() + (size_t)owner * Cfg::S_FLOAT2; }
  float2* stg(int cta) { return stg_.data() + (size_t)cta * (Cfg::STG_FLOAT2 + 2); }
  unsigned int* tab(int cta) { return tab_.data() + (size_t)cta * Cfg::N; }
  float* fbuf(int cta) { return fb_.data() + (size_t)cta * FB; }
  float* cl(int, int owner) { return fbuf(owner) + 4 * Cfg::THREADS + 8 * Cfg::MPC; }
  float2 twiddle(int j, int n) {
    const double a = -2.0 * M_PI * (double)j / (double)n;
    return make_float2((float)std::cos(a), (float)std::sin(a));
  }
};

template <class Cfg, bool DIFF> static void run_form(const FflParams& p) {
  HostEnv<Cfg> env;
  ffl_init_thread<Cfg>(env);
  const long long batches = (p.maps + Cfg::MPC - 1) / Cfg::MPC;
  const bool fast = p.alpha == 1.0f && !p.log_matrix && p.grad_scale >= 0.0f;     // same dispatch as ffl_kernels.cu
  if (FflPipe<Cfg, DIFF>::value && batches > 0) ffl_issue_loads<Cfg, DIFF>(env, p, 0, 0);
  for (long long b = 0; b < batches; ++b) {
    const long long nb = b + 1 < batches ? b + 1 : -1;
    if (fast) ffl_map_batch<Cfg, true, DIFF>(env, p, b, nb);
    else ffl_map_batch<Cfg, false, DIFF>(env, p, b, nb);
  }
}
// target == nullptr: the single-input form (pred is the difference map), as in ffl_kernels.cu
template <class Cfg> static void run(const FflParams& p) {
  if (p.target == nullptr) run_form<Cfg, true>(p);
  else run_form<Cfg, false>(p);
}

extern "C" int ffl_emul(int n, const float* pred, const float* target, long long maps, float alpha,
                        int log_matrix, float grad_scale, float* gp, float* gt, float* map_loss,
                        float* map_max, const float* fmax_override) {
  FflParams p;
  p.pred = pred; p.target = target; p.grad_pred = gp; p.grad_target = gt; p.map_loss = map_loss;
  p.map_max = map_max; p.fmax_override = fmax_override;
  p.maps = maps; p.grad_scale = grad_scale; p.alpha = alpha; p.log_matrix = log_matrix;
  switch (n) {
    case 8: run<FflCfg8>(p); break;
    case 16: run<FflCfg16>(p); break;
    case 32: run<FflCfg32>(p); break;
    case 64: run<FflCfg64>(p); break;
    case 128: run<FflCfg128>(p); break;
    case 256: run<FflCfg256>(p); break;
    case 2564: run<FflCfg256c4>(p); break;
    case 512: run<FflCfg512>(p); break;
    default: return -1;
  }
  return 0;
}
