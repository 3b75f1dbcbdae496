// Phase driver of the fused spectrum-loss kernel (see ffl_core.cuh for the math).
//
// `Env` abstracts the execution model so that the same phases run as a CUDA thread
// block / cluster (ffl_kernels.cu) and as plain loops on the host (tests/emul):
//   env.for_threads(f)   f(cta, tid) for every thread of every CTA of the cluster
//   env.sync_warp/cta/cluster()
//   env.regs(cta, tid)   ThreadRegs that persist across sync points
//   env.S(cta, owner)    float2* to the spectrum buffer of CTA `owner` as seen from `cta`
//   env.stg(cta)         staging for the two-stage 1-D FFT exchange
//   env.fbuf(cta)        float scratch: [0,2T) thread partials, [2T,4T) lane partials,
//                        [4T, 4T+8*MPC) per-map results, then 8*C cluster slots
//   env.cl(cta, owner)   float* to the cluster slots of CTA `owner`
//   env.twiddle(j, n)    e^{-2 pi i j / n}
#pragma once

#include "ffl_core.cuh"

namespace favae {

template <class Cfg, class Env>
FAVAE_HD void ffl_init_thread(Env& env) {
  env.for_threads([&](int cta, int tid) {
    ThreadRegs<Cfg>& r = env.regs(cta, tid);
    const int t = tid % Cfg::TG;
#pragma unroll
    for (int k1 = 0; k1 < Cfg::R1; ++k1) r.tw[k1] = env.twiddle(t * k1, Cfg::N);
  });
}

// One batch = MPC maps (C == 1) or one map shared by the C CTAs of a cluster.
template <class Cfg, class Env>
FAVAE_HD void ffl_map_batch(Env& env, const FflParams& p, long long batch) {
  constexpr int N = Cfg::N, R1 = Cfg::R1, TG = Cfg::TG, NG = Cfg::NG, HALF = Cfg::HALF;
  constexpr int C = Cfg::C, MPC = Cfg::MPC, T = Cfg::THREADS, PASSES = Cfg::PASSES;
  constexpr int GPC = HALF / C;                 // row pairs / column groups per CTA and map
  constexpr int TMAP = T / MPC;                 // threads per map slot
  constexpr int STG = R1 * Cfg::STG_STRIDE;
  static_assert(MPC == 1 || PASSES == 1, "several maps per CTA need a single pass");
  const float inv_nn = 1.0f / (float)(N * N);
  const long long map0 = batch * MPC;

  env.for_threads([&](int cta, int tid) {
    ThreadRegs<Cfg>& r = env.regs(cta, tid);
    r.sum = 0.0f; r.mx = 0.0f;
  });

  // ---------------- P1: packed row FFTs, global -> S ----------------
  for (int pass = 0; pass < PASSES; ++pass) {
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, rp = cta * GPC + item % GPC;
      const long long map = map0 + m;
      if (map < p.maps) {
        const float* pa = p.pred + map * (long long)(N * N) + rp * N;
        const float* ta = p.target + map * (long long)(N * N) + rp * N;
#pragma unroll
        for (int e = 0; e < R1; ++e) {
          const int c = idx_in<Cfg>(t, e);
          r.v[e] = make_float2(pa[c] - ta[c], pa[HALF * N + c] - ta[HALF * N + c]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < R1; ++e) r.v[e] = make_float2(0.f, 0.f);
      }
      fwd_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
    });
    env.sync_warp();
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, rp = cta * GPC + item % GPC;
      fwd_stage2<Cfg>(r, t, env.stg(cta) + g * STG);
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        int owner, off;
        s_locate<Cfg>(idx_out<Cfg>(t, e), m, owner, off);
        env.S(cta, owner)[off + rp] = r.v[e];
      }
    });
    env.sync_warp();
  }
  env.sync_cluster();

  // ---------------- P2: column FFTs + spectrum statistics ----------------
  for (int pass = 0; pass < PASSES; ++pass) {
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, v = cta * GPC + item % GPC;
      int o0, o1, off0, off1;
      s_locate<Cfg>(v == 0 ? 0 : v, m, o0, off0);
      s_locate<Cfg>(v == 0 ? HALF : N - v, m, o1, off1);
      const float2* S = env.S(cta, cta);
#pragma unroll
      for (int e = 0; e < R1 / 2; ++e) {
        const int rr = idx_in<Cfg>(t, e);
        const float2 zv = S[off0 + rr], zw = S[off1 + rr];
        if (v == 0) {
          r.v[e] = make_float2(zv.x, zw.x);
          r.v[e + R1 / 2] = make_float2(zv.y, zw.y);
        } else {
          r.v[e] = make_float2(0.5f * (zv.x + zw.x), 0.5f * (zv.y - zw.y));
          r.v[e + R1 / 2] = make_float2(0.5f * (zv.y + zw.y), 0.5f * (zw.x - zv.x));
        }
      }
      fwd_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
    });
    env.sync_warp();
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, v = cta * GPC + item % GPC;
      int o0, o1, off0, off1;
      s_locate<Cfg>(v == 0 ? 0 : v, m, o0, off0);
      s_locate<Cfg>(v == 0 ? HALF : N - v, m, o1, off1);
      float2* S = env.S(cta, cta);
      fwd_stage2<Cfg>(r, t, env.stg(cta) + g * STG);
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        const int u = idx_out<Cfg>(t, e);
        const float2 z = r.v[e];
        if (v != 0) {
          const float a2 = (z.x * z.x + z.y * z.y) * inv_nn;
          const float f = spectrum_f(a2, p.alpha, p.log_matrix);
          r.sum += 2.0f * f * a2;
          r.mx = fmaxf(r.mx, f);
        }
        S[(u < HALF) ? off0 + u : off1 + (u - HALF)] = z;
      }
    });
    env.sync_warp();
  }
  env.sync_cta();

  // ---------------- P3: statistics of the packed columns v = 0 and v = N/2 ----------------
  env.for_threads([&](int cta, int tid) {
    ThreadRegs<Cfg>& r = env.regs(cta, tid);
    if (cta == 0) {
      const int m = tid / TMAP;
      int o, off0, off1;
      s_locate<Cfg>(0, m, o, off0);
      s_locate<Cfg>(HALF, m, o, off1);
      const float2* S = env.S(cta, cta);
      for (int u = tid % TMAP; u <= HALF; u += TMAP) {
        const int un = (N - u) % N;
        const float2 a = S[(u < HALF) ? off0 + u : off1 + (u - HALF)];
        const float2 b = S[(un < HALF) ? off0 + un : off1 + (un - HALF)];
        const float2 d0 = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
        const float2 dn = make_float2(0.5f * (a.y + b.y), 0.5f * (b.x - a.x));
        const float mult = (u == 0 || u == HALF) ? 1.0f : 2.0f;
        const float a0 = (d0.x * d0.x + d0.y * d0.y) * inv_nn;
        const float an = (dn.x * dn.x + dn.y * dn.y) * inv_nn;
        const float f0 = spectrum_f(a0, p.alpha, p.log_matrix);
        const float fn = spectrum_f(an, p.alpha, p.log_matrix);
        r.sum += mult * (f0 * a0 + fn * an);
        r.mx = fmaxf(r.mx, fmaxf(f0, fn));
      }
    }
    float* fb = env.fbuf(cta);
    fb[tid] = r.sum;
    fb[T + tid] = r.mx;
  });
  env.sync_cta();
  // deterministic two-level reduction per map slot
  env.for_threads([&](int cta, int tid) {
    float* fb = env.fbuf(cta);
    const int m = tid / TMAP, j = tid % TMAP;
    constexpr int L = TMAP < 32 ? TMAP : 32;
    if (j < L) {
      float s = 0.f, mx = 0.f;
      for (int i = j; i < TMAP; i += L) { s += fb[m * TMAP + i]; mx = fmaxf(mx, fb[T + m * TMAP + i]); }
      fb[2 * T + tid] = s;
      fb[3 * T + tid] = mx;
    }
  });
  env.sync_cta();
  env.for_threads([&](int cta, int tid) {
    float* fb = env.fbuf(cta);
    const int m = tid / TMAP, j = tid % TMAP;
    constexpr int L = TMAP < 32 ? TMAP : 32;
    if (j == 0) {
      float s = 0.f, mx = 0.f;
      for (int i = 0; i < L; ++i) { s += fb[2 * T + m * TMAP + i]; mx = fmaxf(mx, fb[3 * T + m * TMAP + i]); }
      if (C == 1) {
        fb[4 * T + 2 * m] = s;
        fb[4 * T + 2 * m + 1] = p.fmax_override ? p.fmax_override[0] : mx;
        fb[4 * T + 2 * MPC + m] = mx;
      } else {
        for (int o = 0; o < C; ++o) {
          float* cl = env.cl(cta, o);
          cl[2 * cta] = s;
          cl[2 * cta + 1] = mx;
        }
      }
    }
  });
  env.sync_cluster();
  env.for_threads([&](int cta, int tid) {
    float* fb = env.fbuf(cta);
    if (C > 1 && tid == 0) {
      const float* cl = env.cl(cta, cta);
      float s = 0.f, mx = 0.f;
      for (int o = 0; o < C; ++o) { s += cl[2 * o]; mx = fmaxf(mx, cl[2 * o + 1]); }
      fb[4 * T] = s;
      fb[4 * T + 1] = p.fmax_override ? p.fmax_override[0] : mx;
      fb[4 * T + 2 * MPC] = mx;
    }
  });
  if (C > 1) env.sync_cta();
  env.for_threads([&](int cta, int tid) {
    const float* fb = env.fbuf(cta);
    const int m = tid / TMAP, j = tid % TMAP;
    if (cta == 0 && j == 0 && map0 + m < p.maps) {
      const float s = fb[4 * T + 2 * m], mx = fb[4 * T + 2 * m + 1];
      p.map_loss[map0 + m] = (mx > 0.0f) ? s / mx : 0.0f;
      if (p.map_max) p.map_max[map0 + m] = fb[4 * T + 2 * MPC + m];
    }
  });
  if (p.grad_pred == nullptr && p.grad_target == nullptr) {
    env.sync_cluster();
    return;
  }

  // ---------------- P4: weight the packed columns in place ----------------
  env.for_threads([&](int cta, int tid) {
    if (cta != 0) return;
    const float* fb = env.fbuf(cta);
    const int m = tid / TMAP;
    const float fmx = fb[4 * T + 2 * m + 1];
    int o, off0, off1;
    s_locate<Cfg>(0, m, o, off0);
    s_locate<Cfg>(HALF, m, o, off1);
    float2* S = env.S(cta, cta);
    for (int u = tid % TMAP; u <= HALF; u += TMAP) {
      const int un = (N - u) % N;
      const int ia = (u < HALF) ? off0 + u : off1 + (u - HALF);
      const int ib = (un < HALF) ? off0 + un : off1 + (un - HALF);
      const float2 a = S[ia], b = S[ib];
      const float2 d0 = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      const float2 dn = make_float2(0.5f * (a.y + b.y), 0.5f * (b.x - a.x));
      const float a0 = (d0.x * d0.x + d0.y * d0.y) * inv_nn;
      const float an = (dn.x * dn.x + dn.y * dn.y) * inv_nn;
      const float w0 = spectrum_w(spectrum_f(a0, p.alpha, p.log_matrix), fmx);
      const float wn = spectrum_w(spectrum_f(an, p.alpha, p.log_matrix), fmx);
      S[ia] = make_float2(w0 * d0.x - wn * dn.y, w0 * d0.y + wn * dn.x);
      if (ib != ia) S[ib] = make_float2(w0 * d0.x + wn * dn.y, wn * dn.x - w0 * d0.y);
    }
  });
  env.sync_cta();

  // ---------------- P5: weight + inverse column FFTs, re-pack into Z' ----------------
  for (int pass = 0; pass < PASSES; ++pass) {
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const float* fb = env.fbuf(cta);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, v = cta * GPC + item % GPC;
      const float fmx = fb[4 * T + 2 * m + 1];
      int o0, o1, off0, off1;
      s_locate<Cfg>(v == 0 ? 0 : v, m, o0, off0);
      s_locate<Cfg>(v == 0 ? HALF : N - v, m, o1, off1);
      const float2* S = env.S(cta, cta);
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        const int u = idx_out<Cfg>(t, e);
        float2 z = S[(u < HALF) ? off0 + u : off1 + (u - HALF)];
        if (v != 0) {
          const float a2 = (z.x * z.x + z.y * z.y) * inv_nn;
          const float w = spectrum_w(spectrum_f(a2, p.alpha, p.log_matrix), fmx);
          z.x *= w; z.y *= w;
        }
        r.v[e] = z;
      }
      inv_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
    });
    env.sync_warp();
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, v = cta * GPC + item % GPC;
      int o0, o1, off0, off1;
      s_locate<Cfg>(v == 0 ? 0 : v, m, o0, off0);
      s_locate<Cfg>(v == 0 ? HALF : N - v, m, o1, off1);
      float2* S = env.S(cta, cta);
      inv_stage2<Cfg>(r, t, env.stg(cta) + g * STG);
#pragma unroll
      for (int e = 0; e < R1 / 2; ++e) {
        const int rr = idx_in<Cfg>(t, e);
        const float2 a = r.v[e], b = r.v[e + R1 / 2];
        if (v == 0) {
          S[off0 + rr] = make_float2(a.x, b.x);
          S[off1 + rr] = make_float2(a.y, b.y);
        } else {
          S[off0 + rr] = make_float2(a.x - b.y, a.y + b.x);
          S[off1 + rr] = make_float2(a.x + b.y, b.x - a.y);
        }
      }
    });
    env.sync_warp();
  }
  env.sync_cluster();

  // ---------------- P6: inverse row FFTs, S -> gradient rows ----------------
  for (int pass = 0; pass < PASSES; ++pass) {
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, rp = cta * GPC + item % GPC;
#pragma unroll
      for (int e = 0; e < R1; ++e) {
        int owner, off;
        s_locate<Cfg>(idx_out<Cfg>(t, e), m, owner, off);
        r.v[e] = env.S(cta, owner)[off + rp];
      }
      inv_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
    });
    env.sync_warp();
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, rp = cta * GPC + item % GPC;
      const long long map = map0 + m;
      inv_stage2<Cfg>(r, t, env.stg(cta) + g * STG);
      if (map < p.maps) {
        const long long base = map * (long long)(N * N) + rp * N;
#pragma unroll
        for (int e = 0; e < R1; ++e) {
          const int c = idx_in<Cfg>(t, e);
          const float ga = r.v[e].x * p.grad_scale, gb = r.v[e].y * p.grad_scale;
          if (p.grad_pred) { p.grad_pred[base + c] = ga; p.grad_pred[base + HALF * N + c] = gb; }
          if (p.grad_target) { p.grad_target[base + c] = -ga; p.grad_target[base + HALF * N + c] = -gb; }
        }
      }
    });
    env.sync_warp();
  }
  env.sync_cluster();
}

}  // namespace favae
