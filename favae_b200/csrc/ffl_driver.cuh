// Phase driver of the fused spectrum-loss kernel (see ffl_core.cuh for the math).
//
// `Env` abstracts the execution model so that the same phases run as a CUDA thread
// block / cluster (ffl_kernels.cu) and as plain loops on the host (tests/emul):
//   env.for_threads(f)   f(cta, tid) for every thread of every CTA of the cluster
//   env.sync_warp/cta/cluster()   env.cluster_arrive[_relaxed]() / env.cluster_wait(): split cluster barrier
//   env.regs(cta, tid)   ThreadRegs that persist across sync points
//   env.S(cta, owner)    float2* to the spectrum buffer of CTA `owner` as seen from `cta`
//   env.stg(cta)         staging for the two-stage 1-D FFT exchange
//   env.fbuf(cta)        float scratch: [0,2T) thread partials, [2T,4T) lane partials,
//                        [4T, 4T+8*MPC) per-map results, then 8*C cluster slots
//   env.cl(cta, owner)   float* to the cluster slots of CTA `owner`
//   env.tab(cta)         per map column w: env.s_entry(cta, owner, off), the S address of entry 0 of
//                        that column in whatever form env.s_put / env.s_get take (the device uses
//                        32-bit shared::cluster addresses, so a row-FFT lane reaches either CTA with
//                        one table load and one add); built by ffl_init_thread
//   env.s_put(cta, entry, at, v) / env.s_get(cta, entry, at)   S[column of entry][at] across the cluster
//   env.s_entry_add(entry, d) / env.s_put_d<D>(...) / env.s_get_d<D>(...)   the same, D column slots further on
//                        (a compile-time distance: an immediate on the device)
//   env.twiddle(j, n)    e^{-2 pi i j / n}
//   env.prefetch_l2(ptr, bytes)   hint: bring a 16-byte-aligned global range into L2
//   env.fence_async() / env.bulk_store(gdst, ssrc, bytes) / env.bulk_commit() / env.bulk_wait_read():
//                        bulk async copy of a staged row to global memory (the host copies at once)
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
    unsigned int* tab = env.tab(cta);
    for (int w = tid; w < Cfg::N; w += Cfg::THREADS) {
      int owner, off;
      s_locate<Cfg>(w, 0, owner, off);
      tab[w] = env.s_entry(cta, owner, off);
    }
    if constexpr (SFast<Cfg>::value) {
      constexpr int GPC = SFast<Cfg>::GPC;
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        r.sp[o] = env.s_entry(cta, o, t);
        r.sm[o] = env.s_entry(cta, o, GPC - t);
      }
      for_each_elem<Cfg>([&](auto ec) {
        using A = SAddr<Cfg, decltype(ec)::value>;
        if constexpr (A::special) {
          static_assert(A::sidx < 2, "two special elements at most");
          r.sx[A::sidx] = (t == 0) ? env.s_entry(cta, A::sp_owner, A::sp_off) : env.s_entry_add(r.sm[A::owner], A::delta);
        }
      });
    }
  });
  env.sync_cta();
}

// Issue the HBM loads of one P1 pass (pred / target rows r', r' + N/2 of every thread group) into
// ThreadRegs; nothing waits for them here.
// DIFF (compile time): pred already holds the difference map (fused DSL op), there is no target: the
// ta / tb registers (32 of the 64 that hold a pass in flight at N = 256) and the subtraction vanish.
template <class Cfg, bool DIFF = false, class Env>
FAVAE_HD void ffl_issue_loads(Env& env, const FflParams& p, long long batch, int pass) {
  constexpr int N = Cfg::N, TG = Cfg::TG, NG = Cfg::NG, HALF = Cfg::HALF, V4 = Cfg::IO_V4;
  constexpr int GPC = HALF / Cfg::C;
  env.for_threads([&](int cta, int tid) {
    ThreadRegs<Cfg>& r = env.regs(cta, tid);
    const int g = tid / TG, t = tid % TG, item = pass * NG + g;
    const int m = item / GPC, rp = cta * GPC + item % GPC;
    const long long map = batch * Cfg::MPC + m;
    const bool live = map < p.maps;
    const long long base = map * (long long)(N * N) + rp * N;
#pragma unroll
    for (int j = 0; j < V4; ++j) {
      const int f = t + TG * j;                          // float4 index inside the row
      r.pa[j] = r.pb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (!DIFF) r.ta[j] = r.tb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        r.pa[j] = *reinterpret_cast<const float4*>(p.pred + base + 4 * f);
        r.pb[j] = *reinterpret_cast<const float4*>(p.pred + base + HALF * N + 4 * f);
        if constexpr (!DIFF) {
          r.ta[j] = *reinterpret_cast<const float4*>(p.target + base + 4 * f);
          r.tb[j] = *reinterpret_cast<const float4*>(p.target + base + HALF * N + 4 * f);
        }
      }
    }
  });
}

// One batch = MPC maps (C == 1) or one map shared by the C CTAs of a cluster.  With
// Cfg::PIPELINE_LOADS, pass 0 of its loads must already be in flight (ffl_issue_loads(batch, 0)) and
// the batch issues pass 0 of next_batch under its last gradient stores.
// FAST: the reference's configuration (alpha == 1, no log weighting), where f(A) = A = sqrt(m)/N with
// m = |D|^2 unnormalised.  The spectrum statistics then run on m (sum m*sqrt(m), max m: the maximum
// commutes with the monotone f) and are converted once per thread, and the weight is
// min(sqrt(m) / sqrt(m_max), 1): 5-6 instructions per bin instead of ~12.
#ifndef FAVAE_FFL_DIFF_PIPE
#define FAVAE_FFL_DIFF_PIPE 1
#endif
// Direct I/O: every thread reads / writes its FFT-layout elements n = R2 e + t of the two rows straight
// from / to global memory with 4-byte accesses (the TG lanes of a transform cover 4 TG contiguous bytes
// per instruction) instead of float4 accesses redistributed through the staging area: the same
// instruction count, two shared-memory passes fewer on the way in and two on the way out (the kernel
// is shared-memory-bandwidth bound, profiles/ncu_r2_summary.md) -- but four times the global requests
// through the same L1 pipe.  Measured on B200, 4096 maps of 256^2: single-input form 1.11 -> 1.06 ms,
// two-input form 1.29 -> 1.35 ms (twice the loads and stores), loss only 0.72 -> 0.74 ms; 128^2
// two-input 0.74 -> 0.78 ms; 512^2 two-input 2.15 -> 2.05 ms.  Hence: the single-input form from
// 256^2 up, and 512^2 always.
#ifndef FAVAE_FFL_DIRECT_LOAD
#define FAVAE_FFL_DIRECT_LOAD 1
#endif
#ifndef FAVAE_FFL_DIRECT_STORE
#define FAVAE_FFL_DIRECT_STORE 1
#endif
template <class Cfg, bool DIFF> struct FflDirect {
  static constexpr bool value = Cfg::DIRECT_IO_OK && ((DIFF && Cfg::N >= 256) || Cfg::N >= 512);
  static constexpr bool load = value && FAVAE_FFL_DIRECT_LOAD, store = value && FAVAE_FFL_DIRECT_STORE;
};
// Bulk stores: the gradient rows of a pass are laid out as whole rows in the group's (idle) FFT staging
// area with 4-byte shared-memory stores and leave through ONE bulk async copy per row (cp.async.bulk,
// 4 N bytes) issued by the group's first lane, instead of 2 R1 four-byte global stores per thread whose
// issue alone took 12.6 % of the map time (the LSU queue backs up: lg_throttle; phase stamps of
// profiles/tools/ffl_phase_timing.cu).  Odd groups shift their rows by 16 floats so that the two groups of
// a warp write disjoint banks.  The staging area is handed back when the copy has READ it
// (wait_group.read), which is awaited just before the next exchange writes there.
#ifndef FAVAE_FFL_BULK_STORE
#define FAVAE_FFL_BULK_STORE 1
#endif
template <class Cfg, bool DIFF> struct FflBulk {
  static constexpr bool value = FAVAE_FFL_BULK_STORE && DIFF && FflDirect<Cfg, DIFF>::load && FflDirect<Cfg, DIFF>::store &&
                                (2 * Cfg::N + 16 <= 2 * Cfg::R1 * Cfg::STG_STRIDE);
};
template <class Cfg, bool DIFF> struct FflPipe {
  // the single-input form has the registers to keep the next map's first pass in flight under the last
  // gradient stores of the current one (the two-input form spills at 512 threads, see Cfg::PIPELINE_LOADS)
  static constexpr bool value =
      !FflDirect<Cfg, DIFF>::load && (Cfg::PIPELINE_LOADS || (DIFF && Cfg::C == 2 && FAVAE_FFL_DIFF_PIPE));
};
// Column group 0 (the packed real columns 0 and N/2) shares a warp with an ordinary group.  0: a branch with
// stage 1 duplicated on both sides (that warp runs stage 1 twice); 1: one stage 1 behind the branch (the
// compiler reconciles the two register layouts with ~55 moves on the common path); 2: branch-free, the
// general separation for every lane and the packed columns selected over it (32 selects per pass).
// Measured, 4096 maps of 256^2, single-input form: 0.967 / 0.974 / 0.958 ms.
#ifndef FAVAE_FFL_MERGE_STAGE1
#define FAVAE_FFL_MERGE_STAGE1 2
#endif
template <class Cfg, bool FAST = false, bool DIFF = false, class Env>
FAVAE_HD void ffl_map_batch(Env& env, const FflParams& p, long long batch, long long next_batch = -1) {
  constexpr bool PIPE = FflPipe<Cfg, DIFF>::value;
  constexpr bool DIRECT = FflDirect<Cfg, DIFF>::load, DIRECT_ST = FflDirect<Cfg, DIFF>::store;
  constexpr bool BULK = FflBulk<Cfg, DIFF>::value;
  constexpr int N = Cfg::N, R1 = Cfg::R1, TG = Cfg::TG, NG = Cfg::NG, HALF = Cfg::HALF;
  constexpr int C = Cfg::C, MPC = Cfg::MPC, T = Cfg::THREADS, PASSES = Cfg::PASSES;
  constexpr int GPC = HALF / C;                 // row pairs / column groups per CTA and map
  constexpr int TMAP = T / MPC;                 // threads per map slot
  constexpr int STG = R1 * Cfg::STG_STRIDE;
  constexpr int IS = Cfg::S_IDX;                // S stride of the row / frequency index
  static_assert(MPC == 1 || PASSES == 1, "several maps per CTA need a single pass");
  const float inv_nn = 1.0f / (float)(N * N);
  const float inv_n = 1.0f / (float)N;
  const long long map0 = batch * MPC;
  // f(A) of the packed columns (P3 / P4): the lean path has f = A = sqrt(A^2), no run-time alpha / log cases
  auto spec_f = [&](float a2) -> float {
    if constexpr (FAST) return favae_fast_sqrt(a2);
    else return spectrum_f(a2, p.alpha, p.log_matrix);
  };

  env.for_threads([&](int cta, int tid) {
    ThreadRegs<Cfg>& r = env.regs(cta, tid);
    r.sum = 0.0f; r.mx = 0.0f;
  });

  env.mark(-1);
  // ---------------- P1: packed row FFTs, global -> S ----------------
  // Inputs are read as float4 (each thread IO_V4 vectors per row and tensor) and redistributed to
  // the FFT's strided layout through the group's staging area.
  constexpr int V4 = Cfg::IO_V4;
  constexpr int IOB4 = Cfg::IO_B2 / 2;          // float4 index of the second half of the I/O staging
  for (int pass = 0; pass < PASSES; ++pass) {
    if constexpr (DIRECT) {
      // straight into the FFT layout: register e of lane t is element n = R2 e + t of the packed row pair
      env.for_threads([&](int cta, int tid) {
        ThreadRegs<Cfg>& r = env.regs(cta, tid);
        const int g = tid / TG, t = tid % TG, item = pass * NG + g;
        const int m = item / GPC, rp = cta * GPC + item % GPC;
        const long long map = batch * MPC + m;
        const bool live = map < p.maps;
        const long long base = map * (long long)(N * N) + rp * N;
#pragma unroll
        for (int e = 0; e < R1; ++e) {
          const int n = idx_in<Cfg>(t, e);
          float a = 0.f, b = 0.f;
          if (live) {
            a = p.pred[base + n];
            b = p.pred[base + HALF * N + n];
            if constexpr (!DIFF) { a -= p.target[base + n]; b -= p.target[base + HALF * N + n]; }
          }
          r.v[e] = make_float2(a, b);
        }
      });
    } else {
    if (pass > 0 || !PIPE) ffl_issue_loads<Cfg, DIFF>(env, p, batch, pass);
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG;
      float4* stg4 = reinterpret_cast<float4*>(env.stg(cta) + g * STG);
      const float4 (&pa)[V4] = r.pa, (&ta)[V4] = r.ta, (&pb)[V4] = r.pb, (&tb)[V4] = r.tb;
#pragma unroll
      for (int j = 0; j < V4; ++j) {
        const int f = t + TG * j;
        float4 a = pa[j], b = pb[j];
        if constexpr (!DIFF) {
          a = make_float4(pa[j].x - ta[j].x, pa[j].y - ta[j].y, pa[j].z - ta[j].z, pa[j].w - ta[j].w);
          b = make_float4(pb[j].x - tb[j].x, pb[j].y - tb[j].y, pb[j].z - tb[j].z, pb[j].w - tb[j].w);
        }
        if constexpr (Cfg::R2 > 1) {                     // interleave (row r', row r'+N/2) pairs
          stg4[f] = make_float4(a.x, b.x, a.y, b.y);          // elements 4f, 4f+1
          stg4[IOB4 + f] = make_float4(a.z, b.z, a.w, b.w);   // elements 4f+2, 4f+3 (bank-shifted half)
        } else {                                         // one thread owns the whole row pair
          r.v[4 * j + 0] = make_float2(a.x, b.x); r.v[4 * j + 1] = make_float2(a.y, b.y);
          r.v[4 * j + 2] = make_float2(a.z, b.z); r.v[4 * j + 3] = make_float2(a.w, b.w);
        }
      }
    });
    env.sync_warp();
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG;
      if constexpr (Cfg::R2 > 1) {
        const float2* stg = env.stg(cta) + g * STG;
#pragma unroll
        for (int e = 0; e < R1; ++e) r.v[e] = stg[io_slot<Cfg>(idx_in<Cfg>(t, e))];
      }
    });
    }
    if constexpr (BULK) {
      // the previous map's last gradient rows must have left the staging area
      if (pass == 0) env.for_threads([&](int, int tid) { if (tid % TG == 0) env.bulk_wait_read(); });
    }
    env.sync_warp();
    env.mark(0);
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG;
      fwd_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
    });
    env.sync_warp();
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, rp = cta * GPC + item % GPC;
      const unsigned int* tab = env.tab(cta);
      fwd_stage2<Cfg>(r, t, env.stg(cta) + g * STG);
      // the previous map's P6 reads of S (arrived at the end of the last batch) must be complete
      // before anybody overwrites S: the wait half of the split barrier sits here, after this
      // map's loads and row FFTs
      if (pass == 0) env.cluster_wait();
      const int at = IS * rp + (Cfg::S_ENTRY_MAJOR ? 0 : m * (GPC * 2 * Cfg::COLSTRIDE));
      if constexpr (SFast<Cfg>::value) {
        for_each_elem<Cfg>([&](auto ec) {
          constexpr int E = decltype(ec)::value;
          using A = SAddr<Cfg, E>;
          const unsigned int ent = A::special ? r.sx[A::sidx] : A::pos ? r.sp[A::owner] : r.sm[A::owner];
          env.template s_put_d<(A::special ? 0 : A::delta)>(cta, ent, at, r.v[E]);
        });
      } else {
#pragma unroll
        for (int e = 0; e < R1; ++e) env.s_put(cta, tab[idx_out<Cfg>(t, e)], at, r.v[e]);
      }
    });
    env.sync_warp();
    env.mark(1);
  }
  env.sync_cluster();
  env.mark(2);
  // pull the next batch's rows of this CTA into L2 while this one is transformed (one bulk
  // prefetch per contiguous run of rows; a hint only)
  auto prefetch_next = [&]() {
    if (!(next_batch >= 0 && next_batch * MPC < p.maps)) return;
    env.for_threads([&](int cta, int tid) {
      if (tid == 0) {
        const long long nmaps = (p.maps - next_batch * MPC < MPC) ? p.maps - next_batch * MPC : MPC;
        if constexpr (C == 1) {
          env.prefetch_l2(p.pred + next_batch * MPC * (long long)(N * N), nmaps * N * N * sizeof(float));
          if (!DIFF) env.prefetch_l2(p.target + next_batch * MPC * (long long)(N * N), nmaps * N * N * sizeof(float));
        } else {
          const long long base = next_batch * (long long)(N * N) + (long long)cta * GPC * N;
          env.prefetch_l2(p.pred + base, GPC * N * sizeof(float));
          env.prefetch_l2(p.pred + base + HALF * N, GPC * N * sizeof(float));
          if (!DIFF) {
            env.prefetch_l2(p.target + base, GPC * N * sizeof(float));
            env.prefetch_l2(p.target + base + HALF * N, GPC * N * sizeof(float));
          }
        }
      }
    });
  };
  // Where the hint is issued matters: right here (a whole map time ahead of its use) 8 % of the
  // prefetched lines are evicted again by the gradient stores streaming through L2 and the kernel
  // is 5 % slower than with the hint at the start of P5 (measured, 4096 maps of 256^2).  Loss-only
  // calls end after P3, so they prefetch here.
#ifndef FAVAE_FFL_PF_POINT
#define FAVAE_FFL_PF_POINT 5
#endif
  const bool want_grad = p.grad_pred != nullptr || p.grad_target != nullptr;
  if (FAVAE_FFL_PF_POINT == 2 || !want_grad) prefetch_next();

  // ---------------- P2: column FFTs + spectrum statistics ----------------
  for (int pass = 0; pass < PASSES; ++pass) {
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, v = cta * GPC + item % GPC;
      int off0, off1;
      s_group_offsets<Cfg>(item % GPC, m, off0, off1);
      const float2* S = env.S(cta, cta);
#if FAVAE_FFL_MERGE_STAGE1 == 2
      // branch-free: the general separation for every lane, the packed real columns of group 0 selected over it
      {
        const bool g0 = v == 0;
#pragma unroll
        for (int e = 0; e < R1 / 2; ++e) {
          const int rr = idx_in<Cfg>(t, e);
          const float2 zv = S[off0 + IS * rr], zw = S[off1 + IS * rr];
          const float2 a = pk_fma(zw, make_float2(0.5f, -0.5f), pk_mul(zv, make_float2(0.5f, 0.5f)));
          const float2 b = pk_fma(pk_swap(zv), make_float2(0.5f, -0.5f), pk_mul(pk_swap(zw), make_float2(0.5f, 0.5f)));
          r.v[e] = g0 ? make_float2(zv.x, zw.x) : a;
          r.v[e + R1 / 2] = g0 ? make_float2(zv.y, zw.y) : b;
        }
        fwd_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
      }
      if (false)
#endif
      if (v == 0) {                                      // packed real columns 0 and N/2
#pragma unroll
        for (int e = 0; e < R1 / 2; ++e) {
          const int rr = idx_in<Cfg>(t, e);
          const float2 zv = S[off0 + IS * rr], zw = S[off1 + IS * rr];
          r.v[e] = make_float2(zv.x, zw.x);
          r.v[e + R1 / 2] = make_float2(zv.y, zw.y);
        }
#if FAVAE_FFL_MERGE_STAGE1 == 0
        // (stage 1 sits inside both branches: they merge after the values have gone to staging,
        // not through two dozen register copies)
        fwd_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
#endif
      } else {                                           // separate the two packed real rows
#pragma unroll
        for (int e = 0; e < R1 / 2; ++e) {
          const int rr = idx_in<Cfg>(t, e);
          const float2 zv = S[off0 + IS * rr], zw = S[off1 + IS * rr];
          // 0.5 (zv + conj zw) and 0.5 (zv - conj zw) / i, as packed operations
          r.v[e] = pk_fma(zw, make_float2(0.5f, -0.5f), pk_mul(zv, make_float2(0.5f, 0.5f)));
          r.v[e + R1 / 2] = pk_fma(pk_swap(zv), make_float2(0.5f, -0.5f), pk_mul(pk_swap(zw), make_float2(0.5f, 0.5f)));
        }
#if FAVAE_FFL_MERGE_STAGE1 == 0
        fwd_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
#endif
      }
#if FAVAE_FFL_MERGE_STAGE1 == 1
      // one copy of stage 1 behind the branch: the warp that holds column group 0 next to an ordinary group
      // would otherwise run the ~200 instructions of stage 1 twice, once per side of the branch, and every
      // other warp of the cluster waits for it at the statistics barrier
      fwd_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
#endif
    });
    env.sync_warp();
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, v = cta * GPC + item % GPC;
      int off0, off1;
      s_group_offsets<Cfg>(item % GPC, m, off0, off1);
      float2* S = env.S(cta, cta);
      fwd_stage2<Cfg>(r, t, env.stg(cta) + g * STG);
      // out-layout register e holds u = idx_out(t, e); u and u + N/2 sit PO_OUT registers apart
      if (v != 0) {
        float sum = 0.f, mx = r.mx;
#pragma unroll
        for (int e = 0; e < R1; ++e) {
          const float2 z = r.v[e];
          if constexpr (FAST) {                          // in units of m; converted after the last pass
            const float mm = fmaf(z.y, z.y, z.x * z.x);
            sum = fmaf(mm, favae_fast_sqrt(mm), sum);
            mx = fmaxf(mx, mm);
          } else {
            const float a2 = (z.x * z.x + z.y * z.y) * inv_nn;
            const float f = spectrum_f(a2, p.alpha, p.log_matrix);
            sum = fmaf(f, a2, sum);
            mx = fmaxf(mx, f);
          }
        }
        r.sum = fmaf(2.0f, sum, r.sum);
        r.mx = mx;
      }
      // The column spectra of the LAST pass stay in the payload registers across the statistics
      // exchange when a gradient follows: P5 starts with that pass and weights them where they are, which
      // saves one write and one read of S for those columns (all of them for the single-pass
      // configurations; half at N >= 256) -- the kernel is shared-memory-bandwidth bound.  Group 0 (the
      // packed real columns 0 and N/2) always goes through S: P3 / P4 work on it there.
      // (Loss-only calls end after P3, which reads group 0 alone: nothing else is written at all.)
      if (!(v != 0 && (!want_grad || pass == PASSES - 1))) {
#pragma unroll
        for (int e = 0; e < R1; ++e) {
          const int u = idx_out<Cfg>(t, e);
          S[(u < HALF) ? off0 + IS * u : off1 + IS * (u - HALF)] = r.v[e];
        }
      }
    });
    env.sync_warp();
  }
  env.mark(3);
  env.sync_cta();

  // ---------------- P3: statistics of the packed columns v = 0 and v = N/2 ----------------
  env.for_threads([&](int cta, int tid) {
    ThreadRegs<Cfg>& r = env.regs(cta, tid);
    if constexpr (FAST) {                                // m units -> (sum f A^2, max f)
      r.sum *= inv_nn * inv_n;
      r.mx = favae_fast_sqrt(r.mx) * inv_n;
    }
    if (cta == 0) {
      const int m = tid / TMAP;
      int o, off0, off1;
      s_locate<Cfg>(0, m, o, off0);
      s_locate<Cfg>(HALF, m, o, off1);
      const float2* S = env.S(cta, cta);
      for (int u = tid % TMAP; u <= HALF; u += TMAP) {
        const int un = (N - u) % N;
        const float2 a = S[(u < HALF) ? off0 + IS * u : off1 + IS * (u - HALF)];
        const float2 b = S[(un < HALF) ? off0 + IS * un : off1 + IS * (un - HALF)];
        const float2 d0 = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
        const float2 dn = make_float2(0.5f * (a.y + b.y), 0.5f * (b.x - a.x));
        const float mult = (u == 0 || u == HALF) ? 1.0f : 2.0f;
        const float a0 = (d0.x * d0.x + d0.y * d0.y) * inv_nn;
        const float an = (dn.x * dn.x + dn.y * dn.y) * inv_nn;
        const float f0 = spec_f(a0);
        const float fn = spec_f(an);
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
  env.sync_warp();                               // the L partial sums of a map slot sit in one warp
  env.for_threads([&](int cta, int tid) {
    float* fb = env.fbuf(cta);
    const int m = tid / TMAP, j = tid % TMAP;
    constexpr int L = TMAP < 32 ? TMAP : 32;
    if (j == 0) {
      float s = 0.f, mx = 0.f;
      for (int i = 0; i < L; ++i) { s += fb[2 * T + m * TMAP + i]; mx = fmaxf(mx, fb[3 * T + m * TMAP + i]); }
      if (C == 1) {
        fb[4 * T + 2 * m] = s;
        const float used = p.fmax_override ? p.fmax_override[0] : mx;
        fb[4 * T + 2 * m + 1] = used;
        fb[4 * T + 2 * MPC + m] = mx;
        fb[4 * T + 3 * MPC + m] = spectrum_inv(used);
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
  // per-map results: (sum f A^2, max used for the weights, max of this map, 1 / used max).  Clusters
  // combine the per-CTA slots on the fly (a few broadcast loads) instead of a third barrier.
  auto map_stats = [&](int cta, int m, float& s, float& used, float& mx, float& finv) {
    if constexpr (C == 1) {
      const float* fb = env.fbuf(cta);
      s = fb[4 * T + 2 * m]; used = fb[4 * T + 2 * m + 1]; mx = fb[4 * T + 2 * MPC + m]; finv = fb[4 * T + 3 * MPC + m];
    } else {
      const float* cl = env.cl(cta, cta);
      s = 0.f; mx = 0.f;
      for (int o = 0; o < C; ++o) { s += cl[2 * o]; mx = fmaxf(mx, cl[2 * o + 1]); }
      used = p.fmax_override ? p.fmax_override[0] : mx;
      finv = spectrum_inv(used);
    }
  };
  env.for_threads([&](int cta, int tid) {
    const int m = tid / TMAP, j = tid % TMAP;
    if (cta == 0 && j == 0 && map0 + m < p.maps) {
      float s, used, mx, finv;
      map_stats(cta, m, s, used, mx, finv);
      p.map_loss[map0 + m] = (used > 0.0f) ? s / used : 0.0f;
      if (p.map_max) p.map_max[map0 + m] = mx;
    }
  });
  if (p.grad_pred == nullptr && p.grad_target == nullptr) {
    env.for_threads([&](int, int) { env.cluster_arrive_relaxed(); });     // S is not read again
    if (PIPE && next_batch >= 0) ffl_issue_loads<Cfg, DIFF>(env, p, next_batch, 0);
    return;
  }

  env.mark(4);
  // ---------------- P4: weight the packed columns in place ----------------
  env.for_threads([&](int cta, int tid) {
    if (cta != 0) return;
    const int m = tid / TMAP;
    float s_, used_, mx_, finv;
    map_stats(cta, m, s_, used_, mx_, finv);
    int o, off0, off1;
    s_locate<Cfg>(0, m, o, off0);
    s_locate<Cfg>(HALF, m, o, off1);
    float2* S = env.S(cta, cta);
    for (int u = tid % TMAP; u <= HALF; u += TMAP) {
      const int un = (N - u) % N;
      const int ia = (u < HALF) ? off0 + IS * u : off1 + IS * (u - HALF);
      const int ib = (un < HALF) ? off0 + IS * un : off1 + IS * (un - HALF);
      const float2 a = S[ia], b = S[ib];
      const float2 d0 = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
      const float2 dn = make_float2(0.5f * (a.y + b.y), 0.5f * (b.x - a.x));
      const float a0 = (d0.x * d0.x + d0.y * d0.y) * inv_nn;
      const float an = (dn.x * dn.x + dn.y * dn.y) * inv_nn;
      const float w0 = spectrum_w(spec_f(a0), finv) * p.grad_scale;
      const float wn = spectrum_w(spec_f(an), finv) * p.grad_scale;
      S[ia] = make_float2(w0 * d0.x - wn * dn.y, w0 * d0.y + wn * dn.x);
      if (ib != ia) S[ib] = make_float2(w0 * d0.x + wn * dn.y, wn * dn.x - w0 * d0.y);
    }
  });
  env.sync_cta();

  env.mark(5);
  if (FAVAE_FFL_PF_POINT == 5) prefetch_next();
  // ---------------- P5: weight + inverse column FFTs, re-pack into Z' ----------------
  for (int pi = 0; pi < PASSES; ++pi) {
    const int pass = (pi == 0) ? PASSES - 1 : pi - 1;   // the pass P2 left in the registers comes first
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, v = cta * GPC + item % GPC;
      float s_, used_, mx_, finv;
      map_stats(cta, m, s_, used_, mx_, finv);
      int off0, off1;
      s_group_offsets<Cfg>(item % GPC, m, off0, off1);
      const float2* S = env.S(cta, cta);
      if (!(pi == 0 && v != 0)) {
#pragma unroll
        for (int e = 0; e < R1; ++e) {
          const int u = idx_out<Cfg>(t, e);
          r.v[e] = S[(u < HALF) ? off0 + IS * u : off1 + IS * (u - HALF)];
        }
      }
      if (v != 0) {                                      // group 0 was weighted in place by P4
        // the gradient scale rides on the weights (FAST: grad_scale >= 0, so it commutes with the clamp)
        const float gs = p.grad_scale, finv_g = finv * inv_n * gs;
#pragma unroll
        for (int e = 0; e < R1; ++e) {
          const float2 z = r.v[e];
          float w;
          if constexpr (FAST) w = fminf(favae_fast_sqrt(fmaf(z.y, z.y, z.x * z.x)) * finv_g, gs);
          else w = spectrum_w(spectrum_f((z.x * z.x + z.y * z.y) * inv_nn, p.alpha, p.log_matrix), finv) * gs;
          r.v[e] = pk_mul(z, pk_dup(w));
        }
      }
      inv_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
    });
    env.sync_warp();
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, v = cta * GPC + item % GPC;
      int off0, off1;
      s_group_offsets<Cfg>(item % GPC, m, off0, off1);
      float2* S = env.S(cta, cta);
      inv_stage2<Cfg>(r, t, env.stg(cta) + g * STG);
      if (v == 0) {
#pragma unroll
        for (int e = 0; e < R1 / 2; ++e) {
          const int rr = idx_in<Cfg>(t, e);
          const float2 a = r.v[e], b = r.v[e + R1 / 2];
          S[off0 + IS * rr] = make_float2(a.x, b.x);
          S[off1 + IS * rr] = make_float2(a.y, b.y);
        }
      } else {
#pragma unroll
        for (int e = 0; e < R1 / 2; ++e) {
          const int rr = idx_in<Cfg>(t, e);
          const float2 a = r.v[e], b = r.v[e + R1 / 2];
          S[off0 + IS * rr] = pk_fma(pk_swap(b), make_float2(-1.0f, 1.0f), a);    // a + i b
          S[off1 + IS * rr] = pk_fma(a, make_float2(1.0f, -1.0f), pk_swap(b));    // conj(a) + i conj(b)
        }
      }
    });
    env.sync_warp();
  }
  env.mark(6);
  env.sync_cluster();
  env.mark(7);
  if (FAVAE_FFL_PF_POINT == 6) prefetch_next();

  // ---------------- P6: inverse row FFTs, S -> gradient rows ----------------
  // The gather of pass p + 1 (half of it distributed-shared-memory loads from the peer) is issued
  // before the gradient stores of pass p, when the payload registers are free again.
  auto gather = [&](int pass) {
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, rp = cta * GPC + item % GPC;
      const unsigned int* tab = env.tab(cta);
      const int at = IS * rp + (Cfg::S_ENTRY_MAJOR ? 0 : m * (GPC * 2 * Cfg::COLSTRIDE));
      if constexpr (SFast<Cfg>::value) {
        for_each_elem<Cfg>([&](auto ec) {
          constexpr int E = decltype(ec)::value;
          using A = SAddr<Cfg, E>;
          const unsigned int ent = A::special ? r.sx[A::sidx] : A::pos ? r.sp[A::owner] : r.sm[A::owner];
          r.v[E] = env.template s_get_d<(A::special ? 0 : A::delta)>(cta, ent, at);
        });
      } else {
#pragma unroll
        for (int e = 0; e < R1; ++e) r.v[e] = env.s_get(cta, tab[idx_out<Cfg>(t, e)], at);
      }
    });
  };
  gather(0);
  for (int pass = 0; pass < PASSES; ++pass) {
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG;
      env.mark(8);
      inv_stage1<Cfg>(r, t, env.stg(cta) + g * STG);
      // last read of S: release it for the next map.  inv_stage1 has consumed every loaded value, so
      // the reads have completed and no fence is needed (a releasing arrive would also wait for the
      // previous pass's gradient stores to drain)
      if (pass == PASSES - 1) env.cluster_arrive_relaxed();
    });
    env.sync_warp();
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG;
      inv_stage2<Cfg>(r, t, env.stg(cta) + g * STG);
    });
    env.sync_warp();
    if constexpr (BULK) {
      env.mark(9);
      env.for_threads([&](int cta, int tid) {
        ThreadRegs<Cfg>& r = env.regs(cta, tid);
        const int g = tid / TG, t = tid % TG;
        float* row = reinterpret_cast<float*>(env.stg(cta) + g * STG) + ((g & 1) ? 16 : 0);
#pragma unroll
        for (int e = 0; e < R1; ++e) {
          const int n = idx_in<Cfg>(t, e);
          row[n] = r.v[e].x;
          row[N + n] = r.v[e].y;
        }
        env.fence_async();                       // generic-proxy writes -> visible to the bulk copy
      });
      env.sync_warp();
      env.for_threads([&](int cta, int tid) {
        const int g = tid / TG, t = tid % TG, item = pass * NG + g;
        const int m = item / GPC, rp = cta * GPC + item % GPC;
        const long long map = map0 + m;
        if (t == 0 && map < p.maps) {
          const float* row = reinterpret_cast<const float*>(env.stg(cta) + g * STG) + ((g & 1) ? 16 : 0);
          const long long base = map * (long long)(N * N) + rp * N;
          env.bulk_store(p.grad_pred + base, row, N * sizeof(float));
          env.bulk_store(p.grad_pred + base + HALF * N, row + N, N * sizeof(float));
          env.bulk_commit();
        }
      });
      if (pass + 1 < PASSES) {
        gather(pass + 1);
        env.for_threads([&](int, int tid) { if (tid % TG == 0) env.bulk_wait_read(); });
      }
    } else if constexpr (DIRECT_ST) {
      // straight from the FFT layout (element n = R2 e + t of the two rows), scaled already: 4-byte
      // stores, 4 TG contiguous bytes per instruction and transform
      env.mark(9);
      env.for_threads([&](int cta, int tid) {
        ThreadRegs<Cfg>& r = env.regs(cta, tid);
        const int g = tid / TG, t = tid % TG, item = pass * NG + g;
        const int m = item / GPC, rp = cta * GPC + item % GPC;
        const long long map = map0 + m;
        if (map < p.maps) {
          const long long base = map * (long long)(N * N) + rp * N;
#pragma unroll
          for (int e = 0; e < R1; ++e) {
            const int n = idx_in<Cfg>(t, e);
            if (p.grad_pred) {
              p.grad_pred[base + n] = r.v[e].x;
              p.grad_pred[base + HALF * N + n] = r.v[e].y;
            }
            if (p.grad_target) {
              p.grad_target[base + n] = -r.v[e].x;
              p.grad_target[base + HALF * N + n] = -r.v[e].y;
            }
          }
        }
      });
      if (pass + 1 < PASSES) gather(pass + 1);
    } else {
    // back to float4 rows through the staging area, scaled, stored as +grad / -grad
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG;
      if constexpr (Cfg::R2 > 1) {
        float2* stg = env.stg(cta) + g * STG;
#pragma unroll
        for (int e = 0; e < R1; ++e) stg[io_slot<Cfg>(idx_in<Cfg>(t, e))] = r.v[e];
      }
    });
    env.sync_warp();
    env.mark(9);
    if (pass + 1 < PASSES) gather(pass + 1);
    // the next batch's first rows start their trip from HBM / L2 under this pass's gradient stores
    if (PIPE && pass == PASSES - 1 && next_batch >= 0) ffl_issue_loads<Cfg, DIFF>(env, p, next_batch, 0);
    env.for_threads([&](int cta, int tid) {
      ThreadRegs<Cfg>& r = env.regs(cta, tid);
      const int g = tid / TG, t = tid % TG, item = pass * NG + g;
      const int m = item / GPC, rp = cta * GPC + item % GPC;
      const long long map = map0 + m;
      if (map < p.maps) {
        const long long base = map * (long long)(N * N) + rp * N;
        const float4* stg4 = reinterpret_cast<const float4*>(env.stg(cta) + g * STG);
#pragma unroll
        for (int j = 0; j < V4; ++j) {
          const int f = t + TG * j;
          float4 a, b;                                   // already scaled: grad_scale rode on the weights
          if constexpr (Cfg::R2 > 1) {
            const float4 lo = stg4[f], hi = stg4[IOB4 + f];
            a = make_float4(lo.x, lo.z, hi.x, hi.z);
            b = make_float4(lo.y, lo.w, hi.y, hi.w);
          } else {
            a = make_float4(r.v[4 * j].x, r.v[4 * j + 1].x, r.v[4 * j + 2].x, r.v[4 * j + 3].x);
            b = make_float4(r.v[4 * j].y, r.v[4 * j + 1].y, r.v[4 * j + 2].y, r.v[4 * j + 3].y);
          }
          if (p.grad_pred) {
            *reinterpret_cast<float4*>(p.grad_pred + base + 4 * f) = a;
            *reinterpret_cast<float4*>(p.grad_pred + base + HALF * N + 4 * f) = b;
          }
          if (p.grad_target) {
            *reinterpret_cast<float4*>(p.grad_target + base + 4 * f) = make_float4(-a.x, -a.y, -a.z, -a.w);
            *reinterpret_cast<float4*>(p.grad_target + base + HALF * N + 4 * f) = make_float4(-b.x, -b.y, -b.z, -b.w);
          }
        }
      }
    });
    }
    env.sync_warp();
    env.mark(10);
  }
}

}  // namespace favae
