// Fused spectrum-loss core: one real 2-D FFT of (pred - target), dynamic spectrum
// weight, weighted energy, and the weighted inverse transform (= the gradient).
//
// Replaces focal_frequency_loss.FocalFrequencyLoss.forward / tensor2freq /
// loss_formulation (pip focal-frequency-loss==0.3.0, called from
// /root/reference/losses/vqgan_losses.py:14,25-26,45-46) and its autograd backward.
//
// The same source is compiled by nvcc for sm_100a and by g++ for the host emulation
// used in tests/test_ffl_emulation.py (thread loops instead of threads), so every
// cross-thread exchange happens at an explicit FAVAE_SYNC point.
//
// Math.  d = pred - target (real, N x N, N a power of two).
//   * rows r' and r'+N/2 are packed as one complex row z = d[r'] + i d[r'+N/2]
//   * row FFT (length N) of the N/2 packed rows -> Z[r'][v]  (stored S[v][r'])
//   * column group v in [1, N/2): separates Dr[r][v] from Z[.][v], Z[.][N-v] on load,
//     FFT over rows -> T[v][u] = unnormalised spectrum D[u][v]
//     column group 0 packs the two real columns v=0 and v=N/2 as one complex column
//   * stats: f(A) with A = |D|/N; sum f(A) A^2 (mirror half counted via weight 2), max f(A)
//   * weight, inverse column FFT, re-pack, inverse row FFT -> grad rows r', r'+N/2
//
// 1-D FFT of length N = R1*R2 on TG = R2 threads, R1 values per thread:
//   forward : x[R2*e + t] --FFT_R1--> *W_N^(t*k1) --exchange--> FFT_R2 --> X[t + R2*m + R1*k2]
//             (register e = m*R2 + k2)
//   inverse : the same pipeline run backwards, so the inverse consumes the forward's
//             output layout and produces its input layout: every phase reads and writes
//             exactly the addresses it touched on the way in.
#pragma once

#include <math.h>
#include <stdint.h>

#include <type_traits>

#if defined(__CUDACC__)
#define FAVAE_HD __host__ __device__ __forceinline__
#else
#define FAVAE_HD inline
#ifndef FAVAE_HOST_FLOAT2
#define FAVAE_HOST_FLOAT2
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif
#endif

#if defined(__CUDA_ARCH__)
#define FAVAE_RSQRT(x) rsqrtf(x)
// single-instruction square root (MUFU.SQRT, relative error <= 2^-23, sqrt(0) = 0)
__device__ __forceinline__ float favae_fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#else
#define FAVAE_RSQRT(x) (1.0f / sqrtf(x))
static inline float favae_fast_sqrt(float x) { return sqrtf(x); }
#endif

namespace favae {

// ----------------------------------------------------------------------------------
// complex helpers
// ----------------------------------------------------------------------------------
// Packed fp32x2 arithmetic: sm_100a executes add/mul/fma on a register pair in one issue slot
// (FADD2 / FMUL2 / FFMA2, same lane throughput as the scalar forms - profiles/tools/ffma2_bench.cu),
// which halves the instruction count of the complex butterflies.  The host build (emulation) uses
// the scalar forms; results are bit-identical (each lane is an IEEE fp32 op either way).
#if defined(__CUDA_ARCH__) && !defined(FAVAE_FFL_NO_PACKED)
#define FAVAE_PK_ASM(op)                                                                              \
  float2 c;                                                                                           \
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; " op " rc, ra, rb; mov.b64 {%0,%1}, rc;}" \
      : "=f"(c.x), "=f"(c.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                              \
  return c;
FAVAE_HD float2 pk_add(float2 a, float2 b) { FAVAE_PK_ASM("add.f32x2") }
FAVAE_HD float2 pk_sub(float2 a, float2 b) { FAVAE_PK_ASM("sub.f32x2") }
FAVAE_HD float2 pk_mul(float2 a, float2 b) { FAVAE_PK_ASM("mul.f32x2") }
#undef FAVAE_PK_ASM
FAVAE_HD float2 pk_fma(float2 a, float2 b, float2 d) {   // a * b + d, fused per lane
  float2 c;
  asm("{.reg .b64 ra, rb, rd, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rd, {%6,%7}; "
      "fma.rn.f32x2 rc, ra, rb, rd; mov.b64 {%0,%1}, rc;}"
      : "=f"(c.x), "=f"(c.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(d.x), "f"(d.y));
  return c;
}
#else
FAVAE_HD float2 pk_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
FAVAE_HD float2 pk_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
FAVAE_HD float2 pk_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
FAVAE_HD float2 pk_fma(float2 a, float2 b, float2 d) { return make_float2(fmaf(a.x, b.x, d.x), fmaf(a.y, b.y, d.y)); }
#endif
FAVAE_HD float2 pk_swap(float2 a) { return make_float2(a.y, a.x); }
FAVAE_HD float2 pk_dup(float v) { return make_float2(v, v); }

FAVAE_HD float2 cadd(float2 a, float2 b) { return pk_add(a, b); }
FAVAE_HD float2 csub(float2 a, float2 b) { return pk_sub(a, b); }
FAVAE_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by -i (DIR=-1, forward) or +i (DIR=+1, inverse)
template <int DIR> FAVAE_HD float2 crot(float2 a) {
  return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

// twiddle e^{DIR * 2*pi*i * j/32}, j = 0..15, as compile-time constants:
// a * (c + i s) = a (.) (c, c) + swap(a) (.) (-s, s), two packed operations
template <int J, int DIR> FAVAE_HD float2 ctw32(float2 a) {
  // cos(pi j / 16), j = 0..8
  constexpr float C[9] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                          0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f,
                          0.19509032201612825f, 0.0f};
  if constexpr (J == 0) return a;
  else if constexpr (J == 8) return crot<DIR>(a);
  else {
    constexpr float c = (J < 8) ? C[J] : -C[16 - J];
    constexpr float s0 = (J < 8) ? C[8 - J] : C[J - 8];
    constexpr float s = DIR < 0 ? -s0 : s0;
    // (the scalar form -- two FMUL + two FFMA -- measured 0.6 % slower: 4352 against 4200 instructions)
    return pk_fma(pk_swap(a), make_float2(-s, s), pk_mul(a, make_float2(c, c)));
  }
}

// ----------------------------------------------------------------------------------
// in-register FFT of R in {1,2,4,8,16,32} values, natural order in and out
// ----------------------------------------------------------------------------------
template <int R, int DIR> struct RegFFT;

template <int DIR> struct RegFFT<1, DIR> {
  static FAVAE_HD void run(float2 (&)[1]) {}
};
template <int DIR> struct RegFFT<2, DIR> {
  static FAVAE_HD void run(float2 (&v)[2]) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
  }
};

template <int R, int DIR, int K> struct Combine {
  static FAVAE_HD void run(float2 (&v)[R], const float2 (&e)[R / 2], const float2 (&o)[R / 2]) {
    float2 t = ctw32<K * (32 / R), DIR>(o[K]);
    v[K] = cadd(e[K], t);
    v[K + R / 2] = csub(e[K], t);
    if constexpr (K + 1 < R / 2) Combine<R, DIR, K + 1>::run(v, e, o);
  }
};

template <int R, int DIR> struct RegFFT {
  static FAVAE_HD void run(float2 (&v)[R]) {
    float2 e[R / 2], o[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; ++k) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
    RegFFT<R / 2, DIR>::run(e);
    RegFFT<R / 2, DIR>::run(o);
    Combine<R, DIR, 0>::run(v, e, o);
  }
};

// ----------------------------------------------------------------------------------
// geometry
// ----------------------------------------------------------------------------------
template <int N> struct FftGeom;
template <> struct FftGeom<8>   { static constexpr int R1 = 8,  R2 = 1; };
template <> struct FftGeom<16>  { static constexpr int R1 = 16, R2 = 1; };
template <> struct FftGeom<32>  { static constexpr int R1 = 8,  R2 = 4; };
template <> struct FftGeom<64>  { static constexpr int R1 = 8,  R2 = 8; };
template <> struct FftGeom<128> { static constexpr int R1 = 16, R2 = 8; };
template <> struct FftGeom<256> { static constexpr int R1 = 16, R2 = 16; };
template <> struct FftGeom<512> { static constexpr int R1 = 32, R2 = 16; };

// N: map side; C: CTAs per cluster sharing one map; MPC: maps per CTA (C==1 only)
template <int N_, int C_, int MPC_, int THREADS_> struct FflCfg {
  static constexpr int N = N_, C = C_, MPC = MPC_, THREADS = THREADS_;
  static constexpr int R1 = FftGeom<N>::R1, R2 = FftGeom<N>::R2;
  static constexpr int TG = R2;                    // threads per 1-D FFT
  static constexpr int NG = THREADS / TG;          // 1-D FFTs in flight per CTA
  static constexpr int HALF = N / 2;
  static constexpr int ITEMS = MPC * HALF / C;     // row pairs (= column groups) per CTA
  static constexpr int PASSES = ITEMS / NG;
  static constexpr int COLSTRIDE = HALF + 1;       // float2 per S column (padded)
  // S holds, per CTA, N/C map columns x HALF entries (row pair r' in P1/P6, frequency u in P2..P5).
  // One CTA per map: column-major (entry index contiguous).  Clusters: entry-major, S[entry][column]
  // with an odd row stride, so that the 16 lanes of a row FFT write / read 16 adjacent columns of
  // the peer CTA as one 128-byte distributed-shared-memory segment instead of 16 scattered 8-byte
  // words, while the column FFTs (stride COLSTRIDE) stay bank-conflict free.
  static constexpr bool S_ENTRY_MAJOR = C > 1;
  static constexpr int S_IDX = S_ENTRY_MAJOR ? (N / C + 1) : 1;
  static constexpr int S_FLOAT2 = S_ENTRY_MAJOR ? HALF * (N / C + 1) : MPC * (N / C) * COLSTRIDE;
  static constexpr int STG_STRIDE = R2 + 1;
  static constexpr int STG_FLOAT2 = (R2 > 1) ? NG * R1 * STG_STRIDE : 1;
  static constexpr int PO_IN = R1 / 2;                         // register offset of n + N/2
  static constexpr int PO_OUT = (R2 >= 2) ? R2 / 2 : R1 / 2;   // register offset of k + N/2
  static constexpr int TMAP = HALF * TG / C;       // threads of one CTA working on one map
  static_assert(ITEMS % NG == 0 && PASSES >= 1, "items must tile the thread groups");
  static_assert(C == 1 || MPC == 1, "clusters hold a single map");
  static_assert(32 % TG == 0, "a 1-D FFT group must sit inside a warp");
  static constexpr size_t SMEM_BYTES =
      sizeof(float2) * (size_t)(S_FLOAT2 + STG_FLOAT2) + sizeof(float) * (size_t)(4 * THREADS + 8 * MPC + 8 * C) + sizeof(unsigned int) * (size_t)N;
  static constexpr int IO_V4 = N / (4 * TG);       // float4 per thread and input row
  // Issue the next batch's first loads under the current batch's last stores.  Measured on B200 for
  // N = 256: at 512 threads the 64 extra live registers spill inside the row FFT (1.39 -> 1.62 ms);
  // at 256 threads / 220 registers it fits but gains 1 % (1.47 -> 1.46 ms), so it stays off.
  static constexpr bool PIPELINE_LOADS = false;
  // I/O staging of a row pair (N float2): elements with (c % 4) < 2 in [0, N/2), the others from
  // IO_B2 on; the 8-slot shift keeps the strided float2 reads of the two halves on different banks
  static constexpr int IO_B2 = N / 2 + 8;
  // Direct I/O (see FflDirect in ffl_driver.cuh) is possible when the TG lanes of a transform cover whole
  // 32-byte sectors with 4-byte accesses
#ifndef FAVAE_FFL_DIRECT_IO
#define FAVAE_FFL_DIRECT_IO 1
#endif
  static constexpr bool DIRECT_IO_OK = (R2 >= 8) && FAVAE_FFL_DIRECT_IO;
  static_assert(R2 == 1 || N + 8 <= R1 * STG_STRIDE, "I/O staging must fit the FFT staging area");
};

// per-thread registers that live across FAVAE_SYNC points
template <class Cfg> struct ThreadRegs {
  float2 v[Cfg::R1];        // FFT payload
  float2 tw[Cfg::R1];       // W_N^(t*k1), forward sign
  // the row pair in flight from HBM: pred / target, rows r' and r' + N/2.  Pass 0 of the next
  // batch is issued before the last gradient stores of the current one
  float4 pa[Cfg::IO_V4], ta[Cfg::IO_V4], pb[Cfg::IO_V4], tb[Cfg::IO_V4];
  float sum, mx;            // running stats
  // S addressing of the row-FFT scatter / gather without the per-element table (SFast below): entries of
  // (owner o, column slot t) and (owner o, slot GPC - t), plus the final entries of the special elements
  unsigned int sp[2], sm[2], sx[2];
};

struct FflParams {
  const float* pred;        // maps * N * N
  const float* target;
  float* grad_pred;         // nullable
  float* grad_target;       // nullable (written as -grad)
  float* map_loss;          // maps: sum_{u,v} w A^2 per map (unscaled by loss_weight / numel)
  float* map_max;           // nullable: max_{u,v} f(A) of each map
  const float* fmax_override;  // nullable: one device scalar used instead of the per-map max
                               // (FocalFrequencyLoss batch_matrix=True)
  long long maps;
  float grad_scale;         // 2*loss_weight/numel / (N*N)
  float alpha;
  int log_matrix;
};

// float2 slot of row-pair element c in the I/O staging layout
template <class Cfg> FAVAE_HD int io_slot(int c) {
  return ((c & 2) ? Cfg::IO_B2 : 0) + 2 * (c >> 2) + (c & 1);
}

// out-layout / in-layout index of register e of lane t
template <class Cfg> FAVAE_HD int idx_in(int t, int e) { return Cfg::R2 * e + t; }
template <class Cfg> FAVAE_HD int idx_out(int t, int e) {
  return t + Cfg::R2 * (e / Cfg::R2) + Cfg::R1 * (e % Cfg::R2);
}

// S addressing: column w in [0,N) of the map -> (owner CTA, float2 offset of row 0)
template <class Cfg> FAVAE_HD void s_locate(int w, int slot_map, int& owner, int& off) {
  constexpr int N = Cfg::N, HALF = Cfg::HALF, GPC = HALF / Cfg::C;   // groups per CTA
  int group, sub;
  if (w == 0) { group = 0; sub = 0; }
  else if (w == HALF) { group = 0; sub = 1; }
  else if (w < HALF) { group = w; sub = 0; }
  else { group = N - w; sub = 1; }
  owner = group / GPC;
  if constexpr (Cfg::S_ENTRY_MAJOR) off = sub * GPC + (group % GPC);   // adjacent map columns stay adjacent
  else off = ((slot_map * GPC + (group % GPC)) * 2 + sub) * Cfg::COLSTRIDE;
}

// The two S columns of column group v (= cta * GPC + local): map column v (sub 0) and N - v (sub 1); for
// v == 0 the packed real columns 0 and N/2.  Same result as two s_locate calls, without their case analysis
// (which the compiler cannot fold for a run-time v: ~35 instructions per call site and pass).
template <class Cfg> FAVAE_HD void s_group_offsets(int local, int slot_map, int& off0, int& off1) {
  constexpr int GPC = Cfg::HALF / Cfg::C;
  if constexpr (Cfg::S_ENTRY_MAJOR) { off0 = local; off1 = GPC + local; }
  else {
    off0 = ((slot_map * GPC + local) * 2) * Cfg::COLSTRIDE;
    off1 = off0 + Cfg::COLSTRIDE;
  }
}

// Row-FFT scatter / gather without a table look-up per element.  Register e of lane t holds map column
// w = t + c_e (c_e = idx_out(0, e), a multiple of R2).  With GPC a multiple of R2 the owner CTA and the column
// slot of w are affine in t with compile-time constants per e:
//   c_e <  N/2 :  owner c_e / GPC,              slot  t + c_e % GPC
//   c_e >= N/2 :  group g0 - t (g0 = N - c_e):  owner (g0 - R2) / GPC,  slot  (GPC - t) + (g0 - R2) % GPC + R2
// so six per-thread base entries (sp / sm per owner) plus an immediate reach every element.  The exception
// is lane 0 when g0 is a multiple of GPC (column N/2, which shares group 0, and the first group of an owner):
// those elements (two at N = 256, C = 2) keep their final entry in a register (sx).
template <class Cfg> struct SFast {
  static constexpr int GPC = Cfg::HALF / Cfg::C;
#ifndef FAVAE_FFL_SFAST
#define FAVAE_FFL_SFAST 1
#endif
  static constexpr bool value = FAVAE_FFL_SFAST && Cfg::S_ENTRY_MAJOR && Cfg::C == 2 && Cfg::R2 > 1 && GPC % Cfg::R2 == 0 && Cfg::R1 == Cfg::R2;
};
template <class Cfg, int E> struct SAddr {
  static constexpr int N = Cfg::N, HALF = Cfg::HALF, R1 = Cfg::R1, R2 = Cfg::R2, GPC = HALF / Cfg::C;
  static constexpr int c = R2 * (E / R2) + R1 * (E % R2);          // idx_out(0, E)
  static constexpr bool pos = c < HALF;
  static constexpr int g0 = N - c;
  static constexpr bool special = !pos && (g0 % GPC == 0);
  static constexpr int owner = pos ? c / GPC : (g0 - R2) / GPC;
  static constexpr int delta = pos ? c % GPC : (g0 - R2) % GPC + R2;
  static constexpr int sp_owner = (g0 == HALF) ? 0 : g0 / GPC;     // lane 0 of a special element
  static constexpr int sp_off = GPC;
  static constexpr int specials_before() {
    int n = 0;
    for (int e = 0; e < E; ++e) {
      const int ce = R2 * (e / R2) + R1 * (e % R2);
      if (ce >= HALF && (N - ce) % GPC == 0) ++n;
    }
    return n;
  }
  static constexpr int sidx = specials_before();
};
template <class Cfg, int E = 0, class F> FAVAE_HD void for_each_elem(F&& f) {
  if constexpr (E < Cfg::R1) {
    f(std::integral_constant<int, E>{});
    for_each_elem<Cfg, E + 1>(f);
  }
}

// f(A) from A^2 (already ortho-normalised)
FAVAE_HD float spectrum_f(float a2, float alpha, int log_matrix) {
  float f;
  if (alpha == 1.0f) f = a2 > 0.0f ? a2 * FAVAE_RSQRT(a2) : 0.0f;     // sqrt without the slow path
  else if (alpha == 2.0f) f = a2;
  else f = powf(sqrtf(a2), alpha);
  if (log_matrix) f = logf(f + 1.0f);
  return f;
}

// weight = clamp(nan_to_0(f / fmax), 0, 1) with inv = 1/fmax (0 when fmax == 0: the NaN -> 0 rule)
FAVAE_HD float spectrum_inv(float fmax) { return fmax > 0.0f ? 1.0f / fmax : 0.0f; }
FAVAE_HD float spectrum_w(float f, float inv) { return fminf(fmaxf(f * inv, 0.0f), 1.0f); }

// ----------------------------------------------------------------------------------
// 1-D FFT stages.  stg points at this group's staging area (R1 * STG_STRIDE float2).
// ----------------------------------------------------------------------------------
template <class Cfg> FAVAE_HD void fwd_stage1(ThreadRegs<Cfg>& r, int t, float2* stg) {
  RegFFT<Cfg::R1, -1>::run(r.v);
  if constexpr (Cfg::R2 > 1) {
#pragma unroll
    for (int k1 = 0; k1 < Cfg::R1; ++k1) {
      float2 y = (k1 == 0) ? r.v[0] : cmul(r.v[k1], r.tw[k1]);
      stg[k1 * Cfg::STG_STRIDE + t] = y;
    }
  }
}
template <class Cfg> FAVAE_HD void fwd_stage2(ThreadRegs<Cfg>& r, int t, const float2* stg) {
  if constexpr (Cfg::R2 > 1) {
    constexpr int M = Cfg::R1 / Cfg::R2;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float2 b[Cfg::R2];
      const int k1 = t + Cfg::R2 * m;
#pragma unroll
      for (int n2 = 0; n2 < Cfg::R2; ++n2) b[n2] = stg[k1 * Cfg::STG_STRIDE + n2];
      RegFFT<Cfg::R2, -1>::run(b);
#pragma unroll
      for (int k2 = 0; k2 < Cfg::R2; ++k2) r.v[m * Cfg::R2 + k2] = b[k2];
    }
  }
}
// inverse: consumes out-layout registers, produces in-layout registers (unnormalised)
template <class Cfg> FAVAE_HD void inv_stage1(ThreadRegs<Cfg>& r, int t, float2* stg) {
  if constexpr (Cfg::R2 > 1) {
    constexpr int M = Cfg::R1 / Cfg::R2;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      float2 b[Cfg::R2];
      const int k1 = t + Cfg::R2 * m;
#pragma unroll
      for (int k2 = 0; k2 < Cfg::R2; ++k2) b[k2] = r.v[m * Cfg::R2 + k2];
      RegFFT<Cfg::R2, +1>::run(b);
#pragma unroll
      for (int n2 = 0; n2 < Cfg::R2; ++n2) stg[k1 * Cfg::STG_STRIDE + n2] = b[n2];
    }
  }
}
template <class Cfg> FAVAE_HD void inv_stage2(ThreadRegs<Cfg>& r, int t, const float2* stg) {
  if constexpr (Cfg::R2 > 1) {
#pragma unroll
    for (int k1 = 0; k1 < Cfg::R1; ++k1) {
      float2 y = stg[k1 * Cfg::STG_STRIDE + t];
      r.v[k1] = (k1 == 0) ? y : cmul(y, make_float2(r.tw[k1].x, -r.tw[k1].y));
    }
  }
  RegFFT<Cfg::R1, +1>::run(r.v);
}

}  // namespace favae
