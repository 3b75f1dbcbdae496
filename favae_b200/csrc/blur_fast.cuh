// Streaming separable Gaussian blur for power-of-two map widths (4 <= W <= 512).
//
// A work item is a strip of TH rows of one map.  W/4 threads own four adjacent columns each
// (float4 HBM accesses); the vertical pass is a KS-row sliding window held in registers (the
// row loop is fully unrolled so the window is renamed, never moved); the horizontal pass reads
// the vertically-filtered row from a double-buffered shared-memory line.  Each input element is
// read once from HBM (strip halos hit L2), each output written once.
//
//   MODE_FWD      y  = H V x          reflect padding          (vqgan_fcm.py:35-41)
//   MODE_ADJ      gx = H^T V^T gy     forward data path + O(1) border corrections (see below)
//   MODE_ADJ_SIG  MODE_ADJ plus  d/dsigma <gy, H V x> = <(H'^T V^T + H^T V'^T) gy, x>  with
//                 k' = dk/dsigma, sharing the single sliding window over gy (x is read once,
//                 without halo)                                (autograd of vqgan_fcm.py:20-41)
//   MODE_SIGMA    the same scalar as <gy, (H'V + HV') x>: forward data path on x, gy read once
#pragma once

#include "common.cuh"
#include "ffl_core.cuh"   // packed fp32x2 helpers (pk_add / pk_mul / pk_fma / pk_dup)

// Tuning constants.  The defaults are the measured optima quoted next to each kernel below; they can be
// overridden on the nvcc command line for experiments (profiles/README.md) and nothing else reads them.
#ifndef FAVAE_ADJ_FULL
#define FAVAE_ADJ_FULL 0   // adjoint kernel: 0 = RS rows unrolled inside a rolled loop
#endif
#ifndef FAVAE_ADJ_MINB
#define FAVAE_ADJ_MINB 5   // adjoint kernel: CTAs per SM
#endif
#ifndef FAVAE_FWD_FULL
#define FAVAE_FWD_FULL 1   // forward kernel: all rows of a strip unrolled (the rolled form is 45 % slower there)
#endif
#ifndef FAVAE_FWD_MINB
#define FAVAE_FWD_MINB 5   // forward kernel: CTAs per SM (96 registers)
#endif
#ifndef FAVAE_FWD_Q
#define FAVAE_FWD_Q 6   // forward kernel: rows in flight from HBM per thread (register prefetch)
#endif
#ifndef FAVAE_FWD_TH
#define FAVAE_FWD_TH 32   // forward kernel: strip height
#endif
#ifndef FAVAE_ADJ_Q
#define FAVAE_ADJ_Q 3   // adjoint kernel: rows in flight per thread
#endif
#ifndef FAVAE_DIFF_MINB
#define FAVAE_DIFF_MINB 4   // blur-difference kernel: CTAs per SM at k <= 9 (126 registers)
#endif
#ifndef FAVAE_ADJSIG_MINB
#define FAVAE_ADJSIG_MINB 4   // adjoint + sigma kernel: CTAs per SM (126 registers)
#endif
#ifndef FAVAE_ADJSIG_FULL
#define FAVAE_ADJSIG_FULL 0   // adjoint + sigma kernel: 0 = RS rows unrolled inside a rolled loop
#endif
#ifndef FAVAE_ADJSIG_Q
#define FAVAE_ADJSIG_Q 3   // adjoint + sigma kernel: register prefetch depth when the cp.async rings are off
#endif
#ifndef FAVAE_ADJSIG_XQ
#define FAVAE_ADJSIG_XQ 1   // adjoint + sigma kernel: x rows in flight in registers when the rings are off
#endif
#ifndef FAVAE_ADJSIG_XASYNC
#define FAVAE_ADJSIG_XASYNC 4   // adjoint + sigma kernel: > 0 = x rows through a cp.async shared ring of (up to) that many rows
#endif
#ifndef FAVAE_ADJSIG_GASYNC
#define FAVAE_ADJSIG_GASYNC 4   // adjoint + sigma kernel: gy AND x rows through cp.async rings of that depth (power of two; 0 = off)
#endif
#ifndef FAVAE_SIGMA_MINB
#define FAVAE_SIGMA_MINB 4   // sigma-only kernel (FAVAE_BLUR_SIGMA=split): CTAs per SM
#endif
#ifndef FAVAE_SIGMA_FULL
#define FAVAE_SIGMA_FULL 1   // sigma-only kernel: all rows unrolled
#endif
#ifndef FAVAE_ADJSIG_TH
#define FAVAE_ADJSIG_TH 64   // adjoint + sigma kernel: strip height on maps of >= 128 rows
#endif
#ifndef FAVAE_DIFF_U
#define FAVAE_DIFF_U 0   // blur-difference kernel: rows per rolled iteration (0 = KS, the renamed ring: its 9-row body still fits)
#endif
#ifndef FAVAE_ADJSIG_U
#define FAVAE_ADJSIG_U 4   // adjoint + sigma kernel (cp.async rings): rows per rolled iteration (0 = KS); even values keep the line / ring parities static
#endif
#ifndef FAVAE_PAIR_U
#define FAVAE_PAIR_U 4   // paired adjoint + sigma kernel: rows per rolled iteration (0 = KS, the renamed ring); measured 1: 1.17, 2: 1.05, 3: 1.08, 4: 1.01, 5: 1.09, 6: 1.02 ms
#endif
#ifndef FAVAE_PAIR_TH
#define FAVAE_PAIR_TH 128   // paired adjoint + sigma kernel: strip height on maps of >= 256 rows
#endif
#ifndef FAVAE_ADJ_TH
#define FAVAE_ADJ_TH 64   // adjoint kernel: strip height on maps of >= 128 rows
#endif

namespace favae {
namespace blurf {

constexpr int THREADS = 128;
constexpr int MODE_FWD = 0, MODE_ADJ = 1, MODE_ADJ_SIG = 2, MODE_SIGMA = 3, MODE_PAIR = 4;
constexpr int LPAD = 8;                          // left halo slots of a shared line (>= p, 16B aligned)

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// 16-byte cp.async (L2 only) under a predicate: no branch around the request
__device__ __forceinline__ void cp_async16_if(unsigned smem_addr, const void* gptr, bool pred) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p cp.async.cg.shared.global [%0], [%1], 16; }" ::"r"(smem_addr),
               "l"(gptr), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void fma4(float4& a, float k, const float4& v) {
  a.x = fmaf(k, v.x, a.x); a.y = fmaf(k, v.y, a.y); a.z = fmaf(k, v.z, a.z); a.w = fmaf(k, v.w, a.w);
}

// weights k[t] and dk[t]/dsigma computed by the first KS threads (vqgan_fcm.py:20-26)
template <int KS>
__device__ __forceinline__ void load_weights(const float* sigma_ptr, float* sk, float* sdk, float (&k)[KS],
                                             float (&dk)[KS]) {
  if (threadIdx.x < 32) {
    const float sigma = sigma_ptr[0];
    const float half = (KS - 1) * 0.5f;
    const float x = -half + (float)threadIdx.x;
    const float q = x / sigma;
    const float e = (threadIdx.x < KS) ? expf(-0.5f * q * q) : 0.f;
    const float sum = warp_sum(e);
    const float kk = e / sum;
    const float m2 = warp_sum(kk * x * x);
    if (threadIdx.x < KS) {
      sk[threadIdx.x] = kk;
      sdk[threadIdx.x] = kk * (x * x - m2) / (sigma * sigma * sigma);
    }
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < KS; ++t) { k[t] = sk[t]; dk[t] = sdk[t]; }
}

// Adjoint of (reflect-pad, correlate) as a FORWARD stencil.  With taps symmetric about the centre,
//   H^T g = D (stencil on the mirrored extension of E g),   E = diag(2, 1, ..., 1, 2),  D = diag(1/2, 1, ..., 1, 1/2):
// the reflected copies of samples 1..P land where the mirror puts them; the border sample is never
// reflected, which the forward stencil makes up for by seeing it twice (E); and the border output,
// which would receive every contribution twice, is halved (D).  (Checked against autograd for
// every P < w, including P = w - 1, in tests/test_abi_and_host.py.)  The adjoint therefore runs the
// forward data path with two scalings per border instead of per-row tap masks and corrections.
// adjoint: rolled RS-row unroll, 3 rows in flight, 96 registers (5 CTAs / SM), 64-row strips on large
// maps: 0.47 ms at 4096 maps of 256^2.  (6 rows in flight at 128 registers / 4 CTAs: 0.57 ms; all rows
// unrolled 0.71 ms; 6 rows capped at 96 registers spills: 0.67 ms.)
template <int KS, int TH, int MODE>
__global__ void __launch_bounds__(THREADS, MODE == MODE_ADJ ? FAVAE_ADJ_MINB : FAVAE_FWD_MINB)
blur_fast_kernel(const float* __restrict__ src, const float* __restrict__ /*aux*/, int h, int w, long long items,
                 int strips, const float* __restrict__ sigma, float* __restrict__ dst,
                 float* __restrict__ /*partials*/, float oscale, const float* __restrict__ oscale_dev) {
  if (oscale_dev) oscale *= oscale_dev[0];       // upstream gradient of the fused DSL level (device scalar)
  // src: x (FWD) or gy (ADJ); dst: y / gx
  static_assert(MODE == MODE_FWD || MODE == MODE_ADJ, "sigma-gradient modes have their own kernels");
  constexpr int P = KS / 2;
  constexpr bool ADJ = MODE == MODE_ADJ;
  extern __shared__ float lines[];               // [2 buffers][groups][w + 2*LPAD]
  __shared__ float sk[32], sdk[32];
  float k[KS], dk[KS];
  load_weights<KS>(sigma, sk, sdk, k, dk);

  const int tpi = w >> 2;                        // threads per item
  const int groups = THREADS / tpi;
  const int grp = threadIdx.x / tpi, tx = threadIdx.x % tpi;
  const int x0 = tx * 4;
  const int ll = w + 2 * LPAD;
  const long long item = (long long)blockIdx.x * groups + grp;
  const bool live = item < items;
  const long long map = live ? item / strips : 0;
  const int y0 = live ? (int)(item % strips) * TH : 0;
  const float* base = src + map * (long long)h * w;
  const float e0 = (ADJ && tx == 0) ? 2.f : 1.f, e3 = (ADJ && tx == tpi - 1) ? 2.f : 1.f;   // E, columns
  const float d0 = (ADJ && tx == 0) ? 0.5f : 1.f, d3 = (ADJ && tx == tpi - 1) ? 0.5f : 1.f; // D, columns

  // Register ring of RS = KS + Q rows: the KS-row window of the vertical pass plus Q rows in flight
  // from HBM (explicit prefetch: 6 rows per thread for the forward kernel, whose NR rows are all
  // unrolled; 3 for the adjoint, which trades prefetch depth for a fifth CTA per SM and unrolls RS
  // rows inside a rolled loop).  Every ring index is a constant.  (A cp.async shared-memory ring as
  // in blur_adjsig_kernel was measured here too: no gain for either mode - 0.474 / 0.393 ms against
  // 0.469 / 0.394 ms - these two kernels are not waiting for their loads.)
  constexpr int Q = ADJ ? FAVAE_ADJ_Q : FAVAE_FWD_Q, RS = KS + Q, NR = TH + KS - 1;
  const long long mapoff = map * (long long)h * w;
  auto load_row = [&](int r) -> float4 {
    const int ry = reflect_idx(y0 - P + r, h);
    float4 in = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) in = ld4(base + (long long)ry * w + x0);
    return in;
  };
  float4 ring[RS];
#pragma unroll
  for (int q = 0; q < RS; ++q) ring[q] = (q < Q) ? load_row(q) : make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int STEP = ((ADJ && !FAVAE_ADJ_FULL) || (!ADJ && !FAVAE_FWD_FULL)) ? RS : NR;
#pragma unroll 1
  for (int r0 = 0; r0 < NR; r0 += STEP)
#pragma unroll
  for (int uu = 0; uu < STEP; ++uu) {
    const int r = r0 + uu;
    const int u = STEP == RS ? uu : uu % RS;
    if (r >= NR) break;
    if (r + Q < NR) ring[(u + Q) % RS] = load_row(r + Q);
    if (ADJ) {
      // E, rows: doubled when the row enters the window (Q iterations after its load was issued:
      // scaling at load time would make every iteration wait for the row it has just requested)
      const int ry = reflect_idx(y0 - P + r, h);
      if (ry == 0 || ry == h - 1) { float4& b = ring[u]; b.x *= 2.f; b.y *= 2.f; b.z *= 2.f; b.w *= 2.f; }
    }
    if (r < KS - 1) continue;
    const int yo = y0 + r - (KS - 1);            // output row of this iteration
    // ---- vertical pass over the window rows yo - P .. yo + P
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < KS; ++t) fma4(v, k[t], ring[(u + RS - (KS - 1) + t) % RS]);
    if (ADJ) { v.x *= e0; v.w *= e3; }
    // ---- horizontal pass through a shared line
    float* line = lines + (size_t)((r & 1) * groups + grp) * ll;
    *reinterpret_cast<float4*>(line + LPAD + x0) = v;
    {
      // mirrored halo entries (reflect): index -j <- j, index w-1+j <- w-1-j
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = x0 + c;
        if (j >= 1 && j <= P) line[LPAD - j] = vv[c];
        const int jr = w - 1 - j;
        if (jr >= 1 && jr <= P) line[LPAD + w - 1 + jr] = vv[c];
      }
    }
    __syncthreads();
    float seg[4 + 2 * P];
#pragma unroll
    for (int i = 0; i < 4 + 2 * P; ++i) seg[i] = line[LPAD + x0 - P + i];
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int t = 0; t < KS; ++t) o[c] = fmaf(k[t], seg[c + t], o[c]);
    }
    if (ADJ) {
      const float dr = ((yo == 0 || yo == h - 1) ? 0.5f : 1.f) * oscale;   // D, rows (and the caller's scale)
      o[0] *= dr * d0; o[1] *= dr; o[2] *= dr; o[3] *= dr * d3;
    }
    if (live && yo < h)
      *reinterpret_cast<float4*>(dst + mapoff + (long long)yo * w + x0) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Reflect halo of a two-plane shared line of float2 pairs (plane A: columns 4q, 4q+1 of thread q; plane
// B: columns 4q+2, 4q+3): column -j <- j and column w-1+j <- w-1-j for j = 1..P.  Only the two threads at
// either end of a row own such columns, and relative to the writer's own slot every destination is a
// compile-time constant, so the halo costs four thread-role tests per row plus a few stores with
// immediate offsets on the border threads (the generic per-column test / index arithmetic was ~40 of
// the ~210 instructions of a row iteration in every warp, since each 32-thread half row owns a border).
// pa / pb: the writer's own float4 slots in the two planes.
template <int REL> __device__ __forceinline__ void halo_put_rel(float4* pa, float4* pb, float2 v) {
  constexpr int Q = REL >= 0 ? REL / 4 : -((-REL + 3) / 4);   // float4 slot relative to the writer's
  constexpr int C = REL - 4 * Q;                               // column inside that slot
  reinterpret_cast<float2*>((C < 2 ? pa : pb) + Q)[C & 1] = v;
}
template <int P> __device__ __forceinline__ void halo_puts(float4* pa, float4* pb, int tx, int tpi,
                                                           const float2 (&acc)[4]) {
  static_assert(P <= 7, "two border threads per side cover at most 7 halo columns");
  if (tx == 0) {                                   // columns 1..3 -> -1..-3
    if constexpr (P >= 1) halo_put_rel<-1>(pa, pb, acc[1]);
    if constexpr (P >= 2) halo_put_rel<-2>(pa, pb, acc[2]);
    if constexpr (P >= 3) halo_put_rel<-3>(pa, pb, acc[3]);
  }
  if (tx == 1) {                                   // columns 4..7 -> -4..-7 (own x0 = 4)
    if constexpr (P >= 4) halo_put_rel<-8>(pa, pb, acc[0]);
    if constexpr (P >= 5) halo_put_rel<-9>(pa, pb, acc[1]);
    if constexpr (P >= 6) halo_put_rel<-10>(pa, pb, acc[2]);
    if constexpr (P >= 7) halo_put_rel<-11>(pa, pb, acc[3]);
  }
  if (tx == tpi - 1) {                             // columns w-2..w-4 -> w..w+2 (own x0 = w - 4)
    if constexpr (P >= 1) halo_put_rel<4>(pa, pb, acc[2]);
    if constexpr (P >= 2) halo_put_rel<5>(pa, pb, acc[1]);
    if constexpr (P >= 3) halo_put_rel<6>(pa, pb, acc[0]);
  }
  if (tx == tpi - 2) {                             // columns w-5..w-8 -> w+3..w+6 (own x0 = w - 8)
    if constexpr (P >= 4) halo_put_rel<11>(pa, pb, acc[3]);
    if constexpr (P >= 5) halo_put_rel<12>(pa, pb, acc[2]);
    if constexpr (P >= 6) halo_put_rel<13>(pa, pb, acc[1]);
    if constexpr (P >= 7) halo_put_rel<14>(pa, pb, acc[0]);
  }
}

// d = B_dec(dec) - B_enc(enc): the blurred pair of the DSL (models/vqgan_fcm.py:131-134 and the codec
// call sites, consumed only by ffl(de, en), losses/vqgan_losses.py:25) written as ONE map.  The spectrum
// loss sees pred - target only, so the two blurred maps never need to exist: 12 B/element (read enc, read
// dec, write d) instead of 16 for two blurs, and the loss kernel then reads 4 B/element instead of 8.
//
// Both maps go through ONE pipeline in lockstep as packed (enc, dec) pairs: every register pair holds
// the two maps' values of one pixel and every tap is the pair (k_enc[t], k_dec[t]), so one FFMA2 / FADD2
// filters both maps (same trick as the (value, d/dsigma) pairs of blur_adjsig_kernel, whose structure
// this kernel shares: cp.async shared-memory rings for the two input streams, KS-row window of pairs in
// registers, mirrored taps added first, two-plane shared line, KS rows unrolled inside a rolled loop).
// First version, kept for the record: the forward kernel run twice over the strip by the same threads
// (pass 0 leaves B_enc(enc) in dst, pass 1 subtracts it): 1.13 ms at 4096 maps of 256^2 with a plain
// re-load (one exposed L2 round trip per row), 1.00 ms with a cp.async ring, 0.99 ms with evict-first
// input loads (ncu: 13.5 B/element of DRAM traffic, 166 M warp instructions per 1024 maps, issue active
// 60 %, `no_instruction` 1.5 cycles per issue: two scalar passes of a fully unrolled 40-row body are
// instruction bound).
template <int KS, int TH>
__global__ void __launch_bounds__(THREADS, KS <= 9 ? FAVAE_DIFF_MINB : KS == 11 ? 3 : 2)   // the window is 8 KS registers
blur_diff_kernel(const float* __restrict__ enc, const float* __restrict__ dec, int h, int w, long long items,
                 int strips, const float* __restrict__ sigma_enc, const float* __restrict__ sigma_dec,
                 float* __restrict__ dst) {
  constexpr int P = KS / 2;
  extern __shared__ float lines[];               // as float2: [2 buffers][groups][w + 2*LPAD]
  __shared__ float sk[2][32], sdk[32];
  float2 kk[P + 1];                              // (k_enc[t], k_dec[t]); tap t and tap KS-1-t are equal
  {
    float k[KS], dk[KS];
    load_weights<KS>(sigma_enc, sk[0], sdk, k, dk);
    load_weights<KS>(sigma_dec, sk[1], sdk, k, dk);
#pragma unroll
    for (int t = 0; t <= P; ++t) kk[t] = make_float2(sk[0][t], sk[1][t]);
  }
  const int tpi = w >> 2;                        // threads per item
  const int groups = THREADS / tpi;
  const int grp = threadIdx.x / tpi, tx = threadIdx.x % tpi;
  const int x0 = tx * 4;
  const int ll = w + 2 * LPAD;                   // float2 per line
  float2* lines2 = reinterpret_cast<float2*>(lines);
  const long long item = (long long)blockIdx.x * groups + grp;
  const bool live = item < items;
  const long long map = live ? item / strips : 0;
  const int y0 = live ? (int)(item % strips) * TH : 0;
  const long long mapoff = map * (long long)h * w;
  const float* ebase = enc + mapoff;
  const float* dbase = dec + mapoff;
  // rows per rolled iteration: see blur_adjsig_pair_kernel (U < KS: shifted ring, small body)
  constexpr int U = (FAVAE_DIFF_U > 0 && FAVAE_DIFF_U < KS) ? FAVAE_DIFF_U : KS;
  constexpr bool SHIFT = U != KS;
  constexpr int GD = 4, RS = SHIFT ? KS - 1 + U : KS, NR = TH + KS - 1;
  // one array for both rings: a slot of the second stream sits a constant GD * SLOT_B bytes after the
  // first one's, and everything that does not depend on the row is computed once (the per-row request
  // sequence was ~45 instructions of a ~190-instruction row iteration)
  __shared__ float4 rings[2][GD][THREADS];
  float4 (*ering)[THREADS] = rings[0];
  float4 (*dring)[THREADS] = rings[1];
  constexpr unsigned SLOT_B = THREADS * sizeof(float4);
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(&rings[0][0][threadIdx.x]);
  const float* ecol = ebase + x0;
  const float* dcol = dbase + x0;
  auto issue_rows = [&](int r) {
    const bool ok = live && r < NR;
    const unsigned sa = ring_s + (unsigned)(r & (GD - 1)) * SLOT_B;
    const int off = reflect_idx(y0 - P + r, h) * w;          // < h * w <= 2^18
    cp_async16_if(sa, ecol + off, ok);
    cp_async16_if(sa + GD * SLOT_B, dcol + off, ok);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int q = 0; q < GD - 1; ++q) issue_rows(q);
  float2 ring[RS][4];
#pragma unroll
  for (int q = 0; q < RS; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) ring[q][c] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int r0 = 0; r0 < NR; r0 += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u;
      if (NR % U != 0 && r >= NR) break;
      issue_rows(r + GD - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(GD - 1) : "memory");
      {
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f), d = e;
        if (live) { e = ering[r & (GD - 1)][threadIdx.x]; d = dring[r & (GD - 1)][threadIdx.x]; }
        float2 (&nw)[4] = ring[SHIFT ? KS - 1 + u : u];
        nw[0] = make_float2(e.x, d.x); nw[1] = make_float2(e.y, d.y);
        nw[2] = make_float2(e.z, d.z); nw[3] = make_float2(e.w, d.w);
      }
      if (r < KS - 1) continue;
      const int yo = y0 + r - (KS - 1);            // output row of this iteration
#define FAVAE_WIN(t) ring[SHIFT ? u + (t) : (u + 1 + (t)) % RS]
      // ---- vertical pass over the mirrored window, both maps per instruction
      float2 acc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[c] = pk_mul(FAVAE_WIN(P)[c], kk[P]);
#pragma unroll
      for (int t = 0; t < P; ++t)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] = pk_fma(pk_add(FAVAE_WIN(t)[c], FAVAE_WIN(KS - 1 - t)[c]), kk[t], acc[c]);
#undef FAVAE_WIN
      // ---- horizontal pass through a shared line of (enc, dec) pairs: two planes of float4 as in
      // blur_adjsig_kernel (plane A: columns 4q, 4q+1 of thread q; plane B: columns 4q+2, 4q+3)
      constexpr int LP = 2, NB = (P + 3) / 4;
      float4* planeA = reinterpret_cast<float4*>(lines2 + (size_t)((r & 1) * groups + grp) * ll);
      float4* planeB = planeA + (tpi + 2 * LP);
      planeA[LP + tx] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
      planeB[LP + tx] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
      halo_puts<P>(planeA + LP + tx, planeB + LP + tx, tx, tpi, acc);
      __syncthreads();
      float2 cols[4 * (2 * NB + 1)];               // columns x0 - 4 NB .. x0 + 4 NB + 3; own from registers
#pragma unroll
      for (int q = -NB; q <= NB; ++q) {
        float2* dd = cols + 4 * (q + NB);
        if (q == 0) { dd[0] = acc[0]; dd[1] = acc[1]; dd[2] = acc[2]; dd[3] = acc[3]; }
        else {
          const float4 a = planeA[LP + tx + q], b = planeB[LP + tx + q];
          dd[0] = make_float2(a.x, a.y); dd[1] = make_float2(a.z, a.w);
          dd[2] = make_float2(b.x, b.y); dd[3] = make_float2(b.z, b.w);
        }
      }
      const float2* seg = cols + (4 * NB - P);     // seg[i] = column x0 - P + i
      float o[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float2 a2 = pk_mul(seg[c + P], kk[P]);
#pragma unroll
        for (int t = 0; t < P; ++t) a2 = pk_fma(pk_add(seg[c + t], seg[c + KS - 1 - t]), kk[t], a2);
        o[c] = a2.y - a2.x;                        // B_dec(dec) - B_enc(enc)
      }
      if (live && yo < h)
        __stcs(reinterpret_cast<float4*>(dst + mapoff + (long long)yo * w + x0), make_float4(o[0], o[1], o[2], o[3]));
    }
    if constexpr (SHIFT) {
#pragma unroll
      for (int q = 0; q < KS - 1; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) ring[q][c] = ring[q + U][c];
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// MODE_ADJ_SIG rewritten around packed fp32x2 arithmetic (FADD2 / FFMA2 issue two fp32 lanes per
// slot).  Every intermediate is carried as the pair (value filtered with k, value filtered with
// dk/dsigma): the vertical pass accumulates (v, v') per column from the window sample broadcast
// against the tap pair (k[t], k'[t]); the shared line holds those pairs interleaved; the horizontal
// pass forms (sum_t k[t] v, sum_t k[t] v') packed plus sum_t k'[t] v scalar.  The taps are symmetric,
// so mirrored window samples are added first (P + 1 multiplies per output instead of 2P + 1).  The
// adjoint borders are the E / D scalings derived above blur_fast_kernel (they apply to k and k'
// alike).  Same data path (register ring, double-buffered line, one barrier per row), well under
// half the instructions of the scalar formulation.
// Occupancy is what this kernel needs: with 3 rows of gy in flight (instead of 6) and the x row
// requested in the iteration that uses it, the kernel fits 126 registers without spills, i.e. 4 CTAs
// (16 warps) per SM: 0.88 ms at 4096 maps of 256^2 against 1.11 ms at 160 registers / 3 CTAs.
// (Capping the 6-row version at 128 registers spilled: 1.29 ms; all rows unrolled 1.26 ms.  The
// stall that remained was the x row (requested one iteration ahead into registers) arriving late;
// two x rows in registers spill at the 128-register cap (0.90 ms), L2 prefetch hints change
// nothing, so x now travels through a cp.async shared-memory ring 3 iterations ahead: 0.74 ms.)
template <int KS, int TH>
__global__ void __launch_bounds__(THREADS, FAVAE_ADJSIG_MINB)
blur_adjsig_kernel(const float* __restrict__ src, const float* __restrict__ aux, int h, int w, long long items,
                   int strips, const float* __restrict__ sigma, float* __restrict__ dst,
                   float* __restrict__ partials, float oscale, const float* __restrict__ oscale_dev) {
  if (oscale_dev) oscale *= oscale_dev[0];       // upstream gradient of the fused DSL level (device scalar)
  // oscale multiplies gx and the sigma partials (the fused DSL op feeds +G to the decoder side and -G
  // to the encoder side, vqgan_losses.py:25: ffl(de, en)); it rides on the D row scaling
  constexpr int P = KS / 2;
  extern __shared__ float lines[];               // as float2: [2 buffers][groups][w + 2*LPAD]
  __shared__ float sk[32], sdk[32];
  __shared__ float wred[THREADS / 32];
  float k[KS], dk[KS];
  load_weights<KS>(sigma, sk, sdk, k, dk);
  float2 kd[P + 1];                              // (k[t], k'[t]); tap t and tap KS-1-t are equal
#pragma unroll
  for (int t = 0; t <= P; ++t) kd[t] = make_float2(k[t], dk[t]);

  const int tpi = w >> 2;                        // threads per item
  const int groups = THREADS / tpi;
  const int grp = threadIdx.x / tpi, tx = threadIdx.x % tpi;
  const int x0 = tx * 4;
  const int ll = w + 2 * LPAD;                   // float2 per line
  float2* lines2 = reinterpret_cast<float2*>(lines);
  const long long item = (long long)blockIdx.x * groups + grp;
  const bool live = item < items;
  const long long map = live ? item / strips : 0;
  const int y0 = live ? (int)(item % strips) * TH : 0;
  const float* base = src + map * (long long)h * w;
  float acc_sigma = 0.f;

// gy AND x rows through cp.async shared-memory rings of that depth (a power of two; 0: register
// prefetch).  Measured at 4096 maps of 256^2: register prefetch 0.79 ms, x ring only 0.74 ms, both
// rings depth 4 0.66 ms, depth 8 0.68 ms (and 33 KB of static shared memory, over the 48 KB launch
// limit together with the line buffers of the narrow maps).
  constexpr int XQ = FAVAE_ADJSIG_XQ;
  constexpr int GD = FAVAE_ADJSIG_GASYNC;
  // rows per rolled iteration with the cp.async rings: see blur_adjsig_pair_kernel (U < KS: shifted ring)
  constexpr int U = (GD && !FAVAE_ADJSIG_FULL && FAVAE_ADJSIG_U > 0 && FAVAE_ADJSIG_U < KS) ? FAVAE_ADJSIG_U : KS;
  constexpr bool SHIFT = U != KS;
  constexpr int Q = GD ? 0 : FAVAE_ADJSIG_Q + ((XQ - (KS + FAVAE_ADJSIG_Q) % XQ) % XQ);
  constexpr int RS = SHIFT ? KS - 1 + U : KS + Q, NR = TH + KS - 1;
  static_assert(RS % XQ == 0, "x prefetch ring must tile the unroll factor");
  const long long mapoff = map * (long long)h * w;
  auto load_row = [&](int r) -> float4 {
    const int ry = reflect_idx(y0 - P + r, h);
    float4 in = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) in = ld4(base + (long long)ry * w + x0);
    return in;
  };
  const float e0 = tx == 0 ? 2.f : 1.f, e3 = tx == tpi - 1 ? 2.f : 1.f;     // E, columns
  const float d0 = tx == 0 ? 0.5f : 1.f, d3 = tx == tpi - 1 ? 0.5f : 1.f;   // D, columns
  auto load_x = [&](int r) -> float4 {           // x row of the output produced at iteration r
    const int yo = y0 + r - (KS - 1);
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live && r >= KS - 1 && yo < h) xv = ld4(aux + mapoff + (long long)yo * w + x0);
    return xv;
  };
  float4 ring[RS];
#pragma unroll
  for (int q = 0; q < RS; ++q) ring[q] = (q < Q) ? load_row(q) : make_float4(0.f, 0.f, 0.f, 0.f);
#if FAVAE_ADJSIG_GASYNC
  // Both input streams travel HBM -> shared memory by cp.async, GD - 1 iterations ahead, holding no
  // registers: one commit group per iteration carries the gy row that enters the window and the x row
  // of that iteration's output, so wait_group GD - 1 means "this iteration's rows have landed".  Each
  // thread reads back only the 16 bytes it requested itself.
  static_assert((GD & (GD - 1)) == 0, "ring depth must be a power of two");
  __shared__ float4 rings[2][GD][THREADS];        // one array: constant distance between the two streams' slots
  float4 (*gring)[THREADS] = rings[0];
  float4 (*xring)[THREADS] = rings[1];
  constexpr unsigned SLOT_B = THREADS * sizeof(float4);
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(&rings[0][0][threadIdx.x]);
  const float* gcol = base + x0;
  const float* xcol = aux ? aux + mapoff + x0 : nullptr;
  auto issue_rows = [&](int r) {
    const bool ok = live && r < NR;
    const unsigned sa = ring_s + (unsigned)(r & (GD - 1)) * SLOT_B;
    cp_async16_if(sa, gcol + reflect_idx(y0 - P + r, h) * w, ok);
    const int yo = y0 + r - (KS - 1);
    cp_async16_if(sa + GD * SLOT_B, xcol + yo * w, ok && r >= KS - 1 && yo < h);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int q = 0; q < GD - 1; ++q) issue_rows(q);
#elif FAVAE_ADJSIG_XASYNC
  // x rows travel HBM -> shared memory by cp.async, XD - 1 iterations ahead and without holding
  // registers (one commit group per iteration, so wait_group XD - 1 means "the row of this iteration
  // has landed"); each thread reads back only the 16 bytes it requested itself.
  constexpr int XW = FAVAE_ADJSIG_XASYNC;        // wanted depth; the ring must tile the unroll factor
  constexpr int XD = (XW >= 6 && RS % 6 == 0) ? 6 : (XW >= 4 && RS % 4 == 0) ? 4 : (RS % 3 == 0) ? 3 : 2;
  static_assert(RS % XD == 0, "x ring must tile the unroll factor");
  __shared__ float4 xring[XD][THREADS];
  auto issue_x = [&](int r, int slot) {
    const int yo = y0 + r - (KS - 1);
    if (live && r >= KS - 1 && r < NR && yo < h) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&xring[slot][threadIdx.x]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(aux + mapoff + (long long)yo * w + x0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int q = 0; q < XD - 1; ++q) issue_x(q, q % XD);
#else
  float4 xq[XQ];
#pragma unroll
  for (int q = 0; q < XQ; ++q) xq[q] = load_x(q);
#endif
  constexpr int STEP = FAVAE_ADJSIG_FULL ? NR : SHIFT ? U : RS;
#pragma unroll 1
  for (int r0 = 0; r0 < NR; r0 += STEP) {
#pragma unroll
    for (int uu = 0; uu < STEP; ++uu) {
      const int r = r0 + uu;
      const int u = SHIFT ? KS - 1 + uu : STEP == RS ? uu : uu % RS;     // ring slot of the row that enters
      if (r >= NR) break;
#if FAVAE_ADJSIG_GASYNC
      issue_rows(r + GD - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(GD - 1) : "memory");
      ring[u] = live ? gring[r & (GD - 1)][threadIdx.x] : make_float4(0.f, 0.f, 0.f, 0.f);
#else
      if (r + Q < NR) ring[(u + Q) % RS] = load_row(r + Q);
#endif
#if FAVAE_ADJSIG_GASYNC
#elif FAVAE_ADJSIG_XASYNC
      issue_x(r + XD - 1, (uu + XD - 1) % XD);
#else
      const float4 xrow = xq[u % XQ];
      if (r + XQ < NR) xq[u % XQ] = load_x(r + XQ);
#endif
      {                                            // E, rows: doubled when the row enters the window
        const int ry = reflect_idx(y0 - P + r, h);
        if (ry == 0 || ry == h - 1) { float4& b = ring[u]; b.x *= 2.f; b.y *= 2.f; b.z *= 2.f; b.w *= 2.f; }
      }
      if (r < KS - 1) continue;
      const int yo = y0 + r - (KS - 1);            // output row of this iteration
#define FAVAE_WIN(t) ring[SHIFT ? uu + (t) : (u + RS - (KS - 1) + (t)) % RS]
      // ---- vertical pass: acc[c] = (sum_t k[t] g, sum_t k'[t] g) over the mirrored window
      float2 acc[4];
      {
        const float4& m = FAVAE_WIN(P);
        acc[0] = pk_mul(pk_dup(m.x), kd[P]); acc[1] = pk_mul(pk_dup(m.y), kd[P]);
        acc[2] = pk_mul(pk_dup(m.z), kd[P]); acc[3] = pk_mul(pk_dup(m.w), kd[P]);
#pragma unroll
        for (int t = 0; t < P; ++t) {
          const float4& a = FAVAE_WIN(t);
          const float4& b = FAVAE_WIN(KS - 1 - t);
          const float2 s01 = pk_add(make_float2(a.x, a.y), make_float2(b.x, b.y));
          const float2 s23 = pk_add(make_float2(a.z, a.w), make_float2(b.z, b.w));
          acc[0] = pk_fma(pk_dup(s01.x), kd[t], acc[0]); acc[1] = pk_fma(pk_dup(s01.y), kd[t], acc[1]);
          acc[2] = pk_fma(pk_dup(s23.x), kd[t], acc[2]); acc[3] = pk_fma(pk_dup(s23.y), kd[t], acc[3]);
        }
      }
      acc[0] = pk_mul(acc[0], pk_dup(e0)); acc[3] = pk_mul(acc[3], pk_dup(e3));
#undef FAVAE_WIN
      // ---- horizontal pass through a shared line of (v, v') pairs
      // Two planes of float4: plane A holds the pairs of columns 4q, 4q+1 of thread q,
      // plane B those of columns 4q+2, 4q+3, LP halo slots on either side.  Every 128-bit access of a
      // quarter warp then covers 128 contiguous bytes (the interleaved single line had a 32-byte thread
      // stride: two-way bank conflicts on every access, 67 % shared-pipe utilisation).
      constexpr int LP = 2, NB = (P + 3) / 4;
      float4* planeA = reinterpret_cast<float4*>(lines2 + (size_t)((r & 1) * groups + grp) * ll);
      float4* planeB = planeA + (tpi + 2 * LP);
      planeA[LP + tx] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
      planeB[LP + tx] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
      halo_puts<P>(planeA + LP + tx, planeB + LP + tx, tx, tpi, acc);
      __syncthreads();
      float2 cols[4 * (2 * NB + 1)];               // columns x0 - 4 NB .. x0 + 4 NB + 3; own from registers
#pragma unroll
      for (int q = -NB; q <= NB; ++q) {
        float2* d = cols + 4 * (q + NB);
        if (q == 0) { d[0] = acc[0]; d[1] = acc[1]; d[2] = acc[2]; d[3] = acc[3]; }
        else {
          const float4 a = planeA[LP + tx + q], b = planeB[LP + tx + q];
          d[0] = make_float2(a.x, a.y); d[1] = make_float2(a.z, a.w);
          d[2] = make_float2(b.x, b.y); d[3] = make_float2(b.z, b.w);
        }
      }
      const float2* seg = cols + (4 * NB - P);     // seg[i] = column x0 - P + i
      float o[4], z[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float2 a2 = pk_mul(pk_dup(kd[P].x), seg[c + P]);       // (sum k v, sum k v')
        float za = kd[P].y * seg[c + P].x;                      // sum k' v
#pragma unroll
        for (int t = 0; t < P; ++t) {
          const float2 q = pk_add(seg[c + t], seg[c + KS - 1 - t]);
          a2 = pk_fma(pk_dup(kd[t].x), q, a2);
          za = fmaf(kd[t].y, q.x, za);
        }
        o[c] = a2.x; z[c] = a2.y + za;
      }
      {
        const float dr = ((yo == 0 || yo == h - 1) ? 0.5f : 1.f) * oscale;   // D, rows (and the caller's scale)
        o[0] *= dr * d0; o[1] *= dr; o[2] *= dr; o[3] *= dr * d3;
        z[0] *= dr * d0; z[1] *= dr; z[2] *= dr; z[3] *= dr * d3;
      }
      if (live && yo < h) {
        *reinterpret_cast<float4*>(dst + mapoff + (long long)yo * w + x0) = make_float4(o[0], o[1], o[2], o[3]);
#if FAVAE_ADJSIG_GASYNC
        const float4 xrow = xring[r & (GD - 1)][threadIdx.x];
#elif FAVAE_ADJSIG_XASYNC
        asm volatile("cp.async.wait_group %0;" ::"n"(XD - 1) : "memory");
        const float4 xrow = xring[uu % XD][threadIdx.x];
#endif
        acc_sigma = fmaf(xrow.x, z[0], fmaf(xrow.y, z[1], fmaf(xrow.z, z[2], fmaf(xrow.w, z[3], acc_sigma))));
      }
    }
    if constexpr (SHIFT) {
#pragma unroll
      for (int q = 0; q < KS - 1; ++q) ring[q] = ring[q + U];
    }
  }
  acc_sigma = warp_sum(acc_sigma);
  if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = acc_sigma;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) s += wred[i];
    partials[blockIdx.x] = s;
  }
}

// Both sides of a fused DSL level in ONE pass over G (favae_b200/spectrum_dsl.py): the decoder side gets
// +s A_dec(G) and the encoder side -s A_enc(G) (A = adjoint blur; s = the upstream gradient of the level),
// each with its sigma gradient.  blur_adjsig_kernel run twice reads G twice (24 B/element for the two
// launches); here the G rows, the window, the E row scaling and the mirrored pre-adds of the vertical
// pass are shared and only the tap-dependent work is done per side: 20 B/element and ~19 % fewer
// instructions than two launches.  Same structure otherwise (cp.async rings for G and the two x streams,
// one barrier per row for both sides' shared lines, taps in uniform registers).  First version (KS rows
// unrolled over a renamed ring), ncu at 1024 maps of 256^2: 208.5 M warp instructions against 2 x 120.6 M,
// 19.9 B/element, issue active 59 %, but `no_instruction` 1.5 cycles per issue -- the 9-row body was 4.4 k
// instructions (70 KB).  Rolling the two side loops (taps re-read from shared memory, own columns re-read
// from the line) cured the fetch stalls and cost more than it saved: 3.3 k static / 238.6 M executed
// instructions, 1.32 ms against 1.23 ms.  What did work: U rows per rolled trip over a ring that is shifted
// with register moves (below), 128-row strips and predicated row requests: 186.6 M instructions, no fetch
// stalls, 1.01 ms = 0.81 of HBM peak (profiles/ncu_r2_summary.md, last section).
template <int KS, int TH>
__global__ void __launch_bounds__(THREADS, KS <= 9 ? FAVAE_ADJSIG_MINB : KS == 11 ? 3 : 2)
blur_adjsig_pair_kernel(const float* __restrict__ gy, const float* __restrict__ x_enc, const float* __restrict__ x_dec,
                        int h, int w, long long items, int strips, const float* __restrict__ sigma_enc,
                        const float* __restrict__ sigma_dec, float* __restrict__ g_enc, float* __restrict__ g_dec,
                        float* __restrict__ partials, const float* __restrict__ scale_dev) {
  constexpr int P = KS / 2;
  extern __shared__ float lines[];               // as float2: [2 sides][2 buffers][groups][w + 2*LPAD]
  __shared__ float sk[2][32], sdk[2][32];
  __shared__ float wred[2][THREADS / 32];
  float2 kd[2][P + 1];                           // side 0 = encoder, 1 = decoder: (k[t], k'[t])
  {
    float k[KS], dk[KS];
    load_weights<KS>(sigma_enc, sk[0], sdk[0], k, dk);
    load_weights<KS>(sigma_dec, sk[1], sdk[1], k, dk);
#pragma unroll
    for (int sd = 0; sd < 2; ++sd)
#pragma unroll
      for (int t = 0; t <= P; ++t) kd[sd][t] = make_float2(sk[sd][t], sdk[sd][t]);
  }
  const float sc = scale_dev ? scale_dev[0] : 1.0f;
  const float oscale[2] = {-sc, sc};             // ffl(pred = blur(dec), target = blur(enc))

  const int tpi = w >> 2;
  const int groups = THREADS / tpi;
  const int grp = threadIdx.x / tpi, tx = threadIdx.x % tpi;
  const int x0 = tx * 4;
  const int ll = w + 2 * LPAD;
  float2* lines2 = reinterpret_cast<float2*>(lines);
  const long long item = (long long)blockIdx.x * groups + grp;
  const bool live = item < items;
  const long long map = live ? item / strips : 0;
  const int y0 = live ? (int)(item % strips) * TH : 0;
  const long long mapoff = map * (long long)h * w;
  const float* base = gy + mapoff;
  const float* xbase[2] = {x_enc + mapoff, x_dec + mapoff};
  float* gout[2] = {g_enc + mapoff, g_dec + mapoff};
  float acc_sigma[2] = {0.f, 0.f};
  const float e0 = tx == 0 ? 2.f : 1.f, e3 = tx == tpi - 1 ? 2.f : 1.f;     // E, columns
  const float d0 = tx == 0 ? 0.5f : 1.f, d3 = tx == tpi - 1 ? 0.5f : 1.f;   // D, columns

  // Rows per rolled iteration.  U == KS: the window ring is renamed (every ring index a constant), at the
  // price of a KS-row body.  U < KS: the ring holds KS - 1 + U rows and is shifted down by U rows with
  // register moves after every iteration (4 (KS - 1) / U moves per row), which keeps the body small.
  constexpr int U = (FAVAE_PAIR_U > 0 && FAVAE_PAIR_U < KS) ? FAVAE_PAIR_U : KS;
  constexpr bool SHIFT = U != KS;
  constexpr int GD = 4, RS = SHIFT ? KS - 1 + U : KS, NR = TH + KS - 1;
  // one array for the three rings (G, x_enc, x_dec): constant distances between the streams' slots, and
  // everything that does not depend on the row computed once (see blur_diff_kernel)
  __shared__ float4 rings[3][GD][THREADS];
  float4 (*gring)[THREADS] = rings[0];
  float4 (*xring)[GD][THREADS] = &rings[1];
  constexpr unsigned SLOT_B = THREADS * sizeof(float4);
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(&rings[0][0][threadIdx.x]);
  const float* gcol = base + x0;
  const float* xcol[2] = {xbase[0] + x0, xbase[1] + x0};
  auto issue_rows = [&](int r) {
    const bool ok = live && r < NR;
    const unsigned sa = ring_s + (unsigned)(r & (GD - 1)) * SLOT_B;
    cp_async16_if(sa, gcol + reflect_idx(y0 - P + r, h) * w, ok);
    const int yo = y0 + r - (KS - 1);
    const bool okx = ok && r >= KS - 1 && yo < h;
    const int xoff = yo * w;
    cp_async16_if(sa + GD * SLOT_B, xcol[0] + xoff, okx);
    cp_async16_if(sa + 2 * GD * SLOT_B, xcol[1] + xoff, okx);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int q = 0; q < GD - 1; ++q) issue_rows(q);
  float4 ring[RS];
#pragma unroll
  for (int q = 0; q < RS; ++q) ring[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int r0 = 0; r0 < NR; r0 += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = r0 + u;
      if (NR % U != 0 && r >= NR) break;
      const int slot_new = SHIFT ? KS - 1 + u : u;
      issue_rows(r + GD - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(GD - 1) : "memory");
      ring[slot_new] = live ? gring[r & (GD - 1)][threadIdx.x] : make_float4(0.f, 0.f, 0.f, 0.f);
      {                                            // E, rows: doubled when the row enters the window
        const int ry = reflect_idx(y0 - P + r, h);
        if (ry == 0 || ry == h - 1) { float4& b = ring[slot_new]; b.x *= 2.f; b.y *= 2.f; b.z *= 2.f; b.w *= 2.f; }
      }
      if (r < KS - 1) continue;
      const int yo = y0 + r - (KS - 1);            // output row of this iteration
#define FAVAE_WIN(t) ring[SHIFT ? u + (t) : (u + 1 + (t)) % RS]
      // ---- vertical pass: the mirrored pre-adds are shared by the two sides
      float2 s01[P > 0 ? P : 1], s23[P > 0 ? P : 1];
#pragma unroll
      for (int t = 0; t < P; ++t) {
        const float4& a = FAVAE_WIN(t);
        const float4& b = FAVAE_WIN(KS - 1 - t);
        s01[t] = pk_add(make_float2(a.x, a.y), make_float2(b.x, b.y));
        s23[t] = pk_add(make_float2(a.z, a.w), make_float2(b.z, b.w));
      }
      const float4 mid = FAVAE_WIN(P);
#undef FAVAE_WIN
      constexpr int LP = 2, NB = (P + 3) / 4;
      float2 acc[2][4];
#pragma unroll
      for (int sd = 0; sd < 2; ++sd) {
        acc[sd][0] = pk_mul(pk_dup(mid.x), kd[sd][P]); acc[sd][1] = pk_mul(pk_dup(mid.y), kd[sd][P]);
        acc[sd][2] = pk_mul(pk_dup(mid.z), kd[sd][P]); acc[sd][3] = pk_mul(pk_dup(mid.w), kd[sd][P]);
#pragma unroll
        for (int t = 0; t < P; ++t) {
          acc[sd][0] = pk_fma(pk_dup(s01[t].x), kd[sd][t], acc[sd][0]); acc[sd][1] = pk_fma(pk_dup(s01[t].y), kd[sd][t], acc[sd][1]);
          acc[sd][2] = pk_fma(pk_dup(s23[t].x), kd[sd][t], acc[sd][2]); acc[sd][3] = pk_fma(pk_dup(s23[t].y), kd[sd][t], acc[sd][3]);
        }
        acc[sd][0] = pk_mul(acc[sd][0], pk_dup(e0)); acc[sd][3] = pk_mul(acc[sd][3], pk_dup(e3));
        // ---- shared line of (v, v') pairs of this side: two planes of float4 (see blur_adjsig_kernel)
        float4* planeA = reinterpret_cast<float4*>(lines2 + (size_t)((sd * 2 + (r & 1)) * groups + grp) * ll);
        float4* planeB = planeA + (tpi + 2 * LP);
        planeA[LP + tx] = make_float4(acc[sd][0].x, acc[sd][0].y, acc[sd][1].x, acc[sd][1].y);
        planeB[LP + tx] = make_float4(acc[sd][2].x, acc[sd][2].y, acc[sd][3].x, acc[sd][3].y);
        halo_puts<P>(planeA + LP + tx, planeB + LP + tx, tx, tpi, acc[sd]);
      }
      __syncthreads();
#pragma unroll
      for (int sd = 0; sd < 2; ++sd) {
        float4* planeA = reinterpret_cast<float4*>(lines2 + (size_t)((sd * 2 + (r & 1)) * groups + grp) * ll);
        float4* planeB = planeA + (tpi + 2 * LP);
        float2 cols[4 * (2 * NB + 1)];             // columns x0 - 4 NB .. x0 + 4 NB + 3; own from registers
#pragma unroll
        for (int q = -NB; q <= NB; ++q) {
          float2* d = cols + 4 * (q + NB);
          if (q == 0) { d[0] = acc[sd][0]; d[1] = acc[sd][1]; d[2] = acc[sd][2]; d[3] = acc[sd][3]; }
          else {
            const float4 a = planeA[LP + tx + q], b = planeB[LP + tx + q];
            d[0] = make_float2(a.x, a.y); d[1] = make_float2(a.z, a.w);
            d[2] = make_float2(b.x, b.y); d[3] = make_float2(b.z, b.w);
          }
        }
        const float2* seg = cols + (4 * NB - P);   // seg[i] = column x0 - P + i
        float o[4], z[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float2 a2 = pk_mul(pk_dup(kd[sd][P].x), seg[c + P]);     // (sum k v, sum k v')
          float za = kd[sd][P].y * seg[c + P].x;                    // sum k' v
#pragma unroll
          for (int t = 0; t < P; ++t) {
            const float2 q = pk_add(seg[c + t], seg[c + KS - 1 - t]);
            a2 = pk_fma(pk_dup(kd[sd][t].x), q, a2);
            za = fmaf(kd[sd][t].y, q.x, za);
          }
          o[c] = a2.x; z[c] = a2.y + za;
        }
        {
          const float dr = ((yo == 0 || yo == h - 1) ? 0.5f : 1.f) * oscale[sd];   // D, rows, and the side's scale
          o[0] *= dr * d0; o[1] *= dr; o[2] *= dr; o[3] *= dr * d3;
          z[0] *= dr * d0; z[1] *= dr; z[2] *= dr; z[3] *= dr * d3;
        }
        if (live && yo < h) {
          *reinterpret_cast<float4*>(gout[sd] + (long long)yo * w + x0) = make_float4(o[0], o[1], o[2], o[3]);
          const float4 xrow = xring[sd][r & (GD - 1)][threadIdx.x];
          acc_sigma[sd] = fmaf(xrow.x, z[0], fmaf(xrow.y, z[1], fmaf(xrow.z, z[2], fmaf(xrow.w, z[3], acc_sigma[sd]))));
        }
      }
    }
    if constexpr (SHIFT) {
#pragma unroll
      for (int q = 0; q < KS - 1; ++q) ring[q] = ring[q + U];
    }
  }
#pragma unroll
  for (int sd = 0; sd < 2; ++sd) {
    const float v = warp_sum(acc_sigma[sd]);
    if ((threadIdx.x & 31) == 0) wred[sd][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) sm += wred[threadIdx.x][i];
    partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = sm;
  }
}

// MODE_SIGMA with the same packed arithmetic: d/dsigma <gy, H V x> = <gy, (H'V + HV') x> on the forward
// data path (reflect halos only, no border corrections, nothing stored): x goes through the
// (V x, V' x) / (H ., H' .) pair pipeline of blur_adjsig_kernel and each finished row is dotted
// with the matching gy row.  Together with MODE_ADJ it is the FAVAE_BLUR_SIGMA=split alternative to
// the fused adjoint + sigma kernel (two lean kernels, 16 instead of 12 B/element).
// Measured on B200 (4096 maps of 256^2, k = 9, together with the plain adjoint): rolled 15-row unroll
// at 154 registers 1.03 ms; capped at 128 registers (4 CTAs / SM) 0.85 ms; all rows unrolled 0.95 ms;
// both 0.69 ms.
template <int KS, int TH>
__global__ void __launch_bounds__(THREADS, FAVAE_SIGMA_MINB)
blur_sigma_kernel(const float* __restrict__ src, const float* __restrict__ aux, int h, int w, long long items,
                  int strips, const float* __restrict__ sigma, float* __restrict__ partials) {
  constexpr int P = KS / 2;
  extern __shared__ float lines[];               // as float2: [2 buffers][groups][w + 2*LPAD]
  __shared__ float sk[32], sdk[32];
  __shared__ float wred[THREADS / 32];
  float k[KS], dk[KS];
  load_weights<KS>(sigma, sk, sdk, k, dk);
  float2 kd[P + 1];
#pragma unroll
  for (int t = 0; t <= P; ++t) kd[t] = make_float2(k[t], dk[t]);

  const int tpi = w >> 2;
  const int groups = THREADS / tpi;
  const int grp = threadIdx.x / tpi, tx = threadIdx.x % tpi;
  const int x0 = tx * 4;
  const int ll = w + 2 * LPAD;
  float2* lines2 = reinterpret_cast<float2*>(lines);
  const long long item = (long long)blockIdx.x * groups + grp;
  const bool live = item < items;
  const long long map = live ? item / strips : 0;
  const int y0 = live ? (int)(item % strips) * TH : 0;
  const float* base = src + map * (long long)h * w;
  float acc_sigma = 0.f;

  constexpr int Q = 6 + ((3 - (KS + 6) % 3) % 3), RS = KS + Q, XQ = 3, NR = TH + KS - 1;
  static_assert(RS % XQ == 0, "gy prefetch ring must tile the unroll factor");
  const long long mapoff = map * (long long)h * w;
  auto load_row = [&](int r) -> float4 {
    const int yi = y0 - P + r;
    float4 in = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) in = ld4(base + (long long)reflect_idx(yi, h) * w + x0);
    return in;
  };
  auto load_g = [&](int r) -> float4 {           // gy row of the output produced at iteration r
    const int yo = y0 + r - (KS - 1);
    float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live && r >= KS - 1 && yo < h) gv = ld4(aux + mapoff + (long long)yo * w + x0);
    return gv;
  };
  float4 ring[RS], gq[XQ];
#pragma unroll
  for (int q = 0; q < RS; ++q) ring[q] = (q < Q) ? load_row(q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < XQ; ++q) gq[q] = load_g(q);
  constexpr int STEP = FAVAE_SIGMA_FULL ? NR : RS;
#pragma unroll 1
  for (int r0 = 0; r0 < NR; r0 += STEP) {
#pragma unroll
    for (int uu = 0; uu < STEP; ++uu) {
      const int r = r0 + uu;
      const int u = STEP == RS ? uu : uu % RS;
      if (r >= NR) break;
      if (r + Q < NR) ring[(u + Q) % RS] = load_row(r + Q);
      const float4 grow = gq[u % XQ];            // zero outside the map: those rows add nothing
      if (r + XQ < NR) gq[u % XQ] = load_g(r + XQ);
      if (r < KS - 1) continue;
#define FAVAE_WIN(t) ring[(u + RS - (KS - 1) + (t)) % RS]
      float2 acc[4];
      {
        const float4& m = FAVAE_WIN(P);
        acc[0] = pk_mul(pk_dup(m.x), kd[P]); acc[1] = pk_mul(pk_dup(m.y), kd[P]);
        acc[2] = pk_mul(pk_dup(m.z), kd[P]); acc[3] = pk_mul(pk_dup(m.w), kd[P]);
#pragma unroll
        for (int t = 0; t < P; ++t) {
          const float4& a = FAVAE_WIN(t);
          const float4& b = FAVAE_WIN(KS - 1 - t);
          const float2 s01 = pk_add(make_float2(a.x, a.y), make_float2(b.x, b.y));
          const float2 s23 = pk_add(make_float2(a.z, a.w), make_float2(b.z, b.w));
          acc[0] = pk_fma(pk_dup(s01.x), kd[t], acc[0]); acc[1] = pk_fma(pk_dup(s01.y), kd[t], acc[1]);
          acc[2] = pk_fma(pk_dup(s23.x), kd[t], acc[2]); acc[3] = pk_fma(pk_dup(s23.y), kd[t], acc[3]);
        }
      }
#undef FAVAE_WIN
      // Shared line in two planes of float4: plane A holds the pairs of columns 4q, 4q+1 of thread q,
      // plane B those of columns 4q+2, 4q+3, LP halo slots on either side.  Every 128-bit access of a
      // quarter warp then covers 128 contiguous bytes (the interleaved single line had a 32-byte thread
      // stride: two-way bank conflicts on every access, 67 % shared-pipe utilisation).
      constexpr int LP = 2, NB = (P + 3) / 4;
      float4* planeA = reinterpret_cast<float4*>(lines2 + (size_t)((r & 1) * groups + grp) * ll);
      float4* planeB = planeA + (tpi + 2 * LP);
      planeA[LP + tx] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
      planeB[LP + tx] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
      halo_puts<P>(planeA + LP + tx, planeB + LP + tx, tx, tpi, acc);
      __syncthreads();
      float2 cols[4 * (2 * NB + 1)];               // columns x0 - 4 NB .. x0 + 4 NB + 3; own from registers
#pragma unroll
      for (int q = -NB; q <= NB; ++q) {
        float2* d = cols + 4 * (q + NB);
        if (q == 0) { d[0] = acc[0]; d[1] = acc[1]; d[2] = acc[2]; d[3] = acc[3]; }
        else {
          const float4 a = planeA[LP + tx + q], b = planeB[LP + tx + q];
          d[0] = make_float2(a.x, a.y); d[1] = make_float2(a.z, a.w);
          d[2] = make_float2(b.x, b.y); d[3] = make_float2(b.z, b.w);
        }
      }
      const float2* seg = cols + (4 * NB - P);     // seg[i] = column x0 - P + i
      const float gv[4] = {grow.x, grow.y, grow.z, grow.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // z = sum_t k'[t] v + sum_t k[t] v': the pair (k'[t], k[t]) against the pair (v, v'), summed over lanes
        float2 a2 = pk_mul(make_float2(dk[P], k[P]), seg[c + P]);
#pragma unroll
        for (int t = 0; t < P; ++t) a2 = pk_fma(make_float2(dk[t], k[t]), pk_add(seg[c + t], seg[c + KS - 1 - t]), a2);
        acc_sigma = fmaf(gv[c], a2.x + a2.y, acc_sigma);
      }
    }
  }
  acc_sigma = warp_sum(acc_sigma);
  if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = acc_sigma;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) s += wred[i];
    partials[blockIdx.x] = s;
  }
}

inline bool supported(int h, int w, int ks) {
  const bool pow2 = w >= 4 && w <= 512 && (w & (w - 1)) == 0;
  const bool kok = ks == 3 || ks == 5 || ks == 9 || ks == 11 || ks == 15;
  // w >= 8 keeps the two border-owning threads distinct; halo slots need ks/2 <= LPAD
  return pow2 && w >= 8 && kok && ks / 2 < h && ks / 2 < w && ks / 2 <= LPAD;
}
// Strip height.  The kernels with a fully unrolled row loop use 32 (16 for small maps); the fused
// adjoint + sigma kernel (rolled loop, so the height is free) takes taller strips on large maps: a
// strip re-reads KS - 1 halo rows, 25 % of its loads at 32 rows, 12.5 % at 64 (1.10 -> 0.79 ms together
// with the occupancy change above).
inline int strip_rows(int h, int mode = MODE_FWD) {
  if (mode == MODE_PAIR) {
    if (FAVAE_PAIR_TH > FAVAE_ADJSIG_TH && h >= 2 * FAVAE_PAIR_TH) return FAVAE_PAIR_TH;
    mode = MODE_ADJ_SIG;
  }
  if (mode == MODE_ADJ_SIG && h >= 2 * FAVAE_ADJSIG_TH) return FAVAE_ADJSIG_TH;
  if (mode == MODE_ADJ && h >= 2 * FAVAE_ADJ_TH) return FAVAE_ADJ_TH;
  if (mode == MODE_FWD && FAVAE_FWD_TH > 32 && h >= 2 * FAVAE_FWD_TH) return FAVAE_FWD_TH;
  return h <= 16 ? 16 : 32;
}
inline long long num_blocks(long long maps, int h, int w, int mode = MODE_FWD) {
  const int th = strip_rows(h, mode), strips = (h + th - 1) / th, groups = THREADS / (w / 4);
  return (maps * strips + groups - 1) / groups;
}

template <int KS, int TH, int MODE>
static int launch_one(const float* src, const float* aux, long long maps, int h, int w, const float* sigma,
                      float* dst, float* partials, cudaStream_t s, float oscale, const float* oscale_dev) {
  const int strips = (h + TH - 1) / TH, groups = THREADS / (w / 4);
  const long long items = maps * strips;
  const long long blocks = (items + groups - 1) / groups;
  const size_t smem = sizeof(float) * 2 * groups * ((MODE == MODE_ADJ_SIG || MODE == MODE_SIGMA) ? 2 : 1) * (size_t)(w + 2 * LPAD);
  if constexpr (MODE == MODE_ADJ_SIG)
    blur_adjsig_kernel<KS, TH><<<(unsigned)blocks, THREADS, smem, s>>>(src, aux, h, w, items, strips, sigma, dst, partials, oscale, oscale_dev);
  else if constexpr (MODE == MODE_SIGMA)
    blur_sigma_kernel<KS, TH><<<(unsigned)blocks, THREADS, smem, s>>>(src, aux, h, w, items, strips, sigma, partials);
  else
    blur_fast_kernel<KS, TH, MODE><<<(unsigned)blocks, THREADS, smem, s>>>(src, aux, h, w, items, strips, sigma, dst,
                                                                        partials, oscale, oscale_dev);
  return check_launch("blur_fast");
}

template <int KS, int TH>
static int launch_diff_one(const float* enc, const float* dec, long long maps, int h, int w, const float* sigma_enc,
                           const float* sigma_dec, float* dst, cudaStream_t s) {
  const int strips = (h + TH - 1) / TH, groups = THREADS / (w / 4);
  const long long items = maps * strips;
  const long long blocks = (items + groups - 1) / groups;
  const size_t smem = sizeof(float) * 2 * groups * 2 * (size_t)(w + 2 * LPAD);      // lines of (enc, dec) pairs
  blur_diff_kernel<KS, TH><<<(unsigned)blocks, THREADS, smem, s>>>(enc, dec, h, w, items, strips, sigma_enc, sigma_dec, dst);
  return check_launch("blur_diff");
}
static int launch_diff(const float* enc, const float* dec, long long maps, int h, int w, int ks, const float* sigma_enc,
                       const float* sigma_dec, float* dst, cudaStream_t s) {
#define FAVAE_BLUR_CASE(KS)                                                                                   \
  case KS:                                                                                                    \
    return h <= 16 ? launch_diff_one<KS, 16>(enc, dec, maps, h, w, sigma_enc, sigma_dec, dst, s)               \
         : h >= 128 ? launch_diff_one<KS, 64>(enc, dec, maps, h, w, sigma_enc, sigma_dec, dst, s)              \
                    : launch_diff_one<KS, 32>(enc, dec, maps, h, w, sigma_enc, sigma_dec, dst, s);
  switch (ks) {
    FAVAE_BLUR_CASE(3)
    FAVAE_BLUR_CASE(5)
    FAVAE_BLUR_CASE(9)
    FAVAE_BLUR_CASE(11)
    FAVAE_BLUR_CASE(15)
  }
#undef FAVAE_BLUR_CASE
  return fail(-22, "favae_b200: %s", "blur_diff: unsupported kernel size");
}

template <int KS, int TH>
static int launch_pair_one(const float* gy, const float* xe, const float* xd, long long maps, int h, int w,
                           const float* se, const float* sd, float* ge, float* gd, float* partials,
                           const float* scale_dev, cudaStream_t s) {
  const int strips = (h + TH - 1) / TH, groups = THREADS / (w / 4);
  const long long items = maps * strips;
  const long long blocks = (items + groups - 1) / groups;
  const size_t smem = sizeof(float) * 2 * 2 * 2 * groups * (size_t)(w + 2 * LPAD);     // [2 sides][2 buffers] pair lines
  // 25 KB of static rings + up to 32 KB of lines (narrow maps: 32 items per CTA) pass the 48 KB default
  static PerDevice<size_t> configured_dev;
  size_t& configured = configured_dev.here();
  if (smem > configured) {
    FAVAE_CUDA_OK(cudaFuncSetAttribute(blur_adjsig_pair_kernel<KS, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  blur_adjsig_pair_kernel<KS, TH><<<(unsigned)blocks, THREADS, smem, s>>>(gy, xe, xd, h, w, items, strips, se, sd, ge, gd,
                                                                         partials, scale_dev);
  return check_launch("blur_adjsig_pair");
}
static int launch_pair(const float* gy, const float* xe, const float* xd, long long maps, int h, int w, int ks,
                       const float* se, const float* sd, float* ge, float* gd, float* partials,
                       const float* scale_dev, cudaStream_t s) {
#define FAVAE_BLUR_CASE(KS)                                                                                        \
  case KS:                                                                                                         \
    if (FAVAE_PAIR_TH > FAVAE_ADJSIG_TH && strip_rows(h, MODE_PAIR) == FAVAE_PAIR_TH)                              \
      return launch_pair_one<KS, FAVAE_PAIR_TH>(gy, xe, xd, maps, h, w, se, sd, ge, gd, partials, scale_dev, s);   \
    return strip_rows(h, MODE_PAIR) == FAVAE_ADJSIG_TH                                                             \
               ? launch_pair_one<KS, FAVAE_ADJSIG_TH>(gy, xe, xd, maps, h, w, se, sd, ge, gd, partials, scale_dev, s) \
           : strip_rows(h, MODE_PAIR) == 16                                                                        \
               ? launch_pair_one<KS, 16>(gy, xe, xd, maps, h, w, se, sd, ge, gd, partials, scale_dev, s)            \
               : launch_pair_one<KS, 32>(gy, xe, xd, maps, h, w, se, sd, ge, gd, partials, scale_dev, s);
  switch (ks) {
    FAVAE_BLUR_CASE(3)
    FAVAE_BLUR_CASE(5)
    FAVAE_BLUR_CASE(9)
    FAVAE_BLUR_CASE(11)
    FAVAE_BLUR_CASE(15)
  }
#undef FAVAE_BLUR_CASE
  return fail(-22, "favae_b200: %s", "blur_adjsig_pair: unsupported kernel size");
}

template <int MODE>
static int launch(const float* src, const float* aux, long long maps, int h, int w, int ks, const float* sigma,
                  float* dst, float* partials, cudaStream_t s, float oscale = 1.0f, const float* oscale_dev = nullptr) {
#define FAVAE_BLUR_CASE(KS)                                                                         \
  case KS:                                                                                          \
    if (MODE == MODE_ADJ_SIG && strip_rows(h, MODE) == FAVAE_ADJSIG_TH)                             \
      return launch_one<KS, (MODE == MODE_ADJ_SIG ? FAVAE_ADJSIG_TH : 32), MODE>(src, aux, maps, h, w, sigma, dst, partials, s, oscale, oscale_dev); \
    if (MODE == MODE_FWD && FAVAE_FWD_TH > 32 && strip_rows(h, MODE) == FAVAE_FWD_TH)               \
      return launch_one<KS, (MODE == MODE_FWD ? FAVAE_FWD_TH : 32), MODE>(src, aux, maps, h, w, sigma, dst, partials, s, oscale, oscale_dev); \
    if (MODE == MODE_ADJ && strip_rows(h, MODE) == FAVAE_ADJ_TH)                                    \
      return launch_one<KS, (MODE == MODE_ADJ ? FAVAE_ADJ_TH : 32), MODE>(src, aux, maps, h, w, sigma, dst, partials, s, oscale, oscale_dev); \
    return strip_rows(h, MODE) == 16 ? launch_one<KS, 16, MODE>(src, aux, maps, h, w, sigma, dst, partials, s, oscale, oscale_dev) \
                                     : launch_one<KS, 32, MODE>(src, aux, maps, h, w, sigma, dst, partials, s, oscale, oscale_dev);
  switch (ks) {
    FAVAE_BLUR_CASE(3)
    FAVAE_BLUR_CASE(5)
    FAVAE_BLUR_CASE(9)
    FAVAE_BLUR_CASE(11)
    FAVAE_BLUR_CASE(15)
  }
#undef FAVAE_BLUR_CASE
  return fail(-22, "favae_b200: %s", "blur_fast: unsupported kernel size");
}

}  // namespace blurf
}  // namespace favae
