// Tensor-core nearest-code search for sm_100a: TMA -> shared memory -> tcgen05.mma (fp16 in,
// fp32 accumulate in TMEM) -> tcgen05.ld epilogue that keeps, per latent, every code whose
// approximate similarity is within a proven error bound of the running maximum.  The N x K
// similarity matrix never exists in HBM (reference: einsum + argmax,
// /root/reference/models/l2_quantize.py:410-411, materialises it twice).
//
// Exactness.  Latents and codes are L2-normalised in fp32, scaled by 16 and rounded to fp16
// (favae_vq_prepare_rows).  With unit roundoff u = 2^-11 the error of one approximate
// similarity is bounded by (2u + u^2)|x||e| plus the fp32 accumulation error, < 1.1e-3, so the
// true fp32 arg-max always lies within TAU = 2.5e-3 of the approximate maximum.  The epilogue
// works on 32-code chunks: every chunk whose maximum falls inside that band is recorded, and all
// codes of a recorded chunk that were inside the band when it was recorded (a 32-bit mask) are
// re-scored with exact fp32 dot products (vq_rescore_kernel, same accumulation order as
// favae_vq_search_exact); ties go to the lowest index.  Rows whose candidate list overflows (pathological ties, all-zero latents)
// are searched exhaustively in fp32 (vq_fallback_kernel).  The result contract is therefore the
// same as favae_vq_search_exact.
//
// L2 traffic.  Streaming the codebook to every CTA needs 64 B/cycle/SM at full MMA rate -- more
// than the ~6300 B/cycle the L2 can deliver to 148 SMs (measured: the single-CTA variant sits at
// exactly that ceiling, 65 % tensor-active; TMA multicast inside a 2-CTA cluster does not reduce
// L2 reads on this part).  The default variant therefore pairs two CTAs (cta_group::2): the pair
// computes a 256-latent x 256-code tile, each CTA stages its own 128 latents and HALF of every
// code block, the leader issues tcgen05.mma.cta_group::2 (M = 256) which reads both halves, and
// each CTA's TMEM receives its own 128 rows.  L2 reads per MMA are halved.
//
// Work decomposition.  A work item is (128-latent tile m, 256-code tile c).  Items are
// linearised m-major and cut into equal contiguous ranges, one per CTA (persistent, 1 CTA/SM),
// so every SM is busy even when there are fewer latent tiles than SMs.  A CTA's range is a short
// list of segments (m, c_begin..c_end); each segment reports a per-row (max, candidates) record
// into its own slot; the re-score kernel merges the slots of a row.
//
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-11 epilogue: one TMEM lane = one latent per thread, two warps per lane quarter, each
// taking one 128-column half of the accumulator (two warps per scheduler hide TMEM latency).  Pipelines: 4-stage smem ring
// (TMA <-> MMA) and a double-buffered 2 x 256-column TMEM accumulator (MMA <-> epilogue).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace favae {
namespace tc {

constexpr int BM = 128;            // latents per tile (UMMA M)
constexpr int BN = 256;            // codes per tile (UMMA N)
constexpr int BK = 64;             // fp16 elements per 128-byte swizzle row
constexpr int UK = 16;             // UMMA K for 16-bit inputs
constexpr int STAGES = 4;
constexpr int CAP = 8;             // candidate chunk records per latent, segment and column half
constexpr int MAX_KB = 4;          // d <= 256
constexpr int THREADS = 384;
constexpr int EPI_WARPS = 8;
constexpr float HALF_SCALE = 16.0f;                 // applied by prepare_rows to xh / eh
constexpr float TAU = 2.5e-3f * HALF_SCALE * HALF_SCALE;   // band in units of the scaled product
constexpr uint32_t A_KB_BYTES = BM * BK * 2;        // 16 KB
constexpr uint32_t B_STAGE_BYTES = BN * BK * 2;     // 32 KB
// instruction descriptor: D=f32 (bit 4), A=B=f16 K-major, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);  // M = 256

struct Params {
  long long n, k;
  int d, kb;                       // kb = d / 64
  int m_tiles, code_tiles;
  int mc;                          // 1: CTA pairs (cta_group::2), work unit = pair of latent tiles
  long long pairs, per_cta;        // work items (m_unit, code tile) and items per CTA / cluster
  int slots;
  // one record per (latent, slot, column half): rec = (row * slots + slot) * 2 + half
  float* ws_max;                   // [recs]       running maximum (scaled units)
  int* ws_cnt;                     // [recs]       (-1 = overflow)
  unsigned int* ws_idx;            // [recs][CAP]  chunk index = code / 32
  unsigned int* ws_mask;           // [recs][CAP]  codes of the chunk inside the band when recorded
  float* ws_val;                   // [recs][CAP]  chunk maximum
  int* err;                        // device error word (pipeline timeout)
  int debug;                       // profiling experiments only (FAVAE_VQ_TC_DEBUG): 1 = epilogue skips TMEM
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a broken pipeline traps (error word set) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {     // ~2 s
      if (err) atomicExch(err, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;        // clears the CTA-rank bit: address in the pair's leader
// 2-SM TMA load: data lands in THIS CTA's shared memory, the bytes are counted on the leader's barrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC2), "r"(accumulate) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one elected lane of a converged warp (the role loops run warp-uniform so that descriptors and
// barrier addresses stay in uniform registers; only the issue instructions are predicated)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max32(const float* v) {
  float t[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) t[j] = fmaxf(v[j], v[j + 16]);
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = fmaxf(t[j], t[j + 8]);
#pragma unroll
  for (int j = 0; j < 4; ++j) t[j] = fmaxf(t[j], t[j + 4]);
  return fmaxf(fmaxf(t[0], t[2]), fmaxf(t[1], t[3]));
}
__device__ __forceinline__ unsigned int band_mask32(const float* v, float thr) {
  unsigned int m = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) m |= (v[j] >= thr ? 1u : 0u) << j;
  return m;
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused with swizzle)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

struct Segment { int m, c_begin, c_end, slot; };

// next segment of this CTA's (cluster's) item range; returns false past the end.  s.m is the latent
// tile of THIS CTA: with multicast the item's unit is a tile pair and rank picks the tile.
__device__ __forceinline__ bool next_segment(const Params& p, long long unit_id, int rank, long long& pair,
                                             long long end, Segment& s) {
  if (pair >= end) return false;
  const int mu = (int)(pair / p.code_tiles);
  s.m = p.mc ? 2 * mu + rank : mu;
  s.c_begin = (int)(pair % p.code_tiles);
  const long long left = end - pair;
  s.c_end = (int)min((long long)p.code_tiles, (long long)s.c_begin + left);
  const long long first = ((long long)mu * p.code_tiles) / p.per_cta;
  s.slot = (int)(unit_id - first);
  pair += s.c_end - s.c_begin;
  return true;
}

template <bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
vq_search_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const Params p) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte-swizzle tiles need 1024-byte alignment
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // layout: A (MAX_KB x 16 KB) | B ring (128 KB: 4 x 32 KB, or 8 x 16 KB halves for CTA pairs) |
  //         candidate lists | barriers
  constexpr int NST = PAIR ? 2 * STAGES : STAGES;
  constexpr uint32_t STB = PAIR ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;
  unsigned char* a_base = smem;
  unsigned char* b_base = smem + MAX_KB * A_KB_BYTES;
  unsigned int* cand_idx = reinterpret_cast<unsigned int*>(b_base + STAGES * B_STAGE_BYTES);
  float* cand_val = reinterpret_cast<float*>(cand_idx + 2 * BM * CAP);
  unsigned int* cand_mask = reinterpret_cast<unsigned int*>(cand_val + 2 * BM * CAP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(cand_mask + 2 * BM * CAP);
  // bars: full[8], empty[8], a_full, a_empty, tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (8 + s); };
  const uint32_t A_FULL = bar0 + 8u * 16, A_EMPTY = bar0 + 8u * 17;
  auto T_FULL = [&](int s) { return bar0 + 8u * (18 + s); };
  auto T_EMPTY = [&](int s) { return bar0 + 8u * (20 + s); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    mbar_init(A_FULL, 1); mbar_init(A_EMPTY, 1);
    // pair: the leader's tmem_empty collects the epilogue warps of both CTAs
    for (int s = 0; s < 2; ++s) { mbar_init(T_FULL(s), 1); mbar_init(T_EMPTY(s), PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                 // both CTAs' barriers exist before anything is signalled
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int rank = PAIR ? (int)cluster_rank() : 0;
  const bool leader = rank == 0;
  const long long unit_id = PAIR ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;
  const long long pair_begin = unit_id * p.per_cta;
  const long long pair_end = min(p.pairs, pair_begin + p.per_cta);

  if (warp == 0) {
    // ================= TMA producer (warp-uniform loop, one elected lane issues) =================
    int stage = 0;
    uint32_t phase = 0, a_phase = 0;
    long long pair = pair_begin;
    Segment s;
    while (next_segment(p, unit_id, rank, pair, pair_end, s)) {
      mbar_wait(A_EMPTY, a_phase ^ 1, p.err, 1);
      if (elect_one()) {
        if (PAIR) {
          if (leader) mbar_expect_tx(A_FULL, 2u * (uint32_t)p.kb * A_KB_BYTES);   // both CTAs' latent tiles
          for (int kb = 0; kb < p.kb; ++kb)
            tma_load_2d_2sm(smem_u32(a_base + kb * A_KB_BYTES), &map_a, A_FULL, kb * BK, s.m * BM);
        } else {
          mbar_expect_tx(A_FULL, (uint32_t)p.kb * A_KB_BYTES);
          for (int kb = 0; kb < p.kb; ++kb)
            tma_load_2d(smem_u32(a_base + kb * A_KB_BYTES), &map_a, A_FULL, kb * BK, s.m * BM);
        }
      }
      __syncwarp();
      a_phase ^= 1;
      for (int c = s.c_begin; c < s.c_end; ++c) {
        for (int kb = 0; kb < p.kb; ++kb) {
          mbar_wait(EMPTY(stage), phase ^ 1, p.err, 2);
          if (elect_one()) {
            if (PAIR) {   // my half of the code block; the leader's barrier counts both halves
              if (leader) mbar_expect_tx(FULL(stage), B_STAGE_BYTES);
              tma_load_2d_2sm(smem_u32(b_base + stage * STB), &map_b, FULL(stage), kb * BK, c * BN + rank * (BN / 2));
            } else {
              mbar_expect_tx(FULL(stage), B_STAGE_BYTES);
              tma_load_2d(smem_u32(b_base + stage * STB), &map_b, FULL(stage), kb * BK, c * BN);
            }
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ================= MMA issuer (the pair's leader issues for both CTAs) =================
    int stage = 0, acc = 0;
    uint32_t phase = 0, a_phase = 0, acc_phase = 0;
    long long pair = pair_begin;
    Segment s;
    while (next_segment(p, unit_id, rank, pair, pair_end, s)) {
      mbar_wait(A_FULL, a_phase, p.err, 3);
      a_phase ^= 1;
      for (int c = s.c_begin; c < s.c_end; ++c) {
        mbar_wait(T_EMPTY(acc), acc_phase ^ 1, p.err, 4);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * BN;
        for (int kb = 0; kb < p.kb; ++kb) {
          mbar_wait(FULL(stage), phase, p.err, 5);
          tc_fence_after();
          const uint64_t da = make_desc(smem_u32(a_base + kb * A_KB_BYTES));
          const uint64_t db = make_desc(smem_u32(b_base + stage * STB));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              // advance both descriptors by k * 32 bytes inside the 128-byte swizzle row
              if (PAIR) tc_mma_2sm(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), (kb | k) != 0 ? 1u : 0u);
              else tc_mma(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), (kb | k) != 0 ? 1u : 0u);
            }
            if (PAIR) tc_commit_2sm(EMPTY(stage), (uint16_t)3);   // free the slot in both CTAs
            else tc_commit(EMPTY(stage));         // smem slot free once these MMAs retire
            if (kb == p.kb - 1) {                 // accumulator complete: wake the epilogue(s)
              if (PAIR) tc_commit_2sm(T_FULL(acc), (uint16_t)3);
              else tc_commit(T_FULL(acc));
            }
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (elect_one()) {
        if (PAIR) tc_commit_2sm(A_EMPTY, (uint16_t)3);
        else tc_commit(A_EMPTY);                // latent tile no longer read
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================= epilogue: running max + candidate chunk band per latent =================
    const int q = warp & 3;                     // TMEM lane quarter == warp_id % 4
    const int half = (warp - 4) >> 2;           // which 128-column half of the accumulator
    const int row = q * 32 + lane;
    unsigned int* my_idx = cand_idx + (half * BM + row) * CAP;
    float* my_val = cand_val + (half * BM + row) * CAP;
    unsigned int* my_mask = cand_mask + (half * BM + row) * CAP;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long pair = pair_begin;
    Segment s;
    while (next_segment(p, unit_id, rank, pair, pair_end, s)) {
      const long long grow = (long long)s.m * BM + row;
      const bool active = grow < p.n;
      float run = -INFINITY;
      int cnt = 0;
      bool overflow = false;
      auto push = [&](unsigned int chunk, float val, float thr, unsigned int mask) {
        if (cnt == CAP) {                       // drop records that fell out of the band
          int keep = 0;
          for (int i = 0; i < CAP; ++i) {
            const float pv = my_val[i];
            if (pv >= thr) { my_val[keep] = pv; my_idx[keep] = my_idx[i]; my_mask[keep] = my_mask[i]; ++keep; }
          }
          cnt = keep;
        }
        if (cnt < CAP) { my_idx[cnt] = chunk; my_val[cnt] = val; my_mask[cnt] = mask; ++cnt; }
        else overflow = true;
      };
      for (int c = s.c_begin; c < s.c_end; ++c) {
        mbar_wait(T_FULL(acc), acc_phase, p.err, 6);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
        // pull this warp's 128 columns into registers and hand the accumulator back to the MMA
        // pipe BEFORE scanning them: the drain time of a TMEM buffer is on the critical path of the
        // two-buffer MMA <-> epilogue ring
        float v[2][64];
        if (p.debug & 1) {
#pragma unroll
          for (int i = 0; i < 64; ++i) { v[0][i] = -1.0e30f - 1.0e27f * (float)c - 1.0e24f * (float)i; v[1][i] = v[0][i] - 1.0e26f; }
        } else {
          tmem_ld64(taddr, v[0]);
          tmem_ld64(taddr + 64, v[1]);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_remote(T_EMPTY(acc), 0);        // the leader's MMA thread waits on it
          else mbar_arrive(T_EMPTY(acc));
        }
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const float ca = max32(v[it]), cb = max32(v[it] + 32);
          const float cm = fmaxf(ca, cb);
          const bool hit = active && !overflow && (cm >= run - TAU);
          if (__any_sync(0xffffffffu, hit)) {
            if (hit) {
              run = fmaxf(run, cm);
              const float thr = run - TAU;
              const unsigned int chunk0 = (unsigned int)c * (BN / 32) + half * (BN / 64) + it * 2;
              if (ca >= thr) push(chunk0, ca, thr, band_mask32(v[it], thr));
              if (cb >= thr) push(chunk0 + 1, cb, thr, band_mask32(v[it] + 32, thr));
            }
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      // publish this segment's record
      const long long o = (((long long)s.m * BM + row) * p.slots + s.slot) * 2 + half;
      p.ws_max[o] = active ? run : -INFINITY;
      p.ws_cnt[o] = active ? (overflow ? -1 : cnt) : 0;
      for (int i = 0; i < cnt; ++i) {
        p.ws_idx[o * CAP + i] = my_idx[i]; p.ws_val[o * CAP + i] = my_val[i]; p.ws_mask[o * CAP + i] = my_mask[i];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                 // no CTA leaves while the pair still uses its memory
  if (warp == 2) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + MAX_KB * A_KB_BYTES + STAGES * B_STAGE_BYTES +
                              2 * BM * CAP * 12 + 24 * 8 + 16;

// ---------------------------------------------------------------- merge + exact re-score
__device__ __forceinline__ unsigned int f_order(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int RESCORE_CAP = 96;      // candidate codes per latent handled in place; more -> fallback

__global__ void __launch_bounds__(256)
vq_rescore_kernel(const Params p, const float* __restrict__ xn, const float* __restrict__ en,
                  long long* __restrict__ idx, int* __restrict__ ovf_count, int* __restrict__ ovf_rows,
                  unsigned long long* __restrict__ keys) {
  __shared__ unsigned int cand[8][RESCORE_CAP];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
  if (row >= p.n) return;
  const int m = (int)(row / BM);
  const int mu = p.mc ? m >> 1 : m;                    // work unit (latent tile or tile pair)
  // records of the unit: slots c0..c1, two column halves each
  const long long c0 = ((long long)mu * p.code_tiles) / p.per_cta;
  const long long c1 = ((long long)(mu + 1) * p.code_tiles - 1) / p.per_cta;
  const int nrec = (int)(c1 - c0 + 1) * 2;
  const long long base = row * p.slots * 2;
  // Everything the first 32 candidate entries need is requested BEFORE the running maximum is known
  // (entries past a record's count hold stale values; the predicate below discards them): one global
  // round trip per latent instead of a chain of four dependent ones (max -> count -> value -> chunk /
  // mask), which was what this warp-per-latent kernel spent its time on -- most latents end with a
  // single candidate and never reach the dot products.
  const int nent = nrec * CAP;
  const long long ebase = base * CAP;
  float pv = 0.f;
  unsigned int pchunk = 0, pmask = 0;
  int pcnt = 0;
  if (lane < nent) {
    pv = p.ws_val[ebase + lane];
    pchunk = p.ws_idx[ebase + lane];
    pmask = p.ws_mask[ebase + lane];
    pcnt = p.ws_cnt[base + lane / CAP];
  }
  float gmax = -INFINITY;
  bool ovf = false;
  for (int s = lane; s < nrec; s += 32) {
    gmax = fmaxf(gmax, p.ws_max[base + s]);
    ovf |= p.ws_cnt[base + s] < 0;
  }
  gmax = warp_max(gmax);
  ovf = __any_sync(0xffffffffu, ovf);           // warp-uniform from here on
  const float thr = gmax - TAU;
  // gather the candidate codes of all records still inside the band
  int total = 0;
  for (int e0 = 0; e0 < nent && !ovf; e0 += 32) {
    const int e = e0 + lane;
    unsigned int chunk = 0, mask = 0;
    if (e0 == 0) {
      if (lane < nent && (lane % CAP) < pcnt && pv >= thr) { chunk = pchunk; mask = pmask; }
    } else if (e < nent) {
      const int s = e / CAP, i = e % CAP;
      if (i < p.ws_cnt[base + s] && p.ws_val[ebase + e] >= thr) {
        chunk = p.ws_idx[ebase + e];
        mask = p.ws_mask[ebase + e];
      }
    }
    const int mine = __popc(mask);
    int incl = mine;                              // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int pos = total + incl - mine;
    total += __shfl_sync(0xffffffffu, incl, 31);
    if (total > RESCORE_CAP) { ovf = true; break; }
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      cand[wib][pos++] = chunk * 32u + b;
    }
  }
  if (ovf) {
    if (lane == 0) {
      keys[row] = 0ull;                          // meeting point of the fallback kernel's code slices
      ovf_rows[atomicAdd(ovf_count, 1)] = (int)row;
    }
    return;
  }
  __syncwarp();
  if (total == 1) {                               // the common case: nothing to re-score
    if (lane == 0) idx[row] = (long long)cand[wib][0];
    return;
  }
  // Exact fp32 re-score, the whole warp on one candidate: lane l owns the float4 columns l, l + 32 of
  // the latent (registers) and of the code row (one coalesced 512-byte request per 128 columns), adds
  // its products in ascending column order and the 32 partial sums go through one fixed xor-shuffle
  // tree -- one accumulation order for every candidate, so identical codes still tie exactly and the
  // lowest index wins.  Four candidates are in flight at a time.  (First version: one lane per
  // candidate walking its whole row, i.e. 2-3 active lanes chasing 64 dependent 16-byte loads each:
  // latency bound, 212 us of the 1.55 ms search at N = 262144.)
  constexpr int XV = MAX_KB * BK / 128;                 // float4 per lane: 2 at d = 256
  const int nv = p.d >> 2;
  float4 xv[XV];
#pragma unroll
  for (int j = 0; j < XV; ++j) {
    const int c4 = lane + 32 * j;
    xv[j] = c4 < nv ? reinterpret_cast<const float4*>(xn + row * p.d)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  unsigned long long best = 0ull;
  for (int i0 = 0; i0 < total; i0 += 4) {
    float4 ev[4][XV];
    unsigned int codes[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      codes[q] = cand[wib][min(i0 + q, total - 1)];
      const float4* er = reinterpret_cast<const float4*>(en + (long long)codes[q] * p.d);
#pragma unroll
      for (int j = 0; j < XV; ++j) {
        const int c4 = lane + 32 * j;
        ev[q][j] = c4 < nv ? er[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < XV; ++j) {
        acc = fmaf(xv[j].x, ev[q][j].x, acc);
        acc = fmaf(xv[j].y, ev[q][j].y, acc);
        acc = fmaf(xv[j].z, ev[q][j].z, acc);
        acc = fmaf(xv[j].w, ev[q][j].w, acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const unsigned long long key = ((unsigned long long)f_order(acc) << 32) | (0xFFFFFFFFu - codes[q]);
      best = key > best ? key : best;            // a candidate repeated past the end changes nothing
    }
  }
  if (lane == 0) idx[row] = (long long)(0xFFFFFFFFu - (unsigned int)(best & 0xFFFFFFFFull));
}

// Exhaustive fp32 search of the (rare) rows whose candidate list overflowed.  A work unit is (row, slice
// of the codebook), with enough slices that a single overflowing latent still uses the whole GPU (one
// block per row took 1 ms for one row at K = 16384: every thread walked 64 code rows with 16-byte
// strided loads).  Within a unit each warp scores one code at a time with the accumulation order of
// vq_rescore_kernel (so ties between identical codes stay exact), the slices meet in a 64-bit
// atomicMax on keys[row] (zeroed by the re-score kernel when it queued the row), and the block that
// finishes the last unit decodes the winners.
__global__ void __launch_bounds__(256)
vq_fallback_kernel(const Params p, const float* __restrict__ xn, const float* __restrict__ en,
                   long long* __restrict__ idx, int* __restrict__ ovf_count, const int* __restrict__ ovf_rows,
                   unsigned long long* __restrict__ keys) {
  __shared__ unsigned long long wbest[8];
  __shared__ int last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int count = ovf_count[0];
  if (count == 0) return;
  int slices = (int)gridDim.x / count;
  slices = slices < 1 ? 1 : (slices > 64 ? 64 : slices);
  const long long per = (p.k + slices - 1) / slices;
  const long long units = (long long)count * slices;
  constexpr int XV = MAX_KB * BK / 128;
  const int nv = p.d >> 2;
  int done = 0;
  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const long long row = ovf_rows[u / slices];
    const long long c0 = (u % slices) * per, c1 = c0 + per < p.k ? c0 + per : p.k;
    float4 xv[XV];
#pragma unroll
    for (int j = 0; j < XV; ++j) {
      const int c4 = lane + 32 * j;
      xv[j] = c4 < nv ? reinterpret_cast<const float4*>(xn + row * p.d)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    unsigned long long best = 0ull;
    for (long long code = c0 + warp; code < c1; code += 8) {
      const float4* er = reinterpret_cast<const float4*>(en + code * p.d);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < XV; ++j) {
        const int c4 = lane + 32 * j;
        const float4 w = c4 < nv ? er[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
        acc = fmaf(xv[j].x, w.x, acc);
        acc = fmaf(xv[j].y, w.y, acc);
        acc = fmaf(xv[j].z, w.z, acc);
        acc = fmaf(xv[j].w, w.w, acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const unsigned long long key = ((unsigned long long)f_order(acc) << 32) | (0xFFFFFFFFu - (unsigned int)code);
      best = key > best ? key : best;
    }
    if (lane == 0) wbest[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long b = 0ull;
      for (int w = 0; w < 8; ++w) b = wbest[w] > b ? wbest[w] : b;
      atomicMax(&keys[row], b);
    }
    __syncthreads();
    ++done;
  }
  // the block that completes the last unit decodes every overflow row
  if (threadIdx.x == 0) {
    __threadfence();
    const long long before = atomicAdd(reinterpret_cast<unsigned long long*>(ovf_count + 2), (unsigned long long)done);
    last = (before + done == units) ? 1 : 0;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    for (int i = threadIdx.x; i < count; i += 256) {
      const long long row = ovf_rows[i];
      const unsigned long long b = *reinterpret_cast<volatile unsigned long long*>(&keys[row]);
      idx[row] = (long long)(0xFFFFFFFFu - (unsigned int)(b & 0xFFFFFFFFull));
    }
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

// TMA descriptors are built lazily and cached by (base pointer, rows, d, box): the codebook's fp16
// image and the per-shape latent scratch keep their addresses from call to call, so a training loop
// encodes each descriptor once.  A descriptor holds no data, only the address and
// the geometry, so a stale entry is harmless as long as the key matches.  Small direct-mapped table,
// one per thread (the ABI is called from one thread per device).
struct MapKey { const void* base; long long rows; int d, box; };
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
static int make_map_uncached(CUtensorMap* map, const void* base, long long rows, int d, int box_rows);
static int make_map(CUtensorMap* map, const void* base, long long rows, int d, int box_rows) {
  constexpr int SLOTS = 16;
  static thread_local MapSlot cache[SLOTS] = {};
  const size_t hsh = ((size_t)(uintptr_t)base >> 8) * 0x9E3779B97F4A7C15ull ^ (size_t)rows * 31u ^ (size_t)box_rows;
  MapSlot& sl = cache[(hsh >> 20) % SLOTS];
  if (sl.used && sl.key.base == base && sl.key.rows == rows && sl.key.d == d && sl.key.box == box_rows) {
    *map = sl.map;
    return 0;
  }
  int rc = make_map_uncached(map, base, rows, d, box_rows);
  if (rc) return rc;
  sl.key = MapKey{base, rows, d, box_rows};
  sl.map = *map;
  sl.used = true;
  return 0;
}
static int make_map_uncached(CUtensorMap* map, const void* base, long long rows, int d, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(-38, "favae_b200: %s", "cuTensorMapEncodeTiled is unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-22, "favae_b200: %s", "cuTensorMapEncodeTiled failed");
  return 0;
}

struct Plan {
  int m_tiles, code_tiles, grid, slots, mc;
  long long pairs, per_cta;
  size_t off_max, off_cnt, off_idx, off_mask, off_val, off_ovf, total;
};

static bool use_multicast() {
  static const bool off = [] { const char* e = getenv("FAVAE_VQ_TC"); return e && e[0] == 'n'; }();
  return !off;                                   // FAVAE_VQ_TC=nopair: single-CTA variant
}

static Plan make_plan(long long n, long long k) {
  Plan pl;
  pl.mc = use_multicast() ? 1 : 0;
  pl.m_tiles = (int)((n + BM - 1) / BM);
  pl.code_tiles = (int)(k / BN);
  const int units = pl.mc ? (pl.m_tiles + 1) / 2 : pl.m_tiles;
  pl.pairs = (long long)units * pl.code_tiles;
  long long g = pl.mc ? num_sms() / 2 : num_sms();          // clusters or CTAs
  if (g > pl.pairs) g = pl.pairs;
  if (g < 1) g = 1;
  pl.per_cta = (pl.pairs + g - 1) / g;
  pl.grid = (int)((pl.pairs + pl.per_cta - 1) / pl.per_cta) * (pl.mc ? 2 : 1);
  pl.slots = (int)((pl.code_tiles + pl.per_cta - 1) / pl.per_cta) + 1;
  const size_t rows = (size_t)units * (pl.mc ? 2 : 1) * BM;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o = 0;
  const size_t recs = rows * pl.slots * 2;
  pl.off_max = o; o = al(o + recs * sizeof(float));
  pl.off_cnt = o; o = al(o + recs * sizeof(int));
  pl.off_idx = o; o = al(o + recs * CAP * sizeof(unsigned int));
  pl.off_mask = o; o = al(o + recs * CAP * sizeof(unsigned int));
  pl.off_val = o; o = al(o + recs * CAP * sizeof(float));
  pl.off_ovf = o; o = al(o + 256 + (size_t)n * sizeof(int));
  pl.total = o;
  return pl;
}

}  // namespace tc
}  // namespace favae

using namespace favae;

extern "C" {

size_t favae_vq_search_tc_workspace_bytes(int64_t n, int64_t k, int d) {
  if (n <= 0 || k <= 0 || d <= 0 || d % tc::BK != 0 || d > tc::MAX_KB * tc::BK || k % tc::BN != 0 ||
      k >= 0x7FFFFFFFll)
    return 0;
  return tc::make_plan(n, k).total;
}

int favae_vq_search_tc_overflow_rows(const void* workspace, int64_t n, int64_t k, int d, int* count_host) {
  FAVAE_REQUIRE(workspace && count_host && favae_vq_search_tc_workspace_bytes(n, k, d) > 0,
                "vq_search_tc_overflow_rows: bad arguments");
  const tc::Plan pl = tc::make_plan(n, k);
  FAVAE_CUDA_OK(cudaMemcpy(count_host, (const unsigned char*)workspace + pl.off_ovf, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int favae_vq_search_tc(const void* xh, const void* eh, const float* xn, const float* en, int64_t n,
                       int64_t k, int d, void* workspace, size_t workspace_bytes, uint64_t* keys,
                       int64_t* idx, void* stream) {
  FAVAE_REQUIRE(xh && eh && xn && en && idx && workspace && keys, "vq_search_tc: null pointer");
  FAVAE_REQUIRE(n > 0 && favae_vq_search_tc_workspace_bytes(n, k, d) > 0,
                "vq_search_tc: needs d % 64 == 0, d <= 256, k % 256 == 0");
  const tc::Plan pl = tc::make_plan(n, k);
  FAVAE_REQUIRE(workspace_bytes >= pl.total, "vq_search_tc: workspace too small");
  FAVAE_REQUIRE(((uintptr_t)xh % 16 == 0) && ((uintptr_t)eh % 16 == 0) && ((uintptr_t)workspace % 256 == 0),
                "vq_search_tc: unaligned buffers");
  cudaStream_t s = (cudaStream_t)stream;
  CUtensorMap map_a, map_b;
  int rc = tc::make_map(&map_a, xh, n, d, tc::BM);
  if (rc) return rc;
  rc = tc::make_map(&map_b, eh, k, d, pl.mc ? tc::BN / 2 : tc::BN);   // CTA pair: each CTA loads half a block
  if (rc) return rc;

  unsigned char* ws = (unsigned char*)workspace;
  tc::Params p;
  p.n = n; p.k = k; p.d = d; p.kb = d / tc::BK;
  p.m_tiles = pl.m_tiles; p.code_tiles = pl.code_tiles; p.pairs = pl.pairs; p.per_cta = pl.per_cta;
  p.slots = pl.slots;
  p.ws_max = (float*)(ws + pl.off_max);
  p.ws_cnt = (int*)(ws + pl.off_cnt);
  p.ws_idx = (unsigned int*)(ws + pl.off_idx);
  p.ws_mask = (unsigned int*)(ws + pl.off_mask);
  p.ws_val = (float*)(ws + pl.off_val);
  int* ovf_count = (int*)(ws + pl.off_ovf);
  int* ovf_rows = ovf_count + 64;
  p.err = ovf_count + 1;
  FAVAE_CUDA_OK(cudaMemsetAsync(ovf_count, 0, 256, s));

  p.mc = pl.mc;
  { const char* e = getenv("FAVAE_VQ_TC_DEBUG"); p.debug = e ? atoi(e) : 0; }
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.here();
  if (!configured) {
    FAVAE_CUDA_OK(cudaFuncSetAttribute(tc::vq_search_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc::SMEM_BYTES));
    FAVAE_CUDA_OK(cudaFuncSetAttribute(tc::vq_search_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc::SMEM_BYTES));
    configured = true;
  }
  if (pl.mc) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pl.grid);
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = tc::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FAVAE_CUDA_OK(cudaLaunchKernelEx(&cfg, tc::vq_search_tc_kernel<true>, map_a, map_b, p));
  } else {
    tc::vq_search_tc_kernel<false><<<pl.grid, tc::THREADS, tc::SMEM_BYTES, s>>>(map_a, map_b, p);
  }
  rc = check_launch("vq_search_tc");
  if (rc) return rc;
  tc::vq_rescore_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(p, xn, en, (long long*)idx, ovf_count, ovf_rows,
                                                                 (unsigned long long*)keys);
  rc = check_launch("vq_rescore");
  if (rc) return rc;
  tc::vq_fallback_kernel<<<num_sms() * 2, 256, 0, s>>>(p, xn, en, (long long*)idx, ovf_count, ovf_rows,
                                                       (unsigned long long*)keys);
  return check_launch("vq_fallback");
}
}
