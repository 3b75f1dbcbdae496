// Vector-quantizer kernels (CUDA-core side): row preparation, exact fp32 search,
// gather / straight-through / commitment loss, code statistics, EMA codebook update,
// backward.  Each kernel cites the reference lines it replaces
// (/root/reference/models/l2_quantize.py).
#include <cuda_fp16.h>

#include "common.cuh"

namespace favae {

// element (n, c) of an (N x D) latent matrix stored row-major (hw == 1) or NCHW
__device__ __forceinline__ long long rows_addr(long long n, int c, int d, long long hw) {
  return (hw == 1) ? n * d + c : (n / hw) * (long long)d * hw + (long long)c * hw + n % hw;
}

// ----------------------------------------------------------------------------------
// l2norm (+ NCHW -> rows rearrange).  32 rows per block staged through shared memory so
// that both the NCHW read (contiguous in the position) and the row-major write
// (contiguous in the channel) are coalesced.     l2_quantize.py:24-25,403,408,540
// ----------------------------------------------------------------------------------
constexpr int PREP_ROWS = 32;
constexpr int PREP_THREADS = 256;

__global__ void __launch_bounds__(PREP_THREADS)
vq_prepare_rows_kernel(const float* __restrict__ x, long long n, int d, long long hw, int normalize,
                       float* __restrict__ xn, __half* __restrict__ xh, float* __restrict__ sq) {
  extern __shared__ float tile[];               // PREP_ROWS x (d + 1)
  const int ld = d + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long row0;
  int rows;
  if (hw == 1) {
    row0 = (long long)blockIdx.x * PREP_ROWS;
    rows = (int)min((long long)PREP_ROWS, n - row0);
    for (int r = warp; r < rows; r += PREP_THREADS / 32)
      for (int c = lane; c < d; c += 32) tile[r * ld + c] = x[(row0 + r) * d + c];
  } else {
    const long long tiles_per_img = (hw + PREP_ROWS - 1) / PREP_ROWS;
    const long long img = blockIdx.x / tiles_per_img, p0 = (blockIdx.x % tiles_per_img) * PREP_ROWS;
    row0 = img * hw + p0;
    rows = (int)min((long long)PREP_ROWS, hw - p0);
    const float* src = x + img * (long long)d * hw + p0;
    for (int c = warp; c < d; c += PREP_THREADS / 32)
      if (lane < rows) tile[lane * ld + c] = src[(long long)c * hw + lane];
  }
  __syncthreads();
  for (int r = warp; r < rows; r += PREP_THREADS / 32) {
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = tile[r * ld + c]; ss += v * v; }
    ss = warp_sum(ss);
    const float denom = normalize ? fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
    float ss_out = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float v = tile[r * ld + c] / denom;
      ss_out += v * v;
      if (xn) xn[(row0 + r) * d + c] = v;
      if (xh) xh[(row0 + r) * d + c] = __float2half_rn(v * 16.0f);   // see vq_search_tc.cu (HALF_SCALE)
    }
    if (sq) {
      ss_out = warp_sum(ss_out);
      if (lane == 0) sq[row0 + r] = ss_out;
    }
  }
}

// ----------------------------------------------------------------------------------
// exact fp32 search: 64 x 64 tiles, 4 x 4 micro-tiles, running packed (score, ~index) key
// per row; code range split across blockIdx.y, merged with a 64-bit atomicMax.
//                                                    l2_quantize.py:410-411 / :280-282
// ----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int float_order(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long pack_key(float score, unsigned int idx) {
  return ((unsigned long long)float_order(score) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

constexpr int SB = 64;      // tile side
constexpr int SK = 16;      // depth chunk

__global__ void __launch_bounds__(256)
vq_search_exact_kernel(const float* __restrict__ xn, const float* __restrict__ en,
                       const float* __restrict__ e_sq, long long n, long long k, int d, int metric,
                       long long codes_per_split, unsigned long long* __restrict__ keys) {
  __shared__ float As[SK][SB + 4];
  __shared__ float Bs[SK][SB + 4];
  __shared__ unsigned long long red[SB][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // tx: code quad, ty: row quad
  const long long row0 = (long long)blockIdx.x * SB;
  const long long c_begin = (long long)blockIdx.y * codes_per_split;
  const long long c_end = min(k, c_begin + codes_per_split);
  unsigned long long best[4] = {0ull, 0ull, 0ull, 0ull};

  const int lr = threadIdx.x >> 2, lc = (threadIdx.x & 3) * 4;   // loader: row lr, depth lc..lc+3
  for (long long c0 = c_begin; c0 < c_end; c0 += SB) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < d; k0 += SK) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kk = k0 + lc + q;
        const long long ra = row0 + lr, rb = c0 + lr;
        As[lc + q][lr] = (ra < n && kk < d) ? xn[ra * d + kk] : 0.f;
        Bs[lc + q][lr] = (rb < c_end && kk < d) ? en[rb * d + kk] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long code = c0 + tx * 4 + j;
      if (code < c_end) {
        const float esq = metric ? e_sq[code] : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float s = metric ? (2.0f * acc[i][j] - esq) : acc[i][j];
          const unsigned long long key = pack_key(s, (unsigned int)code);
          best[i] = key > best[i] ? key : best[i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) red[ty * 4 + i][tx] = best[i];
  __syncthreads();
  if (threadIdx.x < SB) {
    unsigned long long b = 0ull;
#pragma unroll
    for (int j = 0; j < 16; ++j) { const unsigned long long v = red[threadIdx.x][j]; b = v > b ? v : b; }
    const long long row = row0 + threadIdx.x;
    if (row < n) atomicMax(&keys[row], b);
  }
}

__global__ void vq_keys_to_idx_kernel(const unsigned long long* __restrict__ keys, long long n,
                                      long long* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (long long)(0xFFFFFFFFu - (unsigned int)(keys[i] & 0xFFFFFFFFull));
}

// ----------------------------------------------------------------------------------
// gather + straight-through + commitment-loss partials.   l2_quantize.py:415,554,560
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(PREP_THREADS)
vq_gather_st_kernel(const float* __restrict__ x, const float* __restrict__ embed,
                    const long long* __restrict__ idx, long long n, long long k, int d, long long hw,
                    int straight_through, float* __restrict__ out, float* __restrict__ partials) {
  extern __shared__ float tile[];               // PREP_ROWS x (d + 1) gathered code rows
  __shared__ float wsum[PREP_THREADS / 32];
  const int ld = d + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long row0, img = 0, p0 = 0;
  int rows;
  if (hw == 1) {
    row0 = (long long)blockIdx.x * PREP_ROWS;
    rows = (int)min((long long)PREP_ROWS, n - row0);
  } else {
    const long long tiles_per_img = (hw + PREP_ROWS - 1) / PREP_ROWS;
    img = blockIdx.x / tiles_per_img;
    p0 = (blockIdx.x % tiles_per_img) * PREP_ROWS;
    row0 = img * hw + p0;
    rows = (int)min((long long)PREP_ROWS, hw - p0);
  }
  float lsum = 0.f;
  if (hw == 1) {
    for (int r = warp; r < rows; r += PREP_THREADS / 32) {
      long long code = idx[row0 + r];
      code = code < 0 ? 0 : (code >= k ? k - 1 : code);
      const float* e = embed + code * d;
      for (int c = lane; c < d; c += 32) {
        const long long a = (row0 + r) * d + c;
        const float q = e[c];
        if (x) {
          const float xv = x[a];
          const float o = straight_through ? xv + (q - xv) : q;
          const float df = o - xv;
          lsum += df * df;
          out[a] = o;
        } else {
          out[a] = q;
        }
      }
    }
  } else {
    for (int r = warp; r < rows; r += PREP_THREADS / 32) {
      long long code = idx[row0 + r];
      code = code < 0 ? 0 : (code >= k ? k - 1 : code);
      const float* e = embed + code * d;
      for (int c = lane; c < d; c += 32) tile[r * ld + c] = e[c];
    }
    __syncthreads();
    const long long base = img * (long long)d * hw + p0;
    // four channels per warp and trip: their latent loads are in flight together (8 warps x 1 load per
    // SM-resident CTA left the kernel waiting on HBM latency: 24 us for 24 MB at N = 8192, D = 256)
    constexpr int CU = 4, WARPS = PREP_THREADS / 32;
    for (int c0 = warp; c0 < d; c0 += WARPS * CU) {
      float xv[CU];
#pragma unroll
      for (int j = 0; j < CU; ++j) {
        const int c = c0 + j * WARPS;
        xv[j] = (x && lane < rows && c < d) ? x[base + (long long)c * hw + lane] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < CU; ++j) {
        const int c = c0 + j * WARPS;
        if (lane < rows && c < d) {
          const long long a = base + (long long)c * hw + lane;
          const float q = tile[lane * ld + c];
          if (x) {
            const float o = straight_through ? xv[j] + (q - xv[j]) : q;
            const float df = o - xv[j];
            lsum += df * df;
            out[a] = o;
          } else {
            out[a] = q;
          }
        }
      }
    }
  }
  if (partials) {
    lsum = warp_sum(lsum);
    if (lane == 0) wsum[warp] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < PREP_THREADS / 32; ++i) s += wsum[i];
      partials[blockIdx.x] = s;
    }
  }
}

// ----------------------------------------------------------------------------------
// bins + embed_sum by scatter-add (one warp per latent).      l2_quantize.py:412,418,426
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vq_code_stats_kernel(const float* __restrict__ xn, const long long* __restrict__ idx, long long n,
                     long long k, int d, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  float* bins = stats;
  float* esum = stats + k;
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps) {
    long long code = idx[r];
    if (code < 0 || code >= k) continue;
    if (lane == 0) atomicAdd(&bins[code], 1.0f);
    const float* src = xn + r * d;
    float* dst = esum + code * d;
    for (int c = lane; c < d; c += 32) atomicAdd(&dst[c], src[c]);
  }
}

// Deterministic variant (FAVAE_VQ_DETERMINISTIC=1): one warp per CODE scans the index vector (L2
// resident: 8 bytes per latent) and adds the rows of its members in ascending latent order, so bins
// and embed_sum are bit-reproducible from run to run and independent of the launch geometry (the
// atomic kernel above adds in arrival order: ~1e-7 relative noise that the EMA feeds into the next
// search).  No zero-fill needed: every element of stats is written exactly once.
__global__ void __launch_bounds__(256)
vq_code_stats_det_kernel(const float* __restrict__ xn, const long long* __restrict__ idx, long long n,
                         long long k, int d, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long code = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (code >= k) return;
  constexpr int MAXV = 8;                            // d <= 256 in registers; larger d in column chunks
  for (int c0 = 0; c0 < d; c0 += 32 * MAXV) {
    float acc[MAXV];
#pragma unroll
    for (int j = 0; j < MAXV; ++j) acc[j] = 0.f;
    float count = 0.f;
    for (long long r0 = 0; r0 < n; r0 += 32) {
      const long long r = r0 + lane;
      const bool hit = r < n && idx[r] == code;
      unsigned m = __ballot_sync(0xffffffffu, hit);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const float* src = xn + (r0 + b) * d + c0;
#pragma unroll
        for (int j = 0; j < MAXV; ++j)
          if (c0 + lane + 32 * j < d) acc[j] += src[lane + 32 * j];
        count += 1.0f;
      }
    }
#pragma unroll
    for (int j = 0; j < MAXV; ++j)
      if (c0 + lane + 32 * j < d) stats[k + code * d + c0 + lane + 32 * j] = acc[j];
    if (c0 == 0 && lane == 0) stats[code] = count;
  }
}

// ----------------------------------------------------------------------------------
// EMA update of the cosine codebook (one warp per code).          l2_quantize.py:421-438
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vq_ema_cosine_kernel(float* __restrict__ embed, float* __restrict__ cluster, const float* en,
                     const float* __restrict__ stats, long long k, int d, float decay,
                     float* en_next, __half* __restrict__ eh_next) {
  const int lane = threadIdx.x & 31;
  const long long code = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (code >= k) return;
  const float one_minus = (float)(1.0 - (double)decay);
  const float bins = stats[code];
  if (lane == 0) cluster[code] = cluster[code] * decay + bins * one_minus;
  const float* es = stats + k + code * d;
  float* e = embed + code * d;
  if (bins == 0.0f) {
    for (int c = lane; c < d; c += 32) e[c] = e[c] * decay + en[code * d + c] * one_minus;
  } else {
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = es[c] / bins; ss += v * v; }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < d; c += 32) {
      const float v = (es[c] / bins) / denom;
      e[c] = e[c] * decay + v * one_minus;
    }
  }
  if (en_next || eh_next) {
    // The next search needs l2norm(new code) (fp32, and 16x that in fp16): emitted here, with the
    // arithmetic of vq_prepare_rows_kernel (same lane-strided sum, same shuffle tree, same division),
    // so a cached row is bit-identical to a freshly prepared one.  en_next may alias en: each warp
    // has finished reading its row of en above and writes only that row (the lanes re-read their own
    // stores of e).
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = e[c]; ss += v * v; }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < d; c += 32) {
      const float v = e[c] / denom;
      if (en_next) en_next[code * d + c] = v;
      if (eh_next) eh_next[code * d + c] = __float2half_rn(v * 16.0f);
    }
  }
}

// EMA update of the Euclidean codebook.                            l2_quantize.py:292-300
__global__ void __launch_bounds__(1024)
vq_ema_euclid_cluster_kernel(float* __restrict__ cluster, const float* __restrict__ stats, long long k,
                             float decay, float* __restrict__ total) {
  __shared__ double sh[1024];
  const float one_minus = (float)(1.0 - (double)decay);
  double acc = 0.0;
  for (long long i = threadIdx.x; i < k; i += 1024) {
    const float v = cluster[i] * decay + stats[i] * one_minus;
    cluster[i] = v;
    acc += (double)v;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = (float)sh[0];
}
__global__ void __launch_bounds__(256)
vq_ema_euclid_embed_kernel(float* __restrict__ embed, const float* __restrict__ cluster,
                           const float* __restrict__ embed_avg, long long k, int d, float eps,
                           const float* __restrict__ total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * d) return;
  const long long code = i / d;
  const float tot = total[0];
  const float smoothed = (cluster[code] + eps) / (tot + (float)k * eps) * tot;
  embed[i] = embed_avg[i] / smoothed;
}

// backward of straight-through + commitment loss.                   l2_quantize.py:554-561
__global__ void __launch_bounds__(256)
vq_backward_kernel(const float* __restrict__ x, const float* __restrict__ out, const float* __restrict__ g_out,
                   const float* __restrict__ g_loss, long long numel, float coef, float* __restrict__ gx) {
  const float gl = g_loss ? g_loss[0] * coef : 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const float g = g_out ? g_out[i] : 0.f;
    gx[i] = g + gl * (x[i] - out[i]);
  }
}

static inline int ew_blocks(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace favae

using namespace favae;

extern "C" {

int favae_vq_prepare_rows(const float* x, int64_t n, int d, int64_t hw, int normalize, float* xn,
                          void* xh, float* sq, void* stream) {
  FAVAE_REQUIRE(x && n >= 0 && d > 0 && hw >= 1, "vq_prepare_rows: bad arguments");
  FAVAE_REQUIRE(n % hw == 0, "vq_prepare_rows: n must be a multiple of hw");
  if (n == 0) return 0;
  const size_t smem = sizeof(float) * PREP_ROWS * (size_t)(d + 1);
  FAVAE_REQUIRE(smem <= 200 * 1024, "vq_prepare_rows: d too large");
  static PerDevice<size_t> configured_dev;
  size_t& configured = configured_dev.here();
  if (smem > 48 * 1024 && smem > configured) {
    FAVAE_CUDA_OK(cudaFuncSetAttribute(vq_prepare_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long blocks = (hw == 1) ? (n + PREP_ROWS - 1) / PREP_ROWS
                                     : (n / hw) * ((hw + PREP_ROWS - 1) / PREP_ROWS);
  vq_prepare_rows_kernel<<<(unsigned)blocks, PREP_THREADS, smem, (cudaStream_t)stream>>>(
      x, n, d, hw, normalize, xn, (__half*)xh, sq);
  return check_launch("vq_prepare_rows");
}

int favae_vq_search_exact(const float* xn, const float* en, const float* e_sq, int64_t n, int64_t k,
                          int d, int metric, uint64_t* keys, int64_t* idx, void* stream) {
  FAVAE_REQUIRE(xn && en && keys && idx && n >= 0 && k > 0 && d > 0, "vq_search_exact: bad arguments");
  FAVAE_REQUIRE(metric == 0 || e_sq, "vq_search_exact: metric 1 needs e_sq");
  FAVAE_REQUIRE(k < 0xFFFFFFFFll, "vq_search_exact: codebook too large");
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  FAVAE_CUDA_OK(cudaMemsetAsync(keys, 0, sizeof(uint64_t) * (size_t)n, s));
  const long long row_tiles = (n + SB - 1) / SB, code_tiles = (k + SB - 1) / SB;
  long long splits = (2LL * num_sms() + row_tiles - 1) / row_tiles;
  if (splits > code_tiles) splits = code_tiles;
  if (splits < 1) splits = 1;
  const long long cps = ((code_tiles + splits - 1) / splits) * SB;
  splits = (k + cps - 1) / cps;
  FAVAE_REQUIRE(splits <= 65535, "vq_search_exact: too many splits");
  dim3 grid((unsigned)row_tiles, (unsigned)splits);
  vq_search_exact_kernel<<<grid, 256, 0, s>>>(xn, en, e_sq, n, k, d, metric, cps,
                                              (unsigned long long*)keys);
  int rc = check_launch("vq_search_exact");
  if (rc) return rc;
  vq_keys_to_idx_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const unsigned long long*)keys, n,
                                                                  (long long*)idx);
  return check_launch("vq_keys_to_idx");
}

int favae_vq_gather_st(const float* x, const float* embed, const int64_t* idx, int64_t n, int64_t k,
                       int d, int64_t hw, int straight_through, float* out, float* partials,
                       float* loss_sum, void* stream) {
  FAVAE_REQUIRE(embed && idx && out && n >= 0 && d > 0 && hw >= 1 && k > 0, "vq_gather_st: bad arguments");
  FAVAE_REQUIRE(n % hw == 0, "vq_gather_st: n must be a multiple of hw");
  FAVAE_REQUIRE(!loss_sum || (partials && x), "vq_gather_st: loss_sum needs x and partials");
  if (n == 0) {
    if (loss_sum) FAVAE_CUDA_OK(cudaMemsetAsync(loss_sum, 0, sizeof(float), (cudaStream_t)stream));
    return 0;
  }
  const size_t smem = sizeof(float) * PREP_ROWS * (size_t)(d + 1);
  FAVAE_REQUIRE(smem <= 200 * 1024, "vq_gather_st: d too large");
  static PerDevice<size_t> configured_dev;
  size_t& configured = configured_dev.here();
  if (smem > 48 * 1024 && smem > configured) {
    FAVAE_CUDA_OK(cudaFuncSetAttribute(vq_gather_st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long blocks = (hw == 1) ? (n + PREP_ROWS - 1) / PREP_ROWS
                                     : (n / hw) * ((hw + PREP_ROWS - 1) / PREP_ROWS);
  vq_gather_st_kernel<<<(unsigned)blocks, PREP_THREADS, smem, (cudaStream_t)stream>>>(
      x, embed, (const long long*)idx, n, k, d, hw, straight_through, out, loss_sum ? partials : nullptr);
  int rc = check_launch("vq_gather_st");
  if (rc || !loss_sum) return rc;
  return favae_sum_scaled(partials, blocks, 1.0, loss_sum, stream);
}

int favae_vq_gather_rows(const float* embed, const int64_t* idx, int64_t n, int64_t k, int d,
                         int64_t hw, float* out, void* stream) {
  return favae_vq_gather_st(nullptr, embed, idx, n, k, d, hw, 0, out, nullptr, nullptr, stream);
}

int favae_vq_code_stats(const float* xn, const int64_t* idx, int64_t n, int64_t k, int d,
                        int deterministic, float* stats, void* stream) {
  FAVAE_REQUIRE(xn && idx && stats && n >= 0 && k > 0 && d > 0, "vq_code_stats: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (deterministic && n > 0) {
    vq_code_stats_det_kernel<<<(unsigned)((k + 7) / 8), 256, 0, s>>>(xn, (const long long*)idx, n, k, d, stats);
    return check_launch("vq_code_stats_det");
  }
  FAVAE_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(float) * (size_t)k * (size_t)(d + 1), s));
  if (n == 0) return 0;
  long long blocks = (n + 7) / 8;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  vq_code_stats_kernel<<<(unsigned)blocks, 256, 0, s>>>(xn, (const long long*)idx, n, k, d, stats);
  return check_launch("vq_code_stats");
}

int favae_vq_ema_update_cosine(float* embed, float* cluster_size, const float* en, const float* stats,
                               int64_t k, int d, float decay, float* en_next, void* eh_next, void* stream) {
  FAVAE_REQUIRE(embed && cluster_size && en && stats && k > 0 && d > 0, "vq_ema_update_cosine: bad arguments");
  vq_ema_cosine_kernel<<<(unsigned)((k + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      embed, cluster_size, en, stats, k, d, decay, en_next, (__half*)eh_next);
  return check_launch("vq_ema_update_cosine");
}

int favae_vq_ema_update_euclid(float* embed, float* cluster_size, const float* embed_avg,
                               const float* stats, int64_t k, int d, float decay, float eps,
                               float* scratch2, void* stream) {
  FAVAE_REQUIRE(embed && cluster_size && embed_avg && stats && scratch2 && k > 0 && d > 0,
                "vq_ema_update_euclid: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  vq_ema_euclid_cluster_kernel<<<1, 1024, 0, s>>>(cluster_size, stats, k, decay, scratch2);
  int rc = check_launch("vq_ema_euclid_cluster");
  if (rc) return rc;
  vq_ema_euclid_embed_kernel<<<(unsigned)((k * d + 255) / 256), 256, 0, s>>>(embed, cluster_size, embed_avg,
                                                                          k, d, eps, scratch2);
  return check_launch("vq_ema_euclid_embed");
}

int favae_vq_backward(const float* x, const float* out, const float* g_out, const float* g_loss,
                      int64_t numel, float coef, float* gx, void* stream) {
  FAVAE_REQUIRE(x && out && gx && numel >= 0, "vq_backward: bad arguments");
  if (numel == 0) return 0;
  vq_backward_kernel<<<ew_blocks(numel, 256), 256, 0, (cudaStream_t)stream>>>(x, out, g_out, g_loss, numel,
                                                                              coef, gx);
  return check_launch("vq_backward");
}
}
