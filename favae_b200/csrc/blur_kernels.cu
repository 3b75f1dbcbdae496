// Reflect-padded separable Gaussian blur with a device-resident (learnable) sigma:
// forward, adjoint (gradient w.r.t. the input) and the sigma gradient.
// Replaces VQGANFCM._gaussian_blur (/root/reference/models/vqgan_fcm.py:20-41 and the four
// copies in models/codec.py) and T.GaussianBlur (losses/vqgan_losses.py:35).
//
// The reference builds the k x k outer product and runs a depthwise conv2d on the
// reflect-padded map; the outer product of two identical 1-D Gaussians is applied here as
// two 1-D passes over a shared-memory tile (same arithmetic up to fp32 rounding order).
#include <stdlib.h>

#include "blur_fast.cuh"
#include "common.cuh"

namespace favae {

constexpr int BT = 32;           // output tile side
constexpr int BMAXK = 31;        // largest kernel size
constexpr int BMAXP = BMAXK / 2;
constexpr int BTH = BT + 2 * BMAXP;

struct BlurWeights {
  float k[BMAXK];
  float dk[BMAXK];              // d k / d sigma
};

// k1d = exp(-0.5 (x/sigma)^2) / sum,  x = linspace(-(ks-1)/2, (ks-1)/2, ks)   vqgan_fcm.py:20-26
__device__ __forceinline__ void blur_weights(BlurWeights& wt, int ks, float sigma) {
  if (threadIdx.x == 0) {
    const float half = (ks - 1) * 0.5f;
    float sum = 0.f;
    for (int t = 0; t < ks; ++t) {
      const float x = -half + (float)t;
      const float q = x / sigma;
      wt.k[t] = expf(-0.5f * q * q);
      sum += wt.k[t];
    }
    float m2 = 0.f;
    for (int t = 0; t < ks; ++t) {
      wt.k[t] /= sum;
      const float x = -half + (float)t;
      m2 += wt.k[t] * x * x;
    }
    const float s3 = sigma * sigma * sigma;
    for (int t = 0; t < ks; ++t) {
      const float x = -half + (float)t;
      wt.dk[t] = wt.k[t] * (x * x - m2) / s3;
    }
  }
}

__device__ __forceinline__ int reflect_clamp(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);
}

struct TileGeom {
  long long map;
  int y0, x0;
};
__device__ __forceinline__ TileGeom tile_geom(int h, int w) {
  const int tx = (w + BT - 1) / BT, ty = (h + BT - 1) / BT;
  TileGeom g;
  g.map = blockIdx.x / (tx * ty);
  const int t = blockIdx.x % (tx * ty);
  g.y0 = (t / tx) * BT;
  g.x0 = (t % tx) * BT;
  return g;
}

__global__ void __launch_bounds__(256)
blur_forward_kernel(const float* __restrict__ x, int h, int w, int ks, const float* __restrict__ sigma,
                    float* __restrict__ y) {
  __shared__ BlurWeights wt;
  __shared__ float in[BTH][BTH + 1];
  __shared__ float hz[BTH][BT + 1];
  const int p = ks / 2, th = BT + 2 * p;
  const TileGeom g = tile_geom(h, w);
  blur_weights(wt, ks, sigma[0]);
  const float* src = x + g.map * (long long)h * w;
  for (int i = threadIdx.x; i < th * th; i += blockDim.x) {
    const int iy = i / th, ix = i % th;
    in[iy][ix] = src[(long long)reflect_clamp(g.y0 + iy - p, h) * w + reflect_clamp(g.x0 + ix - p, w)];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < th * BT; i += blockDim.x) {
    const int iy = i / BT, ox = i % BT;
    float acc = 0.f;
    for (int t = 0; t < ks; ++t) acc = fmaf(wt.k[t], in[iy][ox + t], acc);
    hz[iy][ox] = acc;
  }
  __syncthreads();
  float* dst = y + g.map * (long long)h * w;
  for (int i = threadIdx.x; i < BT * BT; i += blockDim.x) {
    const int oy = i / BT, ox = i % BT;
    if (g.y0 + oy < h && g.x0 + ox < w) {
      float acc = 0.f;
      for (int t = 0; t < ks; ++t) acc = fmaf(wt.k[t], hz[oy + t][ox], acc);
      dst[(long long)(g.y0 + oy) * w + g.x0 + ox] = acc;
    }
  }
}

// zero-padded correlation at padded coordinate j plus the positions that reflect onto it
template <class F> __device__ __forceinline__ float adjoint_taps(int i, int n, int p, int ks, int lo, int span,
                                                                 const float* k, F value) {
  // value(r) = source sample at tile-local row/col r (0 <= r < span), the tile starting at `lo - p`
  float acc = 0.f;
  int src[3];
  int cnt = 0;
  src[cnt++] = i;
  if (i >= 1 && i <= p) src[cnt++] = -i;
  if (i >= n - 1 - p && i <= n - 2) src[cnt++] = 2 * (n - 1) - i;
  for (int q = 0; q < cnt; ++q) {
    const int base = src[q] - lo;             // tile-local index of tap s = 0
    for (int s = 0; s < ks; ++s) {
      const int r = base + s;
      if (r >= 0 && r < span) acc = fmaf(k[s], value(r), acc);
    }
  }
  return acc;
}

__global__ void __launch_bounds__(256)
blur_adjoint_kernel(const float* __restrict__ gy, int h, int w, int ks, const float* __restrict__ sigma,
                    float* __restrict__ gx) {
  __shared__ BlurWeights wt;
  __shared__ float in[BTH][BTH + 1];
  __shared__ float hz[BTH][BT + 1];
  const int p = ks / 2, th = BT + 2 * p;
  const TileGeom g = tile_geom(h, w);
  blur_weights(wt, ks, sigma[0]);
  const float* src = gy + g.map * (long long)h * w;
  for (int i = threadIdx.x; i < th * th; i += blockDim.x) {
    const int iy = i / th, ix = i % th;
    const int yy = g.y0 + iy - p, xx = g.x0 + ix - p;
    in[iy][ix] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? src[(long long)yy * w + xx] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < th * BT; i += blockDim.x) {
    const int iy = i / BT, ox = i % BT;
    const int col = g.x0 + ox;
    float acc = 0.f;
    if (col < w) acc = adjoint_taps(col, w, p, ks, g.x0, th, wt.k, [&](int r) { return in[iy][r]; });
    hz[iy][ox] = acc;
  }
  __syncthreads();
  float* dst = gx + g.map * (long long)h * w;
  for (int i = threadIdx.x; i < BT * BT; i += blockDim.x) {
    const int oy = i / BT, ox = i % BT;
    const int row = g.y0 + oy, col = g.x0 + ox;
    if (row < h && col < w) {
      // rows of hz outside the map are exact zeros only if the source rows were: the halo
      // rows were loaded as zeros above, so hz[r] is zero there too.
      dst[(long long)row * w + col] =
          adjoint_taps(row, h, p, ks, g.y0, th, wt.k, [&](int r) { return hz[r][ox]; });
    }
  }
}

__global__ void __launch_bounds__(256)
blur_sigma_grad_kernel(const float* __restrict__ gy, const float* __restrict__ x, int h, int w, int ks,
                       const float* __restrict__ sigma, float* __restrict__ partials) {
  __shared__ BlurWeights wt;
  __shared__ float in[BTH][BTH + 1];
  __shared__ float hk[BTH][BT + 1];
  __shared__ float hd[BTH][BT + 1];
  __shared__ float wsum[8];
  const int p = ks / 2, th = BT + 2 * p;
  const TileGeom g = tile_geom(h, w);
  blur_weights(wt, ks, sigma[0]);
  const float* src = x + g.map * (long long)h * w;
  for (int i = threadIdx.x; i < th * th; i += blockDim.x) {
    const int iy = i / th, ix = i % th;
    in[iy][ix] = src[(long long)reflect_clamp(g.y0 + iy - p, h) * w + reflect_clamp(g.x0 + ix - p, w)];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < th * BT; i += blockDim.x) {
    const int iy = i / BT, ox = i % BT;
    float a = 0.f, b = 0.f;
    for (int t = 0; t < ks; ++t) {
      const float v = in[iy][ox + t];
      a = fmaf(wt.k[t], v, a);
      b = fmaf(wt.dk[t], v, b);
    }
    hk[iy][ox] = a;
    hd[iy][ox] = b;
  }
  __syncthreads();
  const float* gsrc = gy + g.map * (long long)h * w;
  float acc = 0.f;
  for (int i = threadIdx.x; i < BT * BT; i += blockDim.x) {
    const int oy = i / BT, ox = i % BT;
    if (g.y0 + oy < h && g.x0 + ox < w) {
      float dy = 0.f;
      for (int t = 0; t < ks; ++t) dy += wt.dk[t] * hk[oy + t][ox] + wt.k[t] * hd[oy + t][ox];
      acc = fmaf(gsrc[(long long)(g.y0 + oy) * w + g.x0 + ox], dy, acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += wsum[i];
    partials[blockIdx.x] = s;
  }
}

static inline long long blur_blocks(long long maps, int h, int w) {
  return maps * ((h + BT - 1) / BT) * ((w + BT - 1) / BT);
}

}  // namespace favae

using namespace favae;

extern "C" {

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int64_t favae_blur_partials(int64_t maps, int h, int w) {
  long long b = blur_blocks(maps, h, w);
  if (w >= 4 && w <= 512 && (w & (w - 1)) == 0) {
    const long long f = blurf::num_blocks(maps, h, w);
    if (f > b) b = f;
  }
  return b;
}

static int blur_check(int64_t maps, int h, int w, int ksize) {
  FAVAE_REQUIRE(maps >= 0 && h > 0 && w > 0, "blur: bad shape");
  FAVAE_REQUIRE(ksize >= 1 && ksize <= BMAXK && (ksize & 1), "blur: kernel size must be odd and <= 31");
  FAVAE_REQUIRE(ksize / 2 < h && ksize / 2 < w, "blur: reflect padding needs kernel_size // 2 < map size");
  FAVAE_REQUIRE(blur_blocks(maps, h, w) < 2147483647ll, "blur: too many tiles");
  return 0;
}

int favae_blur_forward(const float* x, int64_t maps, int h, int w, int ksize, const float* sigma,
                       float* y, void* stream) {
  FAVAE_REQUIRE(x && y && sigma, "blur_forward: null pointer");
  int rc = blur_check(maps, h, w, ksize);
  if (rc || maps == 0) return rc;
  if (blurf::supported(h, w, ksize) && aligned16(x) && aligned16(y))
    return blurf::launch<blurf::MODE_FWD>(x, nullptr, maps, h, w, ksize, sigma, y, nullptr, (cudaStream_t)stream);
  blur_forward_kernel<<<(unsigned)blur_blocks(maps, h, w), 256, 0, (cudaStream_t)stream>>>(x, h, w, ksize,
                                                                                         sigma, y);
  return check_launch("blur_forward");
}

int favae_blur_fast_supported(int h, int w, int ksize) { return blurf::supported(h, w, ksize) ? 1 : 0; }

int favae_blur_diff_forward(const float* enc, const float* dec, int64_t maps, int h, int w, int ksize,
                            const float* sigma_enc, const float* sigma_dec, float* d, void* stream) {
  FAVAE_REQUIRE(enc && dec && d && sigma_enc && sigma_dec, "blur_diff_forward: null pointer");
  int rc = blur_check(maps, h, w, ksize);
  if (rc || maps == 0) return rc;
  FAVAE_REQUIRE(blurf::supported(h, w, ksize) && aligned16(enc) && aligned16(dec) && aligned16(d),
                "blur_diff_forward: needs a power-of-two width in [8,512], kernel size in {3,5,9,11,15} and "
                "16-byte aligned maps (favae_blur_fast_supported)");
  return blurf::launch_diff(enc, dec, maps, h, w, ksize, sigma_enc, sigma_dec, d, (cudaStream_t)stream);
}

int favae_blur_backward_pair(const float* gy, const float* x_enc, const float* x_dec, int64_t maps, int h, int w,
                             int ksize, const float* sigma_enc, const float* sigma_dec, const float* scale_dev,
                             float* g_enc, float* g_dec, float* gsigma_enc, float* gsigma_dec, float* partials,
                             void* stream) {
  FAVAE_REQUIRE(gy && x_enc && x_dec && sigma_enc && sigma_dec && g_enc && g_dec && gsigma_enc && gsigma_dec && partials,
                "blur_backward_pair: null pointer");
  int rc = blur_check(maps, h, w, ksize);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (maps == 0) {
    FAVAE_CUDA_OK(cudaMemsetAsync(gsigma_enc, 0, sizeof(float), s));
    FAVAE_CUDA_OK(cudaMemsetAsync(gsigma_dec, 0, sizeof(float), s));
    return 0;
  }
  FAVAE_REQUIRE(blurf::supported(h, w, ksize) && aligned16(gy) && aligned16(x_enc) && aligned16(x_dec) &&
                    aligned16(g_enc) && aligned16(g_dec),
                "blur_backward_pair: needs favae_blur_fast_supported and 16-byte aligned maps");
  rc = blurf::launch_pair(gy, x_enc, x_dec, maps, h, w, ksize, sigma_enc, sigma_dec, g_enc, g_dec, partials, scale_dev, s);
  if (rc) return rc;
  const long long blocks = blurf::num_blocks(maps, h, w, blurf::MODE_PAIR);
  rc = favae_sum_scaled(partials, blocks, 1.0, gsigma_enc, stream);
  if (rc) return rc;
  return favae_sum_scaled(partials + blocks, blocks, 1.0, gsigma_dec, stream);
}

int favae_blur_backward(const float* gy, const float* x, int64_t maps, int h, int w, int ksize,
                        const float* sigma, float out_scale, const float* out_scale_dev, float* gx, float* gsigma,
                        float* partials, void* stream) {
  FAVAE_REQUIRE(gy && sigma && (gx || gsigma), "blur_backward: null pointer");
  FAVAE_REQUIRE(!gsigma || (x && partials), "blur_backward: sigma gradient needs x and partials");
  int rc = blur_check(maps, h, w, ksize);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (maps == 0) {
    if (gsigma) FAVAE_CUDA_OK(cudaMemsetAsync(gsigma, 0, sizeof(float), s));
    return 0;
  }
  if (gx && blurf::supported(h, w, ksize) && aligned16(gy) && aligned16(gx) && (!gsigma || aligned16(x))) {
    if (!gsigma) return blurf::launch<blurf::MODE_ADJ>(gy, nullptr, maps, h, w, ksize, sigma, gx, nullptr, s, out_scale, out_scale_dev);
    // One fused kernel (12 B/element: gy and x read once, gx written).  FAVAE_BLUR_SIGMA=split runs
    // the plain adjoint and the forward-path sigma-gradient kernel instead (16 B/element).  Measured
    // on B200 at 4096 maps of 256^2, k = 9: fused 1.11 ms, split 0.56 + 0.69 ms.
    static const bool split = [] { const char* e = getenv("FAVAE_BLUR_SIGMA"); return e && e[0] == 's'; }();
    if (!split) {
      rc = blurf::launch<blurf::MODE_ADJ_SIG>(gy, x, maps, h, w, ksize, sigma, gx, partials, s, out_scale, out_scale_dev);
    } else {
      FAVAE_REQUIRE(!out_scale_dev, "blur_backward: FAVAE_BLUR_SIGMA=split does not take a device-side scale");
      rc = blurf::launch<blurf::MODE_ADJ>(gy, nullptr, maps, h, w, ksize, sigma, gx, nullptr, s, out_scale, out_scale_dev);
      if (rc) return rc;
      rc = blurf::launch<blurf::MODE_SIGMA>(x, gy, maps, h, w, ksize, sigma, nullptr, partials, s);
    }
    if (rc) return rc;
    // the fused kernel's partials already carry out_scale (it rides on the D row scaling)
    return favae_sum_scaled(partials, blurf::num_blocks(maps, h, w, split ? blurf::MODE_SIGMA : blurf::MODE_ADJ_SIG),
                            split ? (double)out_scale : 1.0, gsigma, stream);
  }
  FAVAE_REQUIRE(out_scale == 1.0f && !out_scale_dev,
                "blur_backward: an output scale needs the streaming path (favae_blur_fast_supported)");
  const unsigned blocks = (unsigned)blur_blocks(maps, h, w);
  if (gx) {
    blur_adjoint_kernel<<<blocks, 256, 0, s>>>(gy, h, w, ksize, sigma, gx);
    rc = check_launch("blur_adjoint");
    if (rc) return rc;
  }
  if (gsigma) {
    blur_sigma_grad_kernel<<<blocks, 256, 0, s>>>(gy, x, h, w, ksize, sigma, partials);
    rc = check_launch("blur_sigma_grad");
    if (rc) return rc;
    return favae_sum_scaled(partials, blocks, 1.0, gsigma, stream);
  }
  return 0;
}
}
