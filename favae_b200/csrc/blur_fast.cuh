// Streaming separable Gaussian blur for power-of-two map widths (4 <= W <= 512).
//
// A work item is a strip of TH rows of one map.  W/4 threads own four adjacent columns each
// (float4 HBM accesses); the vertical pass is a KS-row sliding window held in registers (the
// row loop is fully unrolled so the window is renamed, never moved); the horizontal pass reads
// the vertically-filtered row from a double-buffered shared-memory line.  Each input element is
// read once from HBM (strip halos hit L2), each output written once.
//
//   MODE_FWD      y  = H V x          reflect padding          (vqgan_fcm.py:35-41)
//   MODE_ADJ      gx = H^T V^T gy     zero padding + the taps that reflect onto border rows/cols
//   MODE_ADJ_SIG  MODE_ADJ plus  d/dsigma <gy, H V x> = <(H'^T V^T + H^T V'^T) gy, x>  with
//                 k' = dk/dsigma, sharing the single sliding window over gy (x is read once,
//                 without halo)                                (autograd of vqgan_fcm.py:20-41)
#pragma once

#include "common.cuh"

namespace favae {
namespace blurf {

constexpr int THREADS = 128;
constexpr int MODE_FWD = 0, MODE_ADJ = 1, MODE_ADJ_SIG = 2;
constexpr int LPAD = 8;                          // left halo slots of a shared line (>= p, 16B aligned)

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
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

// static border fixes of the horizontal adjoint: column j within P of a border also receives the
// taps of the padded columns that reflect onto it.  X0 / R0 are compile-time so every index is.
template <int KS, int X0>
__device__ __forceinline__ void left_fix(float (&o)[4], const float (&k)[KS], const float* line) {
  constexpr int P = KS / 2;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    constexpr int dummy = 0; (void)dummy;
    const int j = X0 + c;
    if (j >= 1 && j <= P) {
#pragma unroll
      for (int s = 0; s < KS; ++s)
        if (s >= j + P) o[c] = fmaf(k[s], line[LPAD + s - j - P], o[c]);
    }
  }
}
// thread owning columns w-4-R0 .. w-1-R0 (R0 = 0 or 4): distance from the right border jr = R0 + 3 - c
template <int KS, int R0>
__device__ __forceinline__ void right_fix(float (&o)[4], const float (&k)[KS], const float* line, int w) {
  constexpr int P = KS / 2;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int jr = R0 + 3 - c;
    if (jr >= 1 && jr <= P) {
#pragma unroll
      for (int s = 0; s < KS; ++s)
        if (s <= P - jr) o[c] = fmaf(k[s], line[LPAD + w - 1 + jr - P + s], o[c]);
    }
  }
}

template <int KS, int TH, int MODE>
__global__ void __launch_bounds__(THREADS)
blur_fast_kernel(const float* __restrict__ src, const float* __restrict__ aux, int h, int w, long long items,
                 int strips, const float* __restrict__ sigma, float* __restrict__ dst,
                 float* __restrict__ partials) {
  // src: x (FWD) or gy (ADJ*); aux: x (ADJ_SIG); dst: y / gx
  constexpr int P = KS / 2;
  constexpr bool ADJ = MODE != MODE_FWD, SIG = MODE == MODE_ADJ_SIG;
  constexpr int NV = SIG ? 2 : 1;
  extern __shared__ float lines[];               // [2 buffers][groups][NV][w + 2*LPAD]
  __shared__ float sk[32], sdk[32];
  __shared__ float wred[THREADS / 32];
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
  float acc_sigma = 0.f;

  float4 win[KS];
#pragma unroll
  for (int t = 0; t < KS; ++t) win[t] = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll
  for (int r = 0; r < TH + KS - 1; ++r) {
    // ---- vertical: slide one input row into the window
    const int yi = y0 - P + r;
    float4 in = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      if (ADJ) {
        if (yi >= 0 && yi < h) in = ld4(base + (long long)yi * w + x0);
      } else {
        in = ld4(base + (long long)reflect_idx(yi, h) * w + x0);
      }
    }
    win[r % KS] = in;
    if (r < KS - 1) continue;
    const int yo = y0 + r - (KS - 1);            // output row of this iteration
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), vd = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < KS; ++t) {
      const float4& wv = win[(r + 1 + t) % KS];  // input row yo - P + t
      fma4(v, k[t], wv);
      if (SIG) fma4(vd, dk[t], wv);
    }
    if (ADJ && live && yo < h) {
      // rows within P of a border also receive the taps of the padded rows that reflect onto them
      if (yo >= 1 && yo <= P) {
        for (int rr = 0; rr <= P - yo; ++rr) {          // G(-yo): taps s = yo + P + rr, source row rr
          if (rr < h) {
            const float4 g = ld4(base + (long long)rr * w + x0);
            fma4(v, sk[yo + P + rr], g);
            if (SIG) fma4(vd, sdk[yo + P + rr], g);
          }
        }
      }
      if (yo >= h - 1 - P && yo <= h - 2) {
        const int jr = h - 1 - yo;                      // G(h-1+jr): taps s = 0 .. P-jr, row h-1+jr-P+s
        for (int s2 = 0; s2 <= P - jr; ++s2) {
          const int rr = h - 1 + jr - P + s2;
          if (rr >= 0) {
            const float4 g = ld4(base + (long long)rr * w + x0);
            fma4(v, sk[s2], g);
            if (SIG) fma4(vd, sdk[s2], g);
          }
        }
      }
    }
    // ---- horizontal through a shared line
    float* line = lines + ((size_t)((r & 1) * groups + grp) * NV) * ll;
    *reinterpret_cast<float4*>(line + LPAD + x0) = v;
    if (SIG) *reinterpret_cast<float4*>(line + ll + LPAD + x0) = vd;
    if (!ADJ) {
      // mirrored halo entries (reflect): index -j <- j, index w-1+j <- w-1-j
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = x0 + c;
        if (j >= 1 && j <= P) line[LPAD - j] = vv[c];
        const int jr = w - 1 - j;
        if (jr >= 1 && jr <= P) line[LPAD + w - 1 + jr] = vv[c];
      }
    } else {
      if (tx == 0) {
#pragma unroll
        for (int j = 1; j <= LPAD; ++j) { line[LPAD - j] = 0.f; if (SIG) line[ll + LPAD - j] = 0.f; }
      }
      if (tx == tpi - 1) {
#pragma unroll
        for (int j = 1; j <= LPAD; ++j) { line[LPAD + w - 1 + j] = 0.f; if (SIG) line[ll + LPAD + w - 1 + j] = 0.f; }
      }
    }
    __syncthreads();
    float seg[4 + 2 * P], segd[SIG ? 4 + 2 * P : 1];
#pragma unroll
    for (int i = 0; i < 4 + 2 * P; ++i) {
      seg[i] = line[LPAD + x0 - P + i];
      if (SIG) segd[i] = line[ll + LPAD + x0 - P + i];
    }
    float o[4] = {0.f, 0.f, 0.f, 0.f}, z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int t = 0; t < KS; ++t) {
        o[c] = fmaf(k[t], seg[c + t], o[c]);
        if (SIG) z[c] = fmaf(dk[t], seg[c + t], fmaf(k[t], segd[c + t], z[c]));
      }
    }
    if (ADJ) {
      if (tx == 0) {
        left_fix<KS, 0>(o, k, line);
        if (SIG) { left_fix<KS, 0>(z, dk, line); left_fix<KS, 0>(z, k, line + ll); }
      } else if (tx == 1 && P > 3) {
        left_fix<KS, 4>(o, k, line);
        if (SIG) { left_fix<KS, 4>(z, dk, line); left_fix<KS, 4>(z, k, line + ll); }
      }
      if (tx == tpi - 1) {
        right_fix<KS, 0>(o, k, line, w);
        if (SIG) { right_fix<KS, 0>(z, dk, line, w); right_fix<KS, 0>(z, k, line + ll, w); }
      } else if (tx == tpi - 2 && P > 3) {
        right_fix<KS, 4>(o, k, line, w);
        if (SIG) { right_fix<KS, 4>(z, dk, line, w); right_fix<KS, 4>(z, k, line + ll, w); }
      }
    }
    if (live && yo < h) {
      const long long a = map * (long long)h * w + (long long)yo * w + x0;
      *reinterpret_cast<float4*>(dst + a) = make_float4(o[0], o[1], o[2], o[3]);
      if (SIG) {
        const float4 xv = ld4(aux + a);
        acc_sigma = fmaf(xv.x, z[0], fmaf(xv.y, z[1], fmaf(xv.z, z[2], fmaf(xv.w, z[3], acc_sigma))));
      }
    }
  }
  if (SIG) {
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
}

inline bool supported(int h, int w, int ks) {
  const bool pow2 = w >= 4 && w <= 512 && (w & (w - 1)) == 0;
  const bool kok = ks == 3 || ks == 5 || ks == 9 || ks == 11 || ks == 15;
  // w >= 8 keeps the two border-owning threads distinct; halo slots need ks/2 <= LPAD
  return pow2 && w >= 8 && kok && ks / 2 < h && ks / 2 < w && ks / 2 <= LPAD;
}
inline int strip_rows(int h) { return h <= 16 ? 16 : 32; }
inline long long num_blocks(long long maps, int h, int w) {
  const int th = strip_rows(h), strips = (h + th - 1) / th, groups = THREADS / (w / 4);
  return (maps * strips + groups - 1) / groups;
}

template <int KS, int TH, int MODE>
static int launch_one(const float* src, const float* aux, long long maps, int h, int w, const float* sigma,
                      float* dst, float* partials, cudaStream_t s) {
  const int strips = (h + TH - 1) / TH, groups = THREADS / (w / 4);
  const long long items = maps * strips;
  const long long blocks = (items + groups - 1) / groups;
  const size_t smem = sizeof(float) * 2 * groups * ((MODE == MODE_ADJ_SIG) ? 2 : 1) * (size_t)(w + 2 * LPAD);
  blur_fast_kernel<KS, TH, MODE><<<(unsigned)blocks, THREADS, smem, s>>>(src, aux, h, w, items, strips, sigma, dst,
                                                                      partials);
  return check_launch("blur_fast");
}

template <int MODE>
static int launch(const float* src, const float* aux, long long maps, int h, int w, int ks, const float* sigma,
                  float* dst, float* partials, cudaStream_t s) {
#define FAVAE_BLUR_CASE(KS)                                                                         \
  case KS:                                                                                          \
    return strip_rows(h) == 16 ? launch_one<KS, 16, MODE>(src, aux, maps, h, w, sigma, dst, partials, s) \
                               : launch_one<KS, 32, MODE>(src, aux, maps, h, w, sigma, dst, partials, s);
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
