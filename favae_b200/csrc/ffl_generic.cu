// Generic spectrum loss for ANY map size H x W (the reference's torch.fft.fft2 takes any size; every
// FA-VAE configuration produces power-of-two squares, which take the fused shared-memory kernel in
// ffl_kernels.cu).  Cold path, kept simple: the 2-D DFT of d = pred - target is evaluated directly as
// two separable passes of O(H W (H + W)) multiply-adds per map with exact twiddle indexing
// ((k * n) mod N looked up in a shared-memory table of the N-th roots of unity), the spectrum lives in
// a caller-provided workspace, the per-map statistics / weights / inverse passes follow the same
// definitions as the fused kernel:
//   U = sum d e^{-2 pi i (uy/H + vx/W)},  A^2 = |U|^2 / (HW),  f = A^alpha [log(1 + f)],
//   w = clamp(nan_to_0(f / max f), 0, 1),  map_loss = sum_f w A^2,
//   grad_pred = grad_scale / (HW) * Re sum_f w U e^{+2 pi i (uy/H + vx/W)} = -grad_target.
// Replaces FocalFrequencyLoss.tensor2freq + loss_formulation (pip focal-frequency-loss==0.3.0) for the
// shapes favae_ffl_supported rejects.
#include "common.cuh"

namespace favae {
namespace fflg {

constexpr int MAXN = 2048;                      // roots table in shared memory: 16 KB

__device__ __forceinline__ void build_roots(float2* tw, int n, float sign) {
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)j / (float)n, &s, &c);
    tw[j] = make_float2(c, sign * s);
  }
  __syncthreads();
}

// pass along the contiguous axis: out[m][r][v] = sum_c in[m][r][c] * root_W^(sign * v * c)
// REAL_IN: in is the real difference map (pred - target); REAL_OUT: out is the scaled real part.
template <bool REAL_IN, bool REAL_OUT>
__global__ void __launch_bounds__(256)
row_pass_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float2* __restrict__ cin,
                float2* __restrict__ cout, float* __restrict__ gp, float* __restrict__ gt, long long rows, int w,
                float sign, float scale) {
  extern __shared__ float2 sm[];                // roots (w) + one input row (w)
  float2* tw = sm;
  float2* row = sm + w;
  build_roots(tw, w, sign);
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    for (int c = threadIdx.x; c < w; c += blockDim.x) {
      if (REAL_IN) row[c] = make_float2(pred[r * w + c] - (target ? target[r * w + c] : 0.f), 0.f);
      else row[c] = cin[r * w + c];
    }
    __syncthreads();
    for (int v = threadIdx.x; v < w; v += blockDim.x) {
      float re = 0.f, im = 0.f;
      int k = 0;                                 // (v * c) mod w, updated incrementally
      for (int c = 0; c < w; ++c) {
        const float2 t = tw[k], a = row[c];
        re = fmaf(a.x, t.x, fmaf(-a.y, t.y, re));
        im = fmaf(a.x, t.y, fmaf(a.y, t.x, im));
        k += v;
        if (k >= w) k -= w;
      }
      if (REAL_OUT) {
        if (gp) gp[r * w + v] = re * scale;
        if (gt) gt[r * w + v] = -re * scale;
      } else {
        cout[r * w + v] = make_float2(re, im);
      }
    }
    __syncthreads();
  }
}

// pass along the strided axis: out[m][u][v] = sum_r in[m][r][v] * root_H^(sign * u * r); the weight
// (per map, per bin) is applied to the INPUT when wmul != nullptr (inverse direction).
__global__ void __launch_bounds__(256)
col_pass_kernel(const float2* __restrict__ cin, float2* __restrict__ cout, const float* __restrict__ wmul,
                long long maps, int h, int w, float sign) {
  extern __shared__ float2 sm[];
  float2* tw = sm;
  build_roots(tw, h, sign);
  const long long per_map = (long long)h * w;
  const long long total = maps * per_map;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / per_map;
    const int u = (int)((i % per_map) / w), v = (int)(i % w);
    const float2* src = cin + m * per_map + v;
    const float* ws = wmul ? wmul + m * per_map + v : nullptr;
    float re = 0.f, im = 0.f;
    int k = 0;
    for (int r = 0; r < h; ++r) {
      float2 a = src[(long long)r * w];
      if (ws) { const float q = ws[(long long)r * w]; a.x *= q; a.y *= q; }
      const float2 t = tw[k];
      re = fmaf(a.x, t.x, fmaf(-a.y, t.y, re));
      im = fmaf(a.x, t.y, fmaf(a.y, t.x, im));
      k += u;
      if (k >= h) k -= h;
    }
    cout[i] = make_float2(re, im);
  }
}

__device__ __forceinline__ float spectrum_f(float a2, float alpha, int log_matrix) {
  float f = (alpha == 1.0f) ? sqrtf(a2) : (alpha == 2.0f) ? a2 : powf(sqrtf(a2), alpha);
  if (log_matrix) f = logf(f + 1.0f);
  return f;
}

// one block per map: max f(A), then weights and sum_f w A^2
__global__ void __launch_bounds__(256)
stats_kernel(const float2* __restrict__ spec, float* __restrict__ wout, float* __restrict__ map_loss,
             float* __restrict__ map_max, const float* __restrict__ fmax_override, int h, int w, float alpha,
             int log_matrix) {
  __shared__ float red[256];
  const long long per_map = (long long)h * w;
  const float2* s = spec + blockIdx.x * per_map;
  const float inv_hw = 1.0f / (float)per_map;
  float mx = 0.f;
  for (long long i = threadIdx.x; i < per_map; i += 256) {
    const float2 u = s[i];
    mx = fmaxf(mx, spectrum_f((u.x * u.x + u.y * u.y) * inv_hw, alpha, log_matrix));
  }
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + st]);
    __syncthreads();
  }
  mx = red[0];
  __syncthreads();
  if (threadIdx.x == 0 && map_max) map_max[blockIdx.x] = mx;
  const float fmax = fmax_override ? fmax_override[0] : mx;
  const float inv = fmax > 0.f ? 1.0f / fmax : 0.f;                 // NaN -> 0 rule of the weight matrix
  double acc = 0.0;
  for (long long i = threadIdx.x; i < per_map; i += 256) {
    const float2 u = s[i];
    const float a2 = (u.x * u.x + u.y * u.y) * inv_hw;
    const float wgt = fminf(fmaxf(spectrum_f(a2, alpha, log_matrix) * inv, 0.f), 1.f);
    if (wout) wout[blockIdx.x * per_map + i] = wgt;
    acc += (double)(wgt * a2);
  }
  __shared__ double dred[256];
  dred[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if ((int)threadIdx.x < st) dred[threadIdx.x] += dred[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) map_loss[blockIdx.x] = (float)dred[0];
}

constexpr size_t CHUNK_BYTES = 256ull << 20;    // workspace cap: maps are processed in chunks

inline long long chunk_maps(long long maps, int h, int w) {
  const size_t per_map = (size_t)h * w * (2 * sizeof(float2) + sizeof(float));
  long long c = (long long)(CHUNK_BYTES / per_map);
  if (c < 1) c = 1;
  return c < maps ? c : maps;
}

}  // namespace fflg
}  // namespace favae

using namespace favae;

extern "C" {

size_t favae_ffl_generic_workspace_bytes(int64_t maps, int h, int w) {
  if (maps <= 0 || h <= 0 || w <= 0 || h > fflg::MAXN || w > fflg::MAXN) return 0;
  return (size_t)fflg::chunk_maps(maps, h, w) * h * w * (2 * sizeof(float2) + sizeof(float));
}

int favae_ffl_forward_generic(const float* pred, const float* target, int64_t maps, int h, int w, float alpha,
                              int log_matrix, float grad_scale, float* map_loss, float* grad_pred,
                              float* grad_target, float* map_max, const float* fmax_override, void* workspace,
                              void* stream) {
  FAVAE_REQUIRE(pred && map_loss && workspace, "ffl_forward_generic: null pointer");
  FAVAE_REQUIRE(target || !grad_target, "ffl_forward_generic: a gradient for a target that was not given");
  FAVAE_REQUIRE(h > 0 && w > 0 && h <= fflg::MAXN && w <= fflg::MAXN, "ffl_forward_generic: map side must be in [1, 2048]");
  if (maps <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const long long chunk = fflg::chunk_maps(maps, h, w);
  const size_t per_map = (size_t)h * w;
  float2* a = (float2*)workspace;
  float2* b = a + chunk * per_map;
  float* wgt = (float*)(b + chunk * per_map);
  const bool want_grad = grad_pred || grad_target;
  const int cap = num_sms() * 8;
  for (long long m0 = 0; m0 < maps; m0 += chunk) {
    const long long n = (maps - m0 < chunk) ? maps - m0 : chunk;
    const long long rows = n * h;
    const unsigned row_blocks = (unsigned)(rows < cap ? rows : cap);
    const long long elems = n * (long long)per_map;
    const unsigned col_blocks = (unsigned)((elems + 255) / 256 < cap ? (elems + 255) / 256 : cap);
    const float* p = pred + m0 * per_map;
    const float* t = target ? target + m0 * per_map : nullptr;
    fflg::row_pass_kernel<true, false><<<row_blocks, 256, sizeof(float2) * 2 * w, s>>>(
        p, t, nullptr, a, nullptr, nullptr, rows, w, -1.0f, 1.0f);
    int rc = check_launch("ffl_generic_rows");
    if (rc) return rc;
    fflg::col_pass_kernel<<<col_blocks, 256, sizeof(float2) * h, s>>>(a, b, nullptr, n, h, w, -1.0f);
    rc = check_launch("ffl_generic_cols");
    if (rc) return rc;
    fflg::stats_kernel<<<(unsigned)n, 256, 0, s>>>(b, want_grad ? wgt : nullptr, map_loss + m0,
                                                   map_max ? map_max + m0 : nullptr, fmax_override, h, w, alpha,
                                                   log_matrix);
    rc = check_launch("ffl_generic_stats");
    if (rc) return rc;
    if (!want_grad) continue;
    fflg::col_pass_kernel<<<col_blocks, 256, sizeof(float2) * h, s>>>(b, a, wgt, n, h, w, 1.0f);
    rc = check_launch("ffl_generic_icols");
    if (rc) return rc;
    fflg::row_pass_kernel<false, true><<<row_blocks, 256, sizeof(float2) * 2 * w, s>>>(
        nullptr, nullptr, a, nullptr, grad_pred ? grad_pred + m0 * per_map : nullptr,
        grad_target ? grad_target + m0 * per_map : nullptr, rows, w, 1.0f, grad_scale / (float)per_map);
    rc = check_launch("ffl_generic_irows");
    if (rc) return rc;
  }
  return 0;
}
}
