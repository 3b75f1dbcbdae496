/* favae_b200 -- C ABI of the B200-native FA-VAE hot path (VQ search + spectrum losses).
 *
 * The reference (oppo-us-research/FA-VAE) has no FFI layer: its boundary is the Python
 * object protocol of models/l2_quantize.py and losses/vqgan_losses.py.  This header is the
 * C boundary underneath the drop-in Python modules in favae_b200/: every entry point names
 * the reference code it replaces (paths relative to /root/reference).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - the caller owns every buffer (including workspaces, sized by the *_workspace_bytes
 *     queries); kernels never allocate;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value 0 = success, otherwise a cudaError_t (>0) or -EINVAL-style code (<0);
 *     favae_last_error() returns a thread-local message;
 *   - "rows layout": a latent matrix (N rows x D channels) is addressed either row-major
 *     (hw == 1) or as an NCHW feature map (N = B*hw, element (n,c) at
 *     x[(n / hw) * D * hw + c * hw + n % hw]), which fuses the reference's
 *     `rearrange(x, 'b c h w -> b (h w) c')` and its inverse (l2_quantize.py:540,593).
 */
#ifndef FAVAE_B200_H
#define FAVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

#define FAVAE_B200_ABI_VERSION 2

int favae_abi_version(void);
const char* favae_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py gpu_launches) */
long long favae_launch_count(void);

/* ------------------------------------------------------------------ vector quantizer */

/* Row preparation: l2norm (l2_quantize.py:24-25, called at :403 and :408) fused with the
 * NCHW -> (N,D) rearrange (:540).  normalize == 0 copies instead (Euclidean codebook).
 * xn  (N*D fp32 row-major, nullable), xh (N*D fp16 row-major, nullable; holds 16 * xn so that
 * small components stay out of the fp16 subnormal range -- input of favae_vq_search_tc),
 * sq  (N fp32, nullable): sum of squares of the OUTPUT rows of xn. */
int favae_vq_prepare_rows(const float* x, int64_t n, int d, int64_t hw, int normalize,
                          float* xn, void* xh, float* sq, void* stream);

/* Exact fp32 nearest-code search (CUDA cores): replaces einsum + argmax
 * (l2_quantize.py:410-411) or -cdist + argmax (:280-282).
 * metric 0: argmax_k <xn_n, en_k>;  metric 1: argmax_k (2<x_n,e_k> - |e_k|^2)  (e_sq = |e_k|^2).
 * Ties resolve to the lowest index.  keys: N uint64 scratch.  idx: N int64. */
int favae_vq_search_exact(const float* xn, const float* en, const float* e_sq, int64_t n,
                          int64_t k, int d, int metric, uint64_t* keys, int64_t* idx,
                          void* stream);

/* Tensor-core nearest-code search (tcgen05 / TMEM / TMA): one fp16 UMMA pass that keeps, per
 * latent, every code whose approximate similarity is within a proven error bound of the
 * maximum, followed by an exact fp32 re-score of those candidates.  Same result contract as
 * favae_vq_search_exact with metric 0.  Needs xh/eh (fp16) and xn/en (fp32) from
 * favae_vq_prepare_rows.  Supported when d % 64 == 0, d <= 256 and k % 256 == 0; the workspace
 * query returns 0 otherwise (callers then use favae_vq_search_exact).  `keys` is unused. */
size_t favae_vq_search_tc_workspace_bytes(int64_t n, int64_t k, int d);
int favae_vq_search_tc(const void* xh, const void* eh, const float* xn, const float* en,
                       int64_t n, int64_t k, int d, void* workspace, size_t workspace_bytes,
                       uint64_t* keys, int64_t* idx, void* stream);

/* Diagnostics (synchronises): number of latents of the last favae_vq_search_tc call on this
 * workspace that took the exhaustive fp32 fallback (candidate list overflow). */
int favae_vq_search_tc_overflow_rows(const void* workspace, int64_t n, int64_t k, int d, int* count_host);

/* Gather + straight-through + commitment-loss partial sums: replaces batched_embedding
 * (l2_quantize.py:166-170, :415), `x + (q - x).detach()` (:554) and the mse numerator (:560).
 * out has the layout of x.  loss_sum (1 float) receives sum((out - x)^2) (deterministic
 * two-level reduction; partials = ceil(n/32) floats of scratch). */
int favae_vq_gather_st(const float* x, const float* embed, const int64_t* idx, int64_t n,
                       int64_t k, int d, int64_t hw, int straight_through, float* out,
                       float* partials, float* loss_sum, void* stream);

/* Code usage statistics: bins = bincount(idx) (l2_quantize.py:412,418) and
 * embed_sum = scatter-add of xn rows by idx (:426, the reference's second dense GEMM).
 * stats = [bins (K) | embed_sum (K*D)] fp32, zeroed by this call: one flat buffer so that the
 * data-parallel exchange is a single all-reduce (reference: two, :419 and :427).
 * deterministic == 0: scatter-add with fp32 atomics (arrival order, ~1e-7 run-to-run noise);
 * deterministic != 0: one warp per code adds its members in ascending latent order --
 * bit-reproducible (the reference's one-hot GEMM, :426, is deterministic too). */
int favae_vq_code_stats(const float* xn, const int64_t* idx, int64_t n, int64_t k, int d,
                        int deterministic, float* stats, void* stream);

/* EMA codebook update, cosine codebook (l2_quantize.py:421-438). en = l2norm(embed) as
 * produced by favae_vq_prepare_rows on the pre-update codebook.
 * en_next (K*D fp32, nullable, may alias en) and eh_next (K*D fp16, nullable) receive
 * l2norm(updated embed) and 16x that in fp16 -- bit-identical to what favae_vq_prepare_rows
 * would produce -- so the next search (l2_quantize.py:408, re-normalising the whole codebook
 * every call) starts without a preparation pass. */
int favae_vq_ema_update_cosine(float* embed, float* cluster_size, const float* en,
                               const float* stats, int64_t k, int d, float decay,
                               float* en_next, void* eh_next, void* stream);

/* EMA update, Euclidean codebook (l2_quantize.py:292-300) including its quirk that
 * embed_avg is never refreshed. */
int favae_vq_ema_update_euclid(float* embed, float* cluster_size, const float* embed_avg,
                               const float* stats, int64_t k, int d, float decay, float eps,
                               float* scratch2, void* stream);

/* Backward of (:554-561): gx = g_out + coef * (x - out) * g_loss[0], coef = 2*w/(N*D).
 * g_out / g_loss nullable (treated as zero).  All tensors share the layout of x. */
int favae_vq_backward(const float* x, const float* out, const float* g_out, const float* g_loss,
                      int64_t numel, float coef, float* gx, void* stream);

/* get_codebook_entry (l2_quantize.py:518-530): rows of embed gathered into an NCHW (hw > 1)
 * or row-major tensor. */
int favae_vq_gather_rows(const float* embed, const int64_t* idx, int64_t n, int64_t k, int d,
                         int64_t hw, float* out, void* stream);

/* ------------------------------------------------------------------ spectrum losses */

/* Fused focal-frequency / spectrum loss over `maps` independent h x w real maps (h == w, a
 * power of two in [8, 512]): replaces FocalFrequencyLoss.tensor2freq + loss_formulation
 * (pip focal-frequency-loss==0.3.0; call sites losses/vqgan_losses.py:14,25-26,45-46) and
 * their autograd backward.  map_loss[m] = sum_{u,v} w * |F(pred - target)|^2 (ortho FFT);
 * grad_pred / grad_target (nullable) receive +/- grad_scale * N^2 * Re ifft2(w . F) -- pass
 * grad_scale = 2 * loss_weight / numel.
 * target == NULL: `pred` already holds the difference map pred - target (the fused DSL op,
 * favae_blur_diff_forward); grad_target must then be NULL and grad_pred receives the single
 * gradient map G = dL/d(difference) (8 instead of 16 bytes per element).
 * map_max (nullable) receives max_{u,v} f(|F|) per map; fmax_override (nullable, one device
 * scalar) replaces the per-map maximum in the weight (batch_matrix=True).
 * alpha == 1 without log weighting and grad_scale >= 0 (every call the reference makes) runs a
 * leaner instantiation of the same kernel; results agree with the general one to rounding.
 * Diagnostics only: FAVAE_FFL256=c4 selects the 4-CTA-cluster variant for 256 x 256 maps. */
int favae_ffl_supported(int h, int w);
int favae_ffl_forward(const float* pred, const float* target, int64_t maps, int h, int w,
                      float alpha, int log_matrix, float grad_scale, float* map_loss,
                      float* grad_pred, float* grad_target, float* map_max,
                      const float* fmax_override, void* stream);

/* The same contract for ANY map size 1 <= h, w <= 2048 (the reference's torch.fft.fft2 takes any
 * size): direct separable DFT passes through a caller-provided workspace of
 * favae_ffl_generic_workspace_bytes(maps, h, w) bytes (maps are processed in chunks of at most
 * 256 MB of spectrum).  Cold path: O(h w (h + w)) multiply-adds per map; every FA-VAE
 * configuration produces power-of-two squares, which take favae_ffl_forward. */
size_t favae_ffl_generic_workspace_bytes(int64_t maps, int h, int w);
int favae_ffl_forward_generic(const float* pred, const float* target, int64_t maps, int h, int w,
                              float alpha, int log_matrix, float grad_scale, float* map_loss,
                              float* grad_pred, float* grad_target, float* map_max,
                              const float* fmax_override, void* workspace, void* stream);

/* out[0] = scale * sum(v[0..n)) accumulated in fp64, deterministic. */
int favae_sum_scaled(const float* v, int64_t n, double scale, float* out, void* stream);

/* a[i] *= s[0], b[i] *= s[0] unless s[0] == 1 (then no memory traffic).  b nullable. */
int favae_scale_inplace(float* a, float* b, int64_t n, const float* s, void* stream);

/* ------------------------------------------------------------------ Gaussian blur */

/* Reflect-padded separable Gaussian blur of `maps` h x w maps with a k x k kernel built from
 * the device scalar sigma[0]: replaces VQGANFCM._gaussian_blur (models/vqgan_fcm.py:20-41; the
 * copies in models/codec.py:255-277,625-646,947-968,1076-1097) and T.GaussianBlur
 * (losses/vqgan_losses.py:35). */
int favae_blur_forward(const float* x, int64_t maps, int h, int w, int ksize, const float* sigma,
                       float* y, void* stream);
/* gx = s * adjoint blur of gy;  gsigma[0] = s * d/dsigma <gy, blur(x)> with
 * s = out_scale * (out_scale_dev ? out_scale_dev[0] : 1)  (gsigma nullable; partials =
 * favae_blur_partials(maps,h,w) floats of scratch).  s != 1 needs
 * favae_blur_fast_supported(h, w, ksize): the fused DSL op passes out_scale = -1 for the encoder side
 * and the upstream gradient of the level as the device scalar, so the gradient map G is never
 * re-scaled in HBM.
 * Diagnostics only: the environment variable FAVAE_BLUR_SIGMA=split computes the two results
 * with two kernels instead of the fused one. */
int64_t favae_blur_partials(int64_t maps, int h, int w);
int favae_blur_backward(const float* gy, const float* x, int64_t maps, int h, int w, int ksize,
                        const float* sigma, float out_scale, const float* out_scale_dev, float* gx,
                        float* gsigma, float* partials, void* stream);

/* Both sides of a fused DSL level in one pass over the gradient map G (20 instead of 24 bytes per
 * element and ~19 % fewer instructions than two favae_blur_backward calls):
 *   g_dec = +s A_dec(G), gsigma_dec = +s d/dsigma_dec <G, blur(x_dec)>,
 *   g_enc = -s A_enc(G), gsigma_enc = -s d/dsigma_enc <G, blur(x_enc)>,   s = scale_dev ? scale_dev[0] : 1
 * (ffl(pred = blur(dec), target = blur(enc)), losses/vqgan_losses.py:25).  partials:
 * 2 * favae_blur_partials(maps, h, w) floats.  Needs favae_blur_fast_supported and 16-byte aligned maps. */
int favae_blur_backward_pair(const float* gy, const float* x_enc, const float* x_dec, int64_t maps, int h,
                             int w, int ksize, const float* sigma_enc, const float* sigma_dec,
                             const float* scale_dev, float* g_enc, float* g_dec, float* gsigma_enc,
                             float* gsigma_dec, float* partials, void* stream);

/* 1 when the streaming blur kernels take this shape: w a power of two in [8, 512],
 * ksize in {3,5,9,11,15}, ksize/2 < min(h, w). */
int favae_blur_fast_supported(int h, int w, int ksize);

/* d = blur(dec, sigma_dec) - blur(enc, sigma_enc) in one pass (12 bytes per element): the
 * blurred FCM feature pair of the DSL (models/vqgan_fcm.py:131-134; models/codec.py:284-309,
 * 655-686, 978-999, 1105-1123) is consumed only by ffl(de_feat, en_feat)
 * (losses/vqgan_losses.py:25), which depends on the difference alone, so the two blurred maps
 * are never materialised.  Needs favae_blur_fast_supported and 16-byte aligned maps. */
int favae_blur_diff_forward(const float* enc, const float* dec, int64_t maps, int h, int w,
                            int ksize, const float* sigma_enc, const float* sigma_dec, float* d,
                            void* stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* FAVAE_B200_H */
