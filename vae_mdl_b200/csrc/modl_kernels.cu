// modl_kernels.cu -- C-ABI entry points of the MoDL log-likelihood / gradient with the green / blue means chained on the
// OBSERVED x (utils/mdl.py:56-207, utils/mdl_openai.py:83-157, utils/mdl_openai_iwae.py:33-67; models/model05.py:141-145).
// Kernels, launchers and the shared host logic live in modl_kernels.cuh.
#include "modl_kernels.cuh"

using namespace vaemdl;

extern "C" size_t vaemdl_modl_workspace_bytes(long long n_img, int H, int W) {
  if (n_img <= 0 || H <= 0 || W <= 0) return 0;
  // one float64 partial per (forward warp, image its run of tiles touches): at most n_img + 3 * warps + 1 of them
  // (also covers the atomic path's one accumulator per image); + the fused finish kernel's block sums / the unfused
  // tail's per-image scratch (n_img) and the arrival counter
  (void)H;
  (void)W;
  return partial_elems(n_img) * sizeof(double) + static_cast<size_t>(n_img) * sizeof(double) + 256;
}

extern "C" size_t vaemdl_modl_step_workspace_bytes(long long n_img, int H, int W) {
  if (n_img <= 0 || H <= 0 || W <= 0) return 0;
  // the forward workspace + one (mixture sum, logit normaliser) pair per pixel-sample for the one-pass gradient
  return vaemdl_modl_workspace_bytes(n_img, H, W) + static_cast<size_t>(n_img) * H * W * sizeof(float2);
}

extern "C" int vaemdl_modl_fwd(const float* params, const void* x, int x_dtype, int x_range, int edge_mode,
                               long long n_img, int x_batch, int H, int W, int M, float* lp_pixel, float* ll_image,
                               double* ll_image_f64, void* workspace, size_t workspace_bytes, void* stream) {
  return modl_fwd_impl<0>(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M, lp_pixel, ll_image, ll_image_f64,
                          IwaeOut{}, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int vaemdl_modl_iwae_fwd(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, int S,
                                    long long B, long long B_total, int x_batch, int H, int W, int M, const float* extra,
                                    float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo,
                                    float* g_ll, void* workspace, size_t workspace_bytes, void* stream) {
  return modl_iwae_fwd_impl<0>(params, x, x_dtype, x_range, edge_mode, S, B, B_total, x_batch, H, W, M, extra, ll_image,
                               ll_image_f64, log_w, lme_b, elbo, g_ll, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int vaemdl_modl_bwd(const float* params, const void* x, int x_dtype, int x_range, int edge_mode,
                               long long n_img, int x_batch, int H, int W, int M, const float* g_image,
                               const float* g_pixel, float* dparams, void* stream) {
  return modl_bwd_impl<0>(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M, g_image, g_pixel, dparams,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int vaemdl_modl_iwae_step(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, int S,
                                     long long B, long long B_total, int x_batch, int H, int W, int M, const float* extra,
                                     float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo,
                                     float* g_ll, float* dparams, void* workspace, size_t workspace_bytes, void* stream,
                                     int* launches) {
  return modl_iwae_step_impl<0>(params, x, x_dtype, x_range, edge_mode, S, B, B_total, x_batch, H, W, M, extra, ll_image,
                                ll_image_f64, log_w, lme_b, elbo, g_ll, dparams, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream), launches);
}

// ---- bfloat16 parameters / gradient; x-conditioned class, all arithmetic in float32 ----------------------------
extern "C" int vaemdl_modl_fwd_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                                    long long n_img, int x_batch, int H, int W, int M, float* lp_pixel, float* ll_image,
                                    double* ll_image_f64, void* workspace, size_t workspace_bytes, void* stream) {
  return modl_fwd_impl<0>(static_cast<const float*>(params_bf16), x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M,
                          lp_pixel, ll_image, ll_image_f64, IwaeOut{}, workspace, workspace_bytes,
                          static_cast<cudaStream_t>(stream), 1);
}

extern "C" int vaemdl_modl_iwae_fwd_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                                         int S, long long B, long long B_total, int x_batch, int H, int W, int M,
                                         const float* extra, float* ll_image, double* ll_image_f64, float* log_w,
                                         float* lme_b, float* elbo, float* g_ll, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  return modl_iwae_fwd_impl<0>(static_cast<const float*>(params_bf16), x, x_dtype, x_range, edge_mode, S, B, B_total,
                               x_batch, H, W, M, extra, ll_image, ll_image_f64, log_w, lme_b, elbo, g_ll, workspace,
                               workspace_bytes, static_cast<cudaStream_t>(stream), 1);
}

extern "C" int vaemdl_modl_bwd_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                                    long long n_img, int x_batch, int H, int W, int M, const float* g_image,
                                    const float* g_pixel, void* dparams_bf16, void* stream) {
  return modl_bwd_impl<0>(static_cast<const float*>(params_bf16), x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M,
                          g_image, g_pixel, static_cast<float*>(dparams_bf16), static_cast<cudaStream_t>(stream), 1);
}

// bfloat16 parameters with the per-pixel mixture sums handed from the forward to the backward call: the backward kernel then
// keeps the tile in bfloat16 (two slots per warp, one-pass gradient rounded once).  pix_stats [n_img*H*W*2] float32.
extern "C" int vaemdl_modl_iwae_fwd_stats_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                                               int S, long long B, long long B_total, int x_batch, int H, int W, int M,
                                               const float* extra, float* ll_image, double* ll_image_f64, float* log_w,
                                               float* lme_b, float* elbo, float* g_ll, float* pix_stats, void* workspace,
                                               size_t workspace_bytes, void* stream) {
  return modl_iwae_fwd_impl<0>(static_cast<const float*>(params_bf16), x, x_dtype, x_range, edge_mode, S, B, B_total,
                               x_batch, H, W, M, extra, ll_image, ll_image_f64, log_w, lme_b, elbo, g_ll, workspace,
                               workspace_bytes, static_cast<cudaStream_t>(stream), 1, pix_stats);
}

extern "C" int vaemdl_modl_bwd_stats_bf16(const void* params_bf16, const void* x, int x_dtype, int x_range, int edge_mode,
                                          long long n_img, int x_batch, int H, int W, int M, const float* g_image,
                                          const float* g_pixel, const float* pix_stats, void* dparams_bf16, void* stream) {
  return modl_bwd_impl<0>(static_cast<const float*>(params_bf16), x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M,
                          g_image, g_pixel, static_cast<float*>(dparams_bf16), static_cast<cudaStream_t>(stream), 1, pix_stats);
}

// ---- forward / backward as two calls that share the per-pixel mixture sums (what vaemdl_modl_iwae_step does inside) -----------
extern "C" int vaemdl_modl_iwae_fwd_stats(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, int S,
                                          long long B, long long B_total, int x_batch, int H, int W, int M,
                                          const float* extra, float* ll_image, double* ll_image_f64, float* log_w,
                                          float* lme_b, float* elbo, float* g_ll, float* pix_stats, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  return modl_iwae_fwd_impl<0>(params, x, x_dtype, x_range, edge_mode, S, B, B_total, x_batch, H, W, M, extra, ll_image,
                               ll_image_f64, log_w, lme_b, elbo, g_ll, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream), 0, pix_stats);
}

extern "C" int vaemdl_modl_bwd_stats(const float* params, const void* x, int x_dtype, int x_range, int edge_mode,
                                     long long n_img, int x_batch, int H, int W, int M, const float* g_image,
                                     const float* g_pixel, const float* pix_stats, float* dparams, void* stream) {
  return modl_bwd_impl<0>(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M, g_image, g_pixel, dparams,
                          static_cast<cudaStream_t>(stream), 0, pix_stats);
}
