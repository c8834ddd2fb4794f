// modl_plain.cu -- C-ABI entry points of the pixel mixture of discretized logistics WITHOUT conditioning on the observed
// x: the green / blue means are chained on the component's own red / green means (utils/mdl_plain.py:7-66, :124-168).
// Same kernels as modl_kernels.cu, instantiated with AR = 1 (modl_kernels.cuh).
#include "modl_kernels.cuh"

using namespace vaemdl;

extern "C" int vaemdl_modl_plain_fwd(const float* params, const void* x, int x_dtype, long long n_img, int x_batch, int H,
                                     int W, int M, float low, float high, float levels, float* lp_pixel, float* ll_image, double* ll_image_f64, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  return modl_fwd_impl<1>(params, x, x_dtype, VAEMDL_RANGE_UNIT, VAEMDL_EDGE_MDL, n_img, x_batch, H, W, M, lp_pixel,
                          ll_image, ll_image_f64, IwaeOut{}, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), 0, nullptr,
                          BinGeom{low, high, levels});
}

extern "C" int vaemdl_modl_plain_iwae_fwd(const float* params, const void* x, int x_dtype, int S, long long B,
                                          long long B_total, int x_batch, int H, int W, int M, float low, float high,
                                          float levels, const float* extra, float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo,
                                          float* g_ll, void* workspace, size_t workspace_bytes, void* stream) {
  return modl_iwae_fwd_impl<1>(params, x, x_dtype, VAEMDL_RANGE_UNIT, VAEMDL_EDGE_MDL, S, B, B_total, x_batch, H, W, M,
                               extra, ll_image, ll_image_f64, log_w, lme_b, elbo, g_ll, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream), 0, nullptr, BinGeom{low, high, levels});
}

extern "C" int vaemdl_modl_plain_bwd(const float* params, const void* x, int x_dtype, long long n_img, int x_batch, int H,
                                     int W, int M, float low, float high, float levels, const float* g_image,
                                     const float* g_pixel, float* dparams, void* stream) {
  return modl_bwd_impl<1>(params, x, x_dtype, VAEMDL_RANGE_UNIT, VAEMDL_EDGE_MDL, n_img, x_batch, H, W, M, g_image, g_pixel,
                          dparams, static_cast<cudaStream_t>(stream), 0, nullptr, BinGeom{low, high, levels});
}

extern "C" int vaemdl_modl_plain_iwae_step(const float* params, const void* x, int x_dtype, int S, long long B,
                                           long long B_total, int x_batch, int H, int W, int M, float low, float high,
                                           float levels, const float* extra, float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo,
                                           float* g_ll, float* dparams, void* workspace, size_t workspace_bytes,
                                           void* stream, int* launches) {
  return modl_iwae_step_impl<1>(params, x, x_dtype, VAEMDL_RANGE_UNIT, VAEMDL_EDGE_MDL, S, B, B_total, x_batch, H, W, M,
                                extra, ll_image, ll_image_f64, log_w, lme_b, elbo, g_ll, dparams, workspace,
                                workspace_bytes, static_cast<cudaStream_t>(stream), launches, BinGeom{low, high, levels});
}
