// latent.cu -- the latent-side terms of the importance log-weights: sums of Normal log-densities over the latent axis.
//
// Replaces  lpz  = reduce_sum(pz.log_prob(z),  axis=pz.axes)    models/loss.py:28
//           lqzx = reduce_sum(qzx.log_prob(z), axis=qzx.axes)   models/loss.py:30
//           beta * (lpz - lqzx)                                  models/loss.py:34
// and model06's four terms  (lpz2 - lqz2z1) + (lpz1z2 - lqz1x)  models/model06.py:40-47,
// plus the derivatives tf.GradientTape takes of them w.r.t. z, loc and scale.  The tensors are tiny ([S,B,D], D ~ 20):
// the point is ONE launch (forward) and ONE launch (backward) instead of ~10 element-wise / reduction launches each,
// and a fixed summation order.
//
//   term t:  T_t[s,b] = sum_d log N(z_t[s,b,d]; loc_t[.,b,d], scale_t[.,b,d])
//            log N(z; m, s) = -((z - m)^2) / (2 s^2) - log s - log sqrt(2 pi)            (tfd.Normal / torch Normal)
//   extra[s,b] = (extra_in ? extra_in[s,b] : 0) + sum_t weight_t * T_t[s,b]
// loc == NULL means the standard normal (loc 0, scale 1: the prior of models/model05.py:208).  loc / scale are either
// [B, D] (shared by the S samples: the encoder's q(z|x), models/model05.py:124) or [S, B, D] (params_per_sample = 1:
// model06's conditional layers).
#include "common.cuh"

namespace vaemdl {

constexpr float kHalfLog2Pi = 0.9189385332046727f;

struct LtArgs {
  vaemdl_latent_term t[VAEMDL_MAX_LATENT_TERMS];
  float* dz[VAEMDL_MAX_LATENT_TERMS];
  float* dloc[VAEMDL_MAX_LATENT_TERMS];
  float* dscale[VAEMDL_MAX_LATENT_TERMS];
  int dz_accumulate[VAEMDL_MAX_LATENT_TERMS];  // this term adds to a dz buffer an earlier term already wrote
  int n;
  int S;
  long long B;
  const float* extra_in;
  float* extra_out;
  float* term_sums;
  const float* g_extra;
  int max_D;
};

__device__ __forceinline__ double lt_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// one warp per (s, b)
__global__ void __launch_bounds__(256) latent_fwd_kernel(const LtArgs a) {
  const long long n = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long n_img = static_cast<long long>(a.S) * a.B;
  if (n >= n_img) return;
  const long long b = n % a.B;
  double extra = a.extra_in ? static_cast<double>(a.extra_in[n]) : 0.0;
  for (int t = 0; t < a.n; ++t) {
    const vaemdl_latent_term& tm = a.t[t];
    const float* z = tm.z + n * tm.D;
    const long long prow = (tm.params_per_sample ? n : b) * tm.D;
    double acc = 0.0;
    for (int d = lane; d < tm.D; d += 32) {
      const float m = tm.loc ? tm.loc[prow + d] : 0.0f;
      const float s = tm.loc ? tm.scale[prow + d] : 1.0f;
      const float r = (z[d] - m) / s;
      acc += static_cast<double>(-0.5f * r * r - logf(s) - kHalfLog2Pi);
    }
    acc = lt_warp_sum(acc);
    if (lane == 0 && a.term_sums) a.term_sums[static_cast<long long>(t) * n_img + n] = static_cast<float>(acc);
    extra += static_cast<double>(tm.weight) * acc;
  }
  if (lane == 0 && a.extra_out) a.extra_out[n] = static_cast<float>(extra);
}

// one thread per (b, d): walks the S samples in order, so the [B, D] parameter gradients are sums in a fixed order and
// every dz element is touched by exactly one thread (terms that share a z tensor accumulate safely)
__global__ void __launch_bounds__(256) latent_bwd_kernel(const LtArgs a) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= a.B * a.max_D) return;
  const long long b = idx / a.max_D;
  const int d = static_cast<int>(idx - b * a.max_D);
  for (int t = 0; t < a.n; ++t) {
    const vaemdl_latent_term& tm = a.t[t];
    if (d >= tm.D) continue;
    double acc_m = 0.0, acc_s = 0.0;
    for (int s = 0; s < a.S; ++s) {
      const long long n = static_cast<long long>(s) * a.B + b;
      const long long zi = n * tm.D + d;
      const long long pi = (tm.params_per_sample ? n : b) * tm.D + d;
      const float m = tm.loc ? tm.loc[pi] : 0.0f;
      const float sc = tm.loc ? tm.scale[pi] : 1.0f;
      const float gw = a.g_extra[n] * tm.weight;
      const float inv = 1.0f / sc;
      const float r = (tm.z[zi] - m) * inv;       // (z - m) / s
      const float dm = gw * r * inv;              // d/dm  = (z - m) / s^2
      const float ds = gw * (r * r - 1.0f) * inv; // d/ds  = ((z - m)^2 / s^2 - 1) / s
      if (a.dz[t]) {
        if (a.dz_accumulate[t])
          a.dz[t][zi] += -dm;
        else
          a.dz[t][zi] = -dm;                      // d/dz  = -(z - m) / s^2
      }
      if (tm.params_per_sample) {
        if (a.dloc[t]) a.dloc[t][pi] = dm;
        if (a.dscale[t]) a.dscale[t][pi] = ds;
      } else {
        acc_m += static_cast<double>(dm);
        acc_s += static_cast<double>(ds);
      }
    }
    if (!tm.params_per_sample) {
      if (a.dloc[t]) a.dloc[t][b * tm.D + d] = static_cast<float>(acc_m);
      if (a.dscale[t]) a.dscale[t][b * tm.D + d] = static_cast<float>(acc_s);
    }
  }
}

static int lt_check(const vaemdl_latent_term* terms, int n_terms, int S, long long B) {
  if (!terms || n_terms < 1 || n_terms > VAEMDL_MAX_LATENT_TERMS || S <= 0 || B <= 0) return VAEMDL_EINVAL;
  for (int t = 0; t < n_terms; ++t) {
    if (!terms[t].z || terms[t].D <= 0) return VAEMDL_EINVAL;
    if ((terms[t].loc == nullptr) != (terms[t].scale == nullptr)) return VAEMDL_EINVAL;
  }
  return VAEMDL_OK;
}

}  // namespace vaemdl

using namespace vaemdl;

extern "C" int vaemdl_latent_terms_fwd(const vaemdl_latent_term* terms, int n_terms, int S, long long B,
                                       const float* extra_in, float* extra_out, float* term_sums, void* stream) {
  int rc = lt_check(terms, n_terms, S, B);
  if (rc) return rc;
  if (!extra_out && !term_sums) return VAEMDL_EINVAL;
  LtArgs a{};
  for (int t = 0; t < n_terms; ++t) a.t[t] = terms[t];
  a.n = n_terms;
  a.S = S;
  a.B = B;
  a.extra_in = extra_in;
  a.extra_out = extra_out;
  a.term_sums = term_sums;
  const long long threads = static_cast<long long>(S) * B * 32;
  latent_fwd_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cuda_rc(cudaGetLastError());
}

extern "C" int vaemdl_latent_terms_bwd(const vaemdl_latent_term* terms, int n_terms, int S, long long B,
                                       const float* g_extra, float* const* dz, float* const* dloc, float* const* dscale,
                                       void* stream) {
  int rc = lt_check(terms, n_terms, S, B);
  if (rc) return rc;
  if (!g_extra || !dz || !dloc || !dscale) return VAEMDL_EINVAL;
  LtArgs a{};
  a.max_D = 0;
  for (int t = 0; t < n_terms; ++t) {
    a.t[t] = terms[t];
    a.dz[t] = dz[t];
    a.dloc[t] = dloc[t];
    a.dscale[t] = dscale[t];
    a.dz_accumulate[t] = 0;
    for (int u = 0; u < t; ++u)
      if (dz[t] && dz[u] == dz[t]) {
        if (terms[u].D != terms[t].D) return VAEMDL_EINVAL;  // a shared gradient buffer means a shared z tensor
        a.dz_accumulate[t] = 1;
      }
    if (terms[t].D > a.max_D) a.max_D = terms[t].D;
  }
  a.n = n_terms;
  a.S = S;
  a.B = B;
  a.g_extra = g_extra;
  const long long threads = B * a.max_D;
  latent_bwd_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return cuda_rc(cudaGetLastError());
}
