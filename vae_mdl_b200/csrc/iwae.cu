// iwae.cu -- log-mean-exp over the importance-sample axis and the fused IWAE tail.
//
// Replaces logmeanexp (utils/utils.py:9-11) and the tail of iwae_loss (models/loss.py:34-43; models/model06.py:47-55).
// The tensors are tiny ([S,B], S <= 5000): the goal is latency and a fixed summation order, not bandwidth.
#include "common.cuh"

namespace vaemdl {

constexpr int kLmeThreads = 256;

// Block = BX columns (b) x SY sample-lanes.  Two passes over the block's S values (max, then sum of exp), each
// combined across the SY lanes through shared memory in a fixed order.
// TIN = float or double: float64 input keeps the differences log_w[s] - max exact to ~1e-12 when |log_w| ~ 2e4,
// where a float32 log_w would carry ~1e-3 (its ulp) straight into the softmax weights.
template <bool TAIL, typename TIN>
__global__ void __launch_bounds__(kLmeThreads)
    lme_kernel(const TIN* __restrict__ in, const float* __restrict__ extra, int S, long long B, int BX, int SY,
               float* __restrict__ log_w_out, float* __restrict__ lme_b, float* __restrict__ g_ll, float b_norm) {
  __shared__ float red[kLmeThreads];
  const int bx = threadIdx.x % BX, sy = threadIdx.x / BX;
  const long long b = static_cast<long long>(blockIdx.x) * BX + bx;
  const bool ok = b < B;
  __shared__ TIN redm[kLmeThreads];
  auto val = [&](int s) -> TIN {
    TIN v = in[static_cast<long long>(s) * B + b];
    if (TAIL && extra) v += static_cast<TIN>(extra[static_cast<long long>(s) * B + b]);  // models/loss.py:34
    return v;
  };
  TIN mx = -INFINITY;
  if (ok)
    for (int s = sy; s < S; s += SY) mx = fmax(mx, val(s));
  redm[threadIdx.x] = mx;
  __syncthreads();
  mx = redm[bx];
  for (int j = 1; j < SY; ++j) mx = fmax(mx, redm[j * BX + bx]);  // utils/utils.py:10
  float sm = 0.0f;
  if (ok)
    for (int s = sy; s < S; s += SY) sm += expf(static_cast<float>(val(s) - mx));
  red[threadIdx.x] = sm;
  __syncthreads();
  sm = red[bx];
  for (int j = 1; j < SY; ++j) sm += red[j * BX + bx];
  if (!ok) return;
  if (sy == 0 && lme_b)
    lme_b[b] = static_cast<float>(static_cast<TIN>(logf(sm / static_cast<float>(S))) + mx);  // utils/utils.py:11
  if (TAIL) {
    const float scale = -1.0f / (sm * b_norm);  // d(-mean_b lme_b)/d log_w = -softmax_s / B
    for (int s = sy; s < S; s += SY) {
      const TIN v = val(s);
      if (log_w_out) log_w_out[static_cast<long long>(s) * B + b] = static_cast<float>(v);
      if (g_ll) g_ll[static_cast<long long>(s) * B + b] = expf(static_cast<float>(v - mx)) * scale;
    }
  }
}

// dlog_w[s,b] = g_out[b] * softmax_s(log_w[:,b])
template <typename TIN>
__global__ void __launch_bounds__(kLmeThreads)
    lme_bwd_kernel(const TIN* __restrict__ log_w, const float* __restrict__ g_out, int S, long long B, int BX, int SY,
                   float* __restrict__ dlog_w) {
  __shared__ float red[kLmeThreads];
  const int bx = threadIdx.x % BX, sy = threadIdx.x / BX;
  const long long b = static_cast<long long>(blockIdx.x) * BX + bx;
  const bool ok = b < B;
  __shared__ TIN redm[kLmeThreads];
  TIN mx = -INFINITY;
  if (ok)
    for (int s = sy; s < S; s += SY) mx = fmax(mx, log_w[static_cast<long long>(s) * B + b]);
  redm[threadIdx.x] = mx;
  __syncthreads();
  mx = redm[bx];
  for (int j = 1; j < SY; ++j) mx = fmax(mx, redm[j * BX + bx]);
  float sm = 0.0f;
  if (ok)
    for (int s = sy; s < S; s += SY) sm += expf(static_cast<float>(log_w[static_cast<long long>(s) * B + b] - mx));
  red[threadIdx.x] = sm;
  __syncthreads();
  sm = red[bx];
  for (int j = 1; j < SY; ++j) sm += red[j * BX + bx];
  if (!ok) return;
  const float scale = g_out[b] / sm;
  for (int s = sy; s < S; s += SY)
    dlog_w[static_cast<long long>(s) * B + b] =
        expf(static_cast<float>(log_w[static_cast<long long>(s) * B + b] - mx)) * scale;
}

// elbo = sum_b lme_b / b_norm, single block, fixed order
__global__ void __launch_bounds__(kLmeThreads)
    mean_kernel(const float* __restrict__ v, long long B, float b_norm, float* __restrict__ out) {
  __shared__ float red[kLmeThreads];
  float acc = 0.0f;
  for (long long i = threadIdx.x; i < B; i += kLmeThreads) acc += v[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kLmeThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0] / b_norm;  // models/loss.py:37
}

static void lme_shape(long long B, int& BX, int& SY) {
  BX = 32;
  while (BX > 1 && BX / 2 >= B) BX /= 2;
  SY = kLmeThreads / BX;
}

}  // namespace vaemdl

using namespace vaemdl;

template <typename TIN>
static int lme_fwd_launch(const TIN* log_w, int S, long long B, float* out_b, void* stream) {
  if (!log_w || !out_b || S <= 0 || B <= 0) return VAEMDL_EINVAL;
  int BX, SY;
  lme_shape(B, BX, SY);
  const long long grid = (B + BX - 1) / BX;
  lme_kernel<false, TIN><<<static_cast<unsigned>(grid), kLmeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      log_w, nullptr, S, B, BX, SY, nullptr, out_b, nullptr, 1.0f);
  return cuda_rc(cudaGetLastError());
}
template <typename TIN>
static int lme_bwd_launch(const TIN* log_w, const float* g_out, int S, long long B, float* dlog_w, void* stream) {
  if (!log_w || !g_out || !dlog_w || S <= 0 || B <= 0) return VAEMDL_EINVAL;
  int BX, SY;
  lme_shape(B, BX, SY);
  const long long grid = (B + BX - 1) / BX;
  lme_bwd_kernel<TIN><<<static_cast<unsigned>(grid), kLmeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      log_w, g_out, S, B, BX, SY, dlog_w);
  return cuda_rc(cudaGetLastError());
}

extern "C" int vaemdl_logmeanexp_fwd(const float* log_w, int S, long long B, float* out_b, void* stream) {
  return lme_fwd_launch<float>(log_w, S, B, out_b, stream);
}
extern "C" int vaemdl_logmeanexp_fwd_f64(const double* log_w, int S, long long B, float* out_b, void* stream) {
  return lme_fwd_launch<double>(log_w, S, B, out_b, stream);
}
extern "C" int vaemdl_logmeanexp_bwd(const float* log_w, const float* g_out, int S, long long B, float* dlog_w,
                                     void* stream) {
  return lme_bwd_launch<float>(log_w, g_out, S, B, dlog_w, stream);
}
extern "C" int vaemdl_logmeanexp_bwd_f64(const double* log_w, const float* g_out, int S, long long B, float* dlog_w,
                                         void* stream) {
  return lme_bwd_launch<double>(log_w, g_out, S, B, dlog_w, stream);
}

namespace vaemdl {
// b_norm: the batch size the mean over b is taken over (differs from B when B is one chunk of a larger batch)
int iwae_tail_norm(const float* ll, const double* ll64, const float* extra, int S, long long B, float b_norm,
                   float* log_w, float* lme_b, float* g_ll, cudaStream_t st) {
  int BX, SY;
  lme_shape(B, BX, SY);
  const long long grid = (B + BX - 1) / BX;
  if (ll64)
    lme_kernel<true, double><<<static_cast<unsigned>(grid), kLmeThreads, 0, st>>>(ll64, extra, S, B, BX, SY, log_w, lme_b,
                                                                                  g_ll, b_norm);
  else
    lme_kernel<true, float><<<static_cast<unsigned>(grid), kLmeThreads, 0, st>>>(ll, extra, S, B, BX, SY, log_w, lme_b,
                                                                                 g_ll, b_norm);
  return cuda_rc(cudaGetLastError());
}
}  // namespace vaemdl

extern "C" int vaemdl_iwae_tail(const float* ll, const double* ll_f64, const float* extra, int S, long long B,
                                long long B_total, float* log_w, float* lme_b, float* elbo, float* g_ll, void* stream) {
  if ((!ll && !ll_f64) || S <= 0 || B <= 0 || B_total < 0) return VAEMDL_EINVAL;
  if (B_total == 0) B_total = B;
  if (elbo && !lme_b) return VAEMDL_EINVAL;  // the mean is taken over the lme_b buffer
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = iwae_tail_norm(ll, ll_f64, extra, S, B, static_cast<float>(B_total), log_w, lme_b, g_ll, st);
  if (rc) return rc;
  if (elbo) {
    mean_kernel<<<1, kLmeThreads, 0, st>>>(lme_b, B, static_cast<float>(B_total), elbo);
    rc = cuda_rc(cudaGetLastError());
  }
  return rc;
}

// ---- importance samples split across ranks (second row) ---------------------------------------------------------
// Each rank holds S_local of the S_total samples of every image.  The log-mean-exp over s (utils/utils.py:9-11) needs one
// exchange of a (max, sum exp) pair per image: split_local forms the rank's pairs, the caller all-gathers them
// ([world, 2, B] float64, rank order), split_combine turns them into the global log-mean-exp, the ELBO and the softmax
// weights of the LOCAL samples (the upstream gradient of the rank's slice).  Every rank computes bit-identical lme / elbo
// (pairs are combined in rank order, the batch mean in a fixed order).
namespace vaemdl {
__device__ __forceinline__ double split_log_w(const double* ll64, const float* extra, long long i) {
  return ll64[i] + (extra ? static_cast<double>(extra[i]) : 0.0);  // models/loss.py:34
}
__global__ void __launch_bounds__(kLmeThreads)
    split_local_kernel(const double* __restrict__ ll64, const float* __restrict__ extra, int S, long long B,
                       double* __restrict__ pair) {
  const long long b = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double mx = -INFINITY;
  for (int s = 0; s < S; ++s) mx = fmax(mx, split_log_w(ll64, extra, static_cast<long long>(s) * B + b));
  double sm = 0.0;
  for (int s = 0; s < S; ++s) sm += exp(split_log_w(ll64, extra, static_cast<long long>(s) * B + b) - mx);
  pair[b] = mx;
  pair[B + b] = sm;
}
// one block: thread t takes images t, t + 256, ...; the batch mean is a fixed-order tree over the block
__global__ void __launch_bounds__(kLmeThreads)
    split_combine_kernel(const double* __restrict__ ll64, const float* __restrict__ extra, int S, long long B,
                         const double* __restrict__ pairs, int world, double s_total, double b_norm,
                         float* __restrict__ log_w_out, float* __restrict__ lme_b, float* __restrict__ elbo,
                         float* __restrict__ g_ll) {
  __shared__ double red[kLmeThreads];
  double acc = 0.0;
  for (long long b = threadIdx.x; b < B; b += kLmeThreads) {
    double gmax = -INFINITY;
    for (int r = 0; r < world; ++r) gmax = fmax(gmax, pairs[(static_cast<long long>(r) * 2) * B + b]);
    double gsum = 0.0;
    for (int r = 0; r < world; ++r)
      gsum += pairs[(static_cast<long long>(r) * 2 + 1) * B + b] * exp(pairs[(static_cast<long long>(r) * 2) * B + b] - gmax);
    const double lme = gmax + log(gsum / s_total);  // utils/utils.py:11 over all S_total samples
    if (lme_b) lme_b[b] = static_cast<float>(lme);
    acc += lme;
    const double scale = -1.0 / (gsum * b_norm);  // d(-mean_b lme_b) / d log_w[s,b] = -softmax_s / B
    for (int s = 0; s < S; ++s) {
      const long long i = static_cast<long long>(s) * B + b;
      const double v = split_log_w(ll64, extra, i);
      if (log_w_out) log_w_out[i] = static_cast<float>(v);
      if (g_ll) g_ll[i] = static_cast<float>(exp(v - gmax) * scale);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kLmeThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0 && elbo) elbo[0] = static_cast<float>(red[0] / b_norm);  // models/loss.py:37
}
}  // namespace vaemdl

extern "C" int vaemdl_iwae_split_local(const double* ll_f64, const float* extra, int S_local, long long B, double* pair_out,
                                       void* stream) {
  if (!ll_f64 || !pair_out || S_local <= 0 || B <= 0) return VAEMDL_EINVAL;
  const long long grid = (B + kLmeThreads - 1) / kLmeThreads;
  split_local_kernel<<<static_cast<unsigned>(grid), kLmeThreads, 0, static_cast<cudaStream_t>(stream)>>>(ll_f64, extra, S_local,
                                                                                                       B, pair_out);
  return cuda_rc(cudaGetLastError());
}

extern "C" int vaemdl_iwae_split_combine(const double* ll_f64, const float* extra, int S_local, long long B,
                                         const double* pairs_all, int world, int S_total, long long B_total, float* log_w,
                                         float* lme_b, float* elbo, float* g_ll, void* stream) {
  if (!ll_f64 || !pairs_all || S_local <= 0 || B <= 0 || world <= 0 || S_total < S_local || B_total < 0) return VAEMDL_EINVAL;
  if (B_total == 0) B_total = B;
  split_combine_kernel<<<1, kLmeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      ll_f64, extra, S_local, B, pairs_all, world, static_cast<double>(S_total), static_cast<double>(B_total), log_w, lme_b,
      elbo, g_ll);
  return cuda_rc(cudaGetLastError());
}
