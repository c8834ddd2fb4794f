// iwae.cu -- log-mean-exp over the importance-sample axis and the fused IWAE tail.
//
// Replaces logmeanexp (utils/utils.py:9-11) and the tail of iwae_loss (models/loss.py:34-43; models/model06.py:47-55).
// The tensors are tiny ([S,B], S <= 5000): the goal is latency and a fixed summation order, not bandwidth.
#include "common.cuh"

namespace vaemdl {

constexpr int kLmeThreads = 256;

// Block = BX columns (b) x SY sample-lanes.  Two passes over the block's S values (max, then sum of exp), each
// combined across the SY lanes through shared memory in a fixed order.
template <bool TAIL>
__global__ void __launch_bounds__(kLmeThreads)
    lme_kernel(const float* __restrict__ in, const float* __restrict__ extra, int S, long long B, int BX, int SY,
               float* __restrict__ log_w_out, float* __restrict__ lme_b, float* __restrict__ g_ll, float b_norm) {
  __shared__ float red[kLmeThreads];
  const int bx = threadIdx.x % BX, sy = threadIdx.x / BX;
  const long long b = static_cast<long long>(blockIdx.x) * BX + bx;
  const bool ok = b < B;
  auto val = [&](int s) -> float {
    float v = in[static_cast<long long>(s) * B + b];
    if (TAIL && extra) v += extra[static_cast<long long>(s) * B + b];  // models/loss.py:34
    return v;
  };
  float mx = -INFINITY;
  if (ok)
    for (int s = sy; s < S; s += SY) mx = fmaxf(mx, val(s));
  red[threadIdx.x] = mx;
  __syncthreads();
  mx = red[bx];
  for (int j = 1; j < SY; ++j) mx = fmaxf(mx, red[j * BX + bx]);  // utils/utils.py:10
  __syncthreads();
  float sm = 0.0f;
  if (ok)
    for (int s = sy; s < S; s += SY) sm += expf(val(s) - mx);
  red[threadIdx.x] = sm;
  __syncthreads();
  sm = red[bx];
  for (int j = 1; j < SY; ++j) sm += red[j * BX + bx];
  if (!ok) return;
  if (sy == 0 && lme_b) lme_b[b] = logf(sm / static_cast<float>(S)) + mx;  // utils/utils.py:11
  if (TAIL) {
    const float scale = -1.0f / (sm * b_norm);  // d(-mean_b lme_b)/d log_w = -softmax_s / B
    for (int s = sy; s < S; s += SY) {
      const float v = val(s);
      if (log_w_out) log_w_out[static_cast<long long>(s) * B + b] = v;
      if (g_ll) g_ll[static_cast<long long>(s) * B + b] = expf(v - mx) * scale;
    }
  }
}

// dlog_w[s,b] = g_out[b] * softmax_s(log_w[:,b])
__global__ void __launch_bounds__(kLmeThreads)
    lme_bwd_kernel(const float* __restrict__ log_w, const float* __restrict__ g_out, int S, long long B, int BX, int SY,
                   float* __restrict__ dlog_w) {
  __shared__ float red[kLmeThreads];
  const int bx = threadIdx.x % BX, sy = threadIdx.x / BX;
  const long long b = static_cast<long long>(blockIdx.x) * BX + bx;
  const bool ok = b < B;
  float mx = -INFINITY;
  if (ok)
    for (int s = sy; s < S; s += SY) mx = fmaxf(mx, log_w[static_cast<long long>(s) * B + b]);
  red[threadIdx.x] = mx;
  __syncthreads();
  mx = red[bx];
  for (int j = 1; j < SY; ++j) mx = fmaxf(mx, red[j * BX + bx]);
  __syncthreads();
  float sm = 0.0f;
  if (ok)
    for (int s = sy; s < S; s += SY) sm += expf(log_w[static_cast<long long>(s) * B + b] - mx);
  red[threadIdx.x] = sm;
  __syncthreads();
  sm = red[bx];
  for (int j = 1; j < SY; ++j) sm += red[j * BX + bx];
  if (!ok) return;
  const float scale = g_out[b] / sm;
  for (int s = sy; s < S; s += SY)
    dlog_w[static_cast<long long>(s) * B + b] = expf(log_w[static_cast<long long>(s) * B + b] - mx) * scale;
}

// elbo = mean_b lme_b, single block, fixed order
__global__ void __launch_bounds__(kLmeThreads) mean_kernel(const float* __restrict__ v, long long B, float* __restrict__ out) {
  __shared__ float red[kLmeThreads];
  float acc = 0.0f;
  for (long long i = threadIdx.x; i < B; i += kLmeThreads) acc += v[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kLmeThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0] / static_cast<float>(B);  // models/loss.py:37
}

static void lme_shape(long long B, int& BX, int& SY) {
  BX = 32;
  while (BX > 1 && BX / 2 >= B) BX /= 2;
  SY = kLmeThreads / BX;
}

}  // namespace vaemdl

using namespace vaemdl;

extern "C" int vaemdl_logmeanexp_fwd(const float* log_w, int S, long long B, float* out_b, void* stream) {
  if (!log_w || !out_b || S <= 0 || B <= 0) return VAEMDL_EINVAL;
  int BX, SY;
  lme_shape(B, BX, SY);
  const long long grid = (B + BX - 1) / BX;
  lme_kernel<false><<<static_cast<unsigned>(grid), kLmeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      log_w, nullptr, S, B, BX, SY, nullptr, out_b, nullptr, 1.0f);
  return cuda_rc(cudaGetLastError());
}

extern "C" int vaemdl_logmeanexp_bwd(const float* log_w, const float* g_out, int S, long long B, float* dlog_w,
                                     void* stream) {
  if (!log_w || !g_out || !dlog_w || S <= 0 || B <= 0) return VAEMDL_EINVAL;
  int BX, SY;
  lme_shape(B, BX, SY);
  const long long grid = (B + BX - 1) / BX;
  lme_bwd_kernel<<<static_cast<unsigned>(grid), kLmeThreads, 0, static_cast<cudaStream_t>(stream)>>>(log_w, g_out, S, B,
                                                                                                    BX, SY, dlog_w);
  return cuda_rc(cudaGetLastError());
}

namespace vaemdl {
// b_norm: the batch size the mean over b is taken over (differs from B when B is one chunk of a larger batch)
int iwae_tail_norm(const float* ll, const float* extra, int S, long long B, float b_norm, float* log_w, float* lme_b,
                   float* g_ll, cudaStream_t st) {
  int BX, SY;
  lme_shape(B, BX, SY);
  const long long grid = (B + BX - 1) / BX;
  lme_kernel<true><<<static_cast<unsigned>(grid), kLmeThreads, 0, st>>>(ll, extra, S, B, BX, SY, log_w, lme_b, g_ll,
                                                                         b_norm);
  return cuda_rc(cudaGetLastError());
}
}  // namespace vaemdl

extern "C" int vaemdl_iwae_tail(const float* ll, const float* extra, int S, long long B, float* log_w, float* lme_b,
                                float* elbo, float* g_ll, void* stream) {
  if (!ll || S <= 0 || B <= 0) return VAEMDL_EINVAL;
  if (elbo && !lme_b) return VAEMDL_EINVAL;  // the mean is taken over the lme_b buffer
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = iwae_tail_norm(ll, extra, S, B, static_cast<float>(B), log_w, lme_b, g_ll, st);
  if (rc) return rc;
  if (elbo) {
    mean_kernel<<<1, kLmeThreads, 0, st>>>(lme_b, B, elbo);
    rc = cuda_rc(cudaGetLastError());
  }
  return rc;
}
