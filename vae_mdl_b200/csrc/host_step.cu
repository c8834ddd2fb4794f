// host_step.cu -- the IWAE observation-model step with HOST buffers (pipelined H2D / kernels / D2H).
//
// Replaces the loss + gradient of one train_step on the observation model (models/model05.py:139-148) for a caller
// whose decoder output lives in host memory.  The batch is cut into chunks of images; every importance sample of a
// chunk's images travels together, so the log-mean-exp over samples and its gradient stay chunk-local
// (models/loss.py:34-37 reduces over s for a fixed b).  Chunks rotate over kSlots device staging slots, each with
// its own stream: the H2D copy of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace vaemdl {

constexpr int kSlots = 3;

struct Slot {
  cudaStream_t stream = nullptr;
  float* params = nullptr;
  float* grads = nullptr;
  float* ll = nullptr;     // [S, cb]
  float* extra = nullptr;  // [S, cb]
  float* g_ll = nullptr;   // [S, cb]
  float* lme = nullptr;    // [cb]
  void* ws = nullptr;
  size_t params_bytes = 0, grads_bytes = 0, small_elems = 0, lme_elems = 0, ws_bytes = 0;
};

struct HostCtx {
  int device = -1;
  Slot slots[kSlots];
  uint8_t* x = nullptr;
  size_t x_bytes = 0;
  // whole-batch copies of the small tensors: ONE host transfer each per step instead of three tiny ones per chunk
  // (a PCIe copy of a few hundred bytes costs ~10 us of queue time, which adds up to ~1 ms over 30 chunks)
  float* extra_full = nullptr;  // [S, B]
  float* ll_full = nullptr;     // [S, B]
  float* lme_full = nullptr;    // [B]
  size_t extra_full_bytes = 0, ll_full_bytes = 0, lme_full_bytes = 0;
  cudaEvent_t x_ready = nullptr;
  std::mutex mu;  // one step at a time per device; callers on different devices do not wait for each other
};

static std::mutex g_mu;
static std::vector<HostCtx*> g_ctxs;

template <typename T>
static cudaError_t grow(T*& p, size_t& have, size_t want) {
  if (have >= want) return cudaSuccess;
  if (p) cudaFree(p);
  p = nullptr;
  have = 0;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want);
  if (e == cudaSuccess) have = want;
  return e;
}

// the per-device context (staging slots, streams); *err reports a failed stream / event creation
static HostCtx* get_ctx(cudaError_t* err) {
  *err = cudaSuccess;
  int dev = 0;
  if ((*err = cudaGetDevice(&dev)) != cudaSuccess) return nullptr;
  for (HostCtx* c : g_ctxs)
    if (c->device == dev) return c;
  HostCtx* c = new HostCtx();
  c->device = dev;
  for (auto& s : c->slots) {
    if (*err == cudaSuccess) *err = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
  }
  if (*err == cudaSuccess) *err = cudaEventCreateWithFlags(&c->x_ready, cudaEventDisableTiming);
  if (*err != cudaSuccess) {  // never hand out a context whose null streams would alias the legacy stream
    for (auto& s : c->slots)
      if (s.stream) cudaStreamDestroy(s.stream);
    if (c->x_ready) cudaEventDestroy(c->x_ready);
    delete c;
    return nullptr;
  }
  g_ctxs.push_back(c);
  return c;
}

// An error return must not leave copies in flight that still read / write the caller's host buffers.
static int drain(HostCtx* c, int rc) {
  for (auto& s : c->slots)
    if (s.stream) cudaStreamSynchronize(s.stream);
  cudaGetLastError();
  return rc;
}

}  // namespace vaemdl

using namespace vaemdl;

// from the first enqueued copy on: wait for everything in flight before reporting the error
#define VAEMDL_TRY_DRAIN(expr)                            \
  do {                                                    \
    cudaError_t e__ = (expr);                             \
    if (e__ != cudaSuccess) return drain(c, cuda_rc(e__)); \
  } while (0)

#define VAEMDL_TRY(expr)                       \
  do {                                         \
    cudaError_t e__ = (expr);                  \
    if (e__ != cudaSuccess) return cuda_rc(e__); \
  } while (0)

extern "C" int vaemdl_modl_iwae_step_host(const float* params_host, const uint8_t* x_host, const float* extra_host, int S,
                                          int B, int H, int W, int M, float* dparams_host, float* ll_host,
                                          float* lme_host, float* elbo_host, int chunk_b) {
  if (!params_host || !x_host || !ll_host || !lme_host || !elbo_host) return VAEMDL_EINVAL;
  if (S <= 0 || B <= 0 || H <= 0 || W <= 0) return VAEMDL_EINVAL;
  if (M < 1 || M > VAEMDL_MAX_MIX) return VAEMDL_EUNSUPPORTED;
  cudaError_t ctx_err = cudaSuccess;
  HostCtx* c;
  {
    std::lock_guard<std::mutex> lock(g_mu);  // the registry only
    c = get_ctx(&ctx_err);
  }
  if (!c) return cuda_rc(ctx_err);
  std::lock_guard<std::mutex> step_lock(c->mu);
  const size_t HW = static_cast<size_t>(H) * W;
  const size_t img_bytes = HW * 10 * M * sizeof(float);  // one (s,b) image of parameters
  if (chunk_b <= 0) {
    const size_t target = 64u << 20;  // ~50-64 MB of parameters per chunk (tools/tune_e2e.py: best on PCIe Gen5)
    chunk_b = static_cast<int>(target / (img_bytes * S));
    if (chunk_b < 1) chunk_b = 1;
  }
  if (chunk_b > B) chunk_b = B;
  const size_t chunk_param_bytes = img_bytes * S * chunk_b;

  VAEMDL_TRY(grow(c->x, c->x_bytes, static_cast<size_t>(B) * HW * 3));
  const size_t sb_bytes = static_cast<size_t>(S) * B * sizeof(float);
  VAEMDL_TRY(grow(c->extra_full, c->extra_full_bytes, sb_bytes));
  VAEMDL_TRY(grow(c->ll_full, c->ll_full_bytes, sb_bytes));
  VAEMDL_TRY(grow(c->lme_full, c->lme_full_bytes, static_cast<size_t>(B) * sizeof(float)));
  for (auto& s : c->slots) {
    VAEMDL_TRY(grow(s.params, s.params_bytes, chunk_param_bytes));
    if (dparams_host) VAEMDL_TRY(grow(s.grads, s.grads_bytes, chunk_param_bytes));
    const size_t small = static_cast<size_t>(S) * chunk_b * sizeof(float);
    if (s.small_elems < small) {
      if (s.ll) cudaFree(s.ll);
      if (s.extra) cudaFree(s.extra);
      if (s.g_ll) cudaFree(s.g_ll);
      s.ll = s.extra = s.g_ll = nullptr;
      s.small_elems = 0;
      VAEMDL_TRY(cudaMalloc(reinterpret_cast<void**>(&s.ll), small));
      VAEMDL_TRY(cudaMalloc(reinterpret_cast<void**>(&s.extra), small));
      VAEMDL_TRY(cudaMalloc(reinterpret_cast<void**>(&s.g_ll), small));
      s.small_elems = small;
    }
    VAEMDL_TRY(grow(s.lme, s.lme_elems, static_cast<size_t>(chunk_b) * sizeof(float)));
    VAEMDL_TRY(grow(s.ws, s.ws_bytes, vaemdl_modl_workspace_bytes(static_cast<long long>(S) * chunk_b, H, W)));
  }

  // the observed images: one small copy, every slot stream waits for it
  VAEMDL_TRY_DRAIN(cudaMemcpyAsync(c->x, x_host, static_cast<size_t>(B) * HW * 3, cudaMemcpyHostToDevice, c->slots[0].stream));
  if (extra_host)
    VAEMDL_TRY_DRAIN(cudaMemcpyAsync(c->extra_full, extra_host, sb_bytes, cudaMemcpyHostToDevice, c->slots[0].stream));
  VAEMDL_TRY_DRAIN(cudaEventRecord(c->x_ready, c->slots[0].stream));
  for (int k = 1; k < kSlots; ++k) VAEMDL_TRY_DRAIN(cudaStreamWaitEvent(c->slots[k].stream, c->x_ready, 0));

  const size_t host_pitch = img_bytes * B;  // bytes between consecutive s in the host tensors
  // Chunk schedule: the first chunk's upload and the last chunk's download have nothing to overlap with, so the two
  // ends of the batch travel in half-size chunks.
  const int edge_b = chunk_b >= 2 ? chunk_b / 2 : chunk_b;
  int ci = 0;
  for (int b0 = 0; b0 < B; ++ci) {
    Slot& s = c->slots[ci % kSlots];
    int want = chunk_b;
    if (b0 == 0 || B - b0 <= chunk_b) want = edge_b;  // first chunk, and the tail split in two
    const int cb = (B - b0) < want ? (B - b0) : want;
    const size_t width = img_bytes * cb;
    const long long n_img = static_cast<long long>(S) * cb;
    // [S, B, ...] host  ->  [S, cb, ...] device: S rows of `width` bytes
    VAEMDL_TRY_DRAIN(cudaMemcpy2DAsync(s.params, width, reinterpret_cast<const char*>(params_host) + img_bytes * b0, host_pitch,
                                 width, S, cudaMemcpyHostToDevice, s.stream));
    if (extra_host)  // [S, B] -> the chunk's dense [S, cb], device to device
      VAEMDL_TRY_DRAIN(cudaMemcpy2DAsync(s.extra, cb * sizeof(float), c->extra_full + b0, static_cast<size_t>(B) * sizeof(float),
                                   cb * sizeof(float), S, cudaMemcpyDeviceToDevice, s.stream));
    // forward + fused finish (per-image sums, log-mean-exp, softmax weights normalised by the WHOLE batch): 2 launches
    int rc = vaemdl_modl_iwae_fwd(s.params, c->x + static_cast<size_t>(b0) * HW * 3, VAEMDL_X_U8, VAEMDL_RANGE_UNIT,
                                  VAEMDL_EDGE_MDL, S, cb, B, cb, H, W, M, extra_host ? s.extra : nullptr, s.ll, nullptr,
                                  nullptr, s.lme, nullptr, dparams_host ? s.g_ll : nullptr, s.ws, s.ws_bytes, s.stream);
    if (rc) return drain(c, rc);
    VAEMDL_TRY_DRAIN(cudaMemcpy2DAsync(c->ll_full + b0, static_cast<size_t>(B) * sizeof(float), s.ll, cb * sizeof(float),
                                 cb * sizeof(float), S, cudaMemcpyDeviceToDevice, s.stream));
    VAEMDL_TRY_DRAIN(cudaMemcpyAsync(c->lme_full + b0, s.lme, cb * sizeof(float), cudaMemcpyDeviceToDevice, s.stream));
    if (dparams_host) {
      rc = vaemdl_modl_bwd(s.params, c->x + static_cast<size_t>(b0) * HW * 3, VAEMDL_X_U8, VAEMDL_RANGE_UNIT,
                           VAEMDL_EDGE_MDL, n_img, cb, H, W, M, s.g_ll, nullptr, s.grads, s.stream);
      if (rc) return drain(c, rc);
      VAEMDL_TRY_DRAIN(cudaMemcpy2DAsync(reinterpret_cast<char*>(dparams_host) + img_bytes * b0, host_pitch, s.grads, width,
                                   width, S, cudaMemcpyDeviceToHost, s.stream));
    }
    b0 += cb;
  }
  for (auto& s : c->slots) VAEMDL_TRY_DRAIN(cudaStreamSynchronize(s.stream));
  VAEMDL_TRY_DRAIN(cudaMemcpyAsync(ll_host, c->ll_full, sb_bytes, cudaMemcpyDeviceToHost, c->slots[0].stream));
  VAEMDL_TRY_DRAIN(cudaMemcpyAsync(lme_host, c->lme_full, static_cast<size_t>(B) * sizeof(float), cudaMemcpyDeviceToHost,
                             c->slots[0].stream));
  VAEMDL_TRY_DRAIN(cudaStreamSynchronize(c->slots[0].stream));
  double acc = 0.0;
  for (int b = 0; b < B; ++b) acc += static_cast<double>(lme_host[b]);  // models/loss.py:37 (mean over the batch)
  elbo_host[0] = static_cast<float>(acc / B);
  return VAEMDL_OK;
}

extern "C" void vaemdl_host_release(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (HostCtx* c : g_ctxs) {
    { std::lock_guard<std::mutex> wait_for_step(c->mu); }  // a step in flight on that device finishes first
    cudaSetDevice(c->device);
    for (auto& s : c->slots) {
      if (s.stream) cudaStreamSynchronize(s.stream);
      cudaFree(s.params);
      cudaFree(s.grads);
      cudaFree(s.ll);
      cudaFree(s.extra);
      cudaFree(s.g_ll);
      cudaFree(s.lme);
      cudaFree(s.ws);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
    cudaFree(c->x);
    cudaFree(c->extra_full);
    cudaFree(c->ll_full);
    cudaFree(c->lme_full);
    if (c->x_ready) cudaEventDestroy(c->x_ready);
    delete c;
  }
  g_ctxs.clear();
  cudaSetDevice(cur);
}
