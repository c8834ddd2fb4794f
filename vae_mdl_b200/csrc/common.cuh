// common.cuh -- PTX wrappers (mbarrier, 1-D TMA bulk copies, MUFU approximations) for sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdint.h>

#include "../../include/vaemdl.h"

namespace vaemdl {

constexpr int kWarp = 32;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr unsigned kFull = 0xffffffffu;

// ---- special-function unit (MUFU) approximations, flush-to-zero -----------------------------
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2a(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpa(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- uint8 pixel -> float32 / 255 (utils/data.py:15-16), bit-exact in three FMA-pipe instructions ---------------
// q0 = k * fl(1/255) is off by one ulp for 126 of the 256 byte values; one residual step (rem = k - 255*q0 exactly, by
// FMA) gives the correctly rounded quotient for all 256 (tests/test_host_logic.py enumerates them against k / 255.f).
__device__ __forceinline__ float u8_to_unit(unsigned k) {
  const float kf = static_cast<float>(k);
  const float r = 1.0f / 255.0f;  // compile-time constant, correctly rounded
  const float q0 = __fmul_rn(kf, r);
  const float rem = __fmaf_rn(-q0, 255.0f, kf);
  return __fmaf_rn(rem, r, q0);
}

// ---- shared-memory addresses -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  // make the initialised barrier visible to the async (TMA) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- L2 cache policies -------------------------------------------------------------------------
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---- 1-D bulk copies (TMA engine, SASS UBLKCP) ----------------------------------------------------
// global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, both addresses 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* gmem_dst, const void* smem_src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// A kernel launched with launch_pdl() may be scheduled while its predecessor in the stream is still draining; it must
// execute pdl_wait() before touching anything the predecessor wrote (it blocks until the predecessor has completed and
// its memory is visible).  pdl_trigger() in the predecessor allows the dependent launch to begin early.  Both are no-ops
// for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- host-side helpers ---------------------------------------------------------------------------
inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : static_cast<int>(e); }

// launch `kern<<<grid, block, smem, st>>>(arg)` with the programmatic-stream-serialization attribute (VAEMDL_PDL=0: off)
template <typename Arg>
inline cudaError_t launch_pdl(void (*kern)(Arg), unsigned grid, unsigned block, size_t smem, cudaStream_t st, const Arg& arg) {
  static const bool on = [] {
    const char* e = getenv("VAEMDL_PDL");
    return !(e && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, arg);
}

struct DeviceInfo {
  int sm_count;
  int max_smem_optin;
};
const DeviceInfo& device_info();  // cached per current device (api_misc.cu)

// ---- per-image sums of a forward kernel whose warps own runs of consecutive tiles (finish.cu) -------------------------------
// Warp w of the forward grid owned tiles [w*tw_base + min(w, tw_rem), + tw_base + (w < tw_rem)) of PPT rows each and left
// one float64 partial per image its run touched, at partial[n*K + (w - first warp whose run touches image n)]: the
// partials of one image are contiguous and in warp order.  HW = rows per image.
constexpr long long kMaxGridWarps = 8192;  // bound on (CTAs x warps) of such a grid: sizes the partial-sum workspace
// n_img * K <= 2 * n_img + 6 * (grid warps) for every plan the launchers can pick (K = tpi / tw_base + 2 with
// tpi <= 3 * HW / PPT and tw_base >= tiles / (2 * warps)); the launchers check the plan against this before enqueuing
inline size_t partial_elems(long long n_img) { return 3 * static_cast<size_t>(n_img) + 6 * kMaxGridWarps + 64; }
inline bool partials_fit(long long n_img, int K) { return static_cast<size_t>(n_img) * static_cast<size_t>(K) <= partial_elems(n_img); }
// K: the largest number of warp runs one image can intersect (runs are tw_base or tw_base + 1 tiles long)
inline int partial_K(long long HW, int PPT, long long tw_base) {
  const long long tpi = (HW + PPT - 1) / PPT + 1;  // tiles an image can overlap
  return static_cast<int>((tw_base >= 1 ? tpi / tw_base : tpi) + 2);
}
struct PartialGeom {
  const double* partial;
  long long tw_base, tw_rem;
  int K, PPT;
  long long HW;
};
template <typename I>
__device__ __forceinline__ I run_of_tile(I t, I base, I rem) {  // which warp's run holds tile t
  const I cut = rem * (base + 1);
  return t < cut ? t / (base + 1) : rem + (t - cut) / base;
}
__device__ __forceinline__ long long partial_slot(long long n, long long w, long long HW, int PPT, long long base,
                                                  long long rem, int K, bool small) {
  long long w_lo;
  if (small)
    w_lo = run_of_tile<unsigned>(static_cast<unsigned>(n * HW) / static_cast<unsigned>(PPT), static_cast<unsigned>(base),
                                 static_cast<unsigned>(rem));
  else
    w_lo = run_of_tile<long long>((n * HW) / PPT, base, rem);
  return n * K + (w - w_lo);
}
// image n's sum: the partials of the forward warps whose tile runs touched it -- contiguous, added in warp order
template <typename I>
__device__ __forceinline__ double image_sum_t(const PartialGeom& g, long long n64) {
  const I n = static_cast<I>(n64), HW = static_cast<I>(g.HW), PPT = static_cast<I>(g.PPT);
  const I base = static_cast<I>(g.tw_base), rem = static_cast<I>(g.tw_rem);
  const I first = n * HW, last = first + HW - 1;
  const I w_lo = run_of_tile<I>(first / PPT, base, rem), w_hi = run_of_tile<I>(last / PPT, base, rem);
  const int cnt = static_cast<int>(w_hi - w_lo) + 1;
  const double* p = g.partial + n64 * g.K;
  double acc = 0.0;
  for (int j0 = 0; j0 < cnt; j0 += 8) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (j0 + j < cnt) ? p[j0 + j] : 0.0;  // all loads first, then the adds (fixed order)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  }
  return acc;
}
__device__ __forceinline__ double image_sum(const PartialGeom& g, long long n, bool small) {
  return small ? image_sum_t<unsigned>(g, n) : image_sum_t<long long>(g, n);
}

// ---- the ELBO share of a batch shard, published to every rank of the box over peer memory (NVLink P2P stores) ---------------
// With the batch split across ranks each rank's step ends with its additive share of the ELBO,
// sum_b lme_b / B_total (models/loss.py:37); the global value is the sum of the shares.  Instead of a separate collective
// launch, the kernel that forms the share stores it -- one self-describing 8-byte word {sequence number, value} -- into slot
// [seq % ring][rank] of EVERY rank's exchange buffer (the peers' buffers are mapped through CUDA IPC).  A reader sums the
// `n` words of step `seq` in rank order (vaemdl_peer_elbo_sum): bit-identical on every rank.
struct PeerOut {
  unsigned long long* slots[VAEMDL_MAX_PEERS];  // slots[r]: rank r's buffer, [ring][n] words, as mapped in THIS process
  int n = 0;                                    // ranks (0: nothing to publish)
  int rank = 0;
  int ring = 1;
  unsigned seq = 0;
};
__device__ __forceinline__ void peer_publish(const PeerOut& p, float value) {  // one thread
  if (p.n <= 0) return;
  const unsigned long long word = (static_cast<unsigned long long>(p.seq) << 32) | __float_as_uint(value);
  const size_t at = static_cast<size_t>(p.seq % static_cast<unsigned>(p.ring)) * p.n + p.rank;
  for (int r = 0; r < p.n; ++r) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p.slots[r] + at), "l"(word) : "memory");
  }
}
// host: the exchange attached to the next ELBO-producing call of this host thread (vaemdl_peer_next), consumed by it
PeerOut take_peer();
void give_peer(const PeerOut& p);  // hand it back (a one-launch step that falls back to separate launches)
// fallback for the routes whose ELBO comes out of vaemdl_iwae_tail: one tiny launch that publishes elbo[0]
int peer_push(const PeerOut& p, const float* elbo, cudaStream_t st);

// ---- the IWAE finish inside a cooperative one-launch step (modl_step_kernel, dl_step_kernel) --------------------------------
struct StepFinish {
  PartialGeom geom;
  const float* extra;  // [S,B] nullable
  float* ll;           // [S,B] nullable
  double* ll64;        // [S,B] nullable
  float* log_w;        // [S,B] nullable
  float* lme_b;        // [B]
  float* elbo;         // [1] nullable
  float* g_ll;         // [S,B]
  double* lme64;       // [B] scratch
  long long B;
  int S;  // <= 32: one importance sample per lane
  float b_norm;
  bool small;
  PeerOut peer;  // where the ELBO share also goes (n = 0: nowhere)
};
// warp `gw` of `total_warps` takes batch elements gw, gw + total_warps, ...: same arithmetic, in the same order, as
// finish_kernel steps (1) and (2) with S <= 32
__device__ __forceinline__ void step_finish(const StepFinish& f, long long gw, long long total_warps, int lane) {
  for (long long b = gw; b < f.B; b += total_warps) {
    const long long n = static_cast<long long>(lane) * f.B + b;
    double v = -INFINITY;
    if (lane < f.S) {
      const double acc = image_sum(f.geom, n, f.small);
      if (f.ll) f.ll[n] = static_cast<float>(acc);
      if (f.ll64) f.ll64[n] = acc;
      v = acc + (f.extra ? static_cast<double>(f.extra[n]) : 0.0);  // models/loss.py:34
      if (f.log_w) f.log_w[n] = static_cast<float>(v);
    }
    double mx = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(kFull, mx, o));  // utils/utils.py:10
    const float e = lane < f.S ? expf(static_cast<float>(v - mx)) : 0.0f;
    float sm = e;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(kFull, sm, o);
    const double lme = static_cast<double>(logf(sm / static_cast<float>(f.S))) + mx;  // utils/utils.py:11
    if (lane == 0) {
      f.lme_b[b] = static_cast<float>(lme);
      f.lme64[b] = lme;
    }
    if (lane < f.S) f.g_ll[n] = e * (-1.0f / (sm * f.b_norm));  // d(-mean_b lme_b)/d log_w = -softmax_s / B
  }
}

struct IwaeOut {  // outputs of the fused IWAE finish (all nullable); active when S > 0
  int S = 0;
  long long B = 0, B_total = 0;
  const float* extra = nullptr;
  float *log_w = nullptr, *lme_b = nullptr, *elbo = nullptr, *g_ll = nullptr;
  PeerOut peer;
};
// ll / ll64 [n_img] nullable.  iw.S == 0: one reduction launch.  Otherwise also log_w, log-mean-exp, elbo and
// g_ll = d(-elbo)/d ll: ONE fused launch when S <= 512, else reduction + IWAE tail.  scratch: n_img doubles;
// counter: a zeroed word (the forward kernel clears it).
int finish_partials(const PartialGeom& g, long long n_img, float* ll, double* ll64, const IwaeOut& iw, double* scratch,
                    unsigned* counter, cudaStream_t st);

// iwae.cu: fused IWAE tail with an explicit batch normaliser
int iwae_tail_norm(const float* ll, const double* ll64, const float* extra, int S, long long B, float b_norm,
                   float* log_w, float* lme_b, float* g_ll, cudaStream_t st);

}  // namespace vaemdl
