// dlogistic.cu -- plain (single-component, per-sub-pixel) discretized logistic: log-prob, gradient, sampler.
//
// Replaces DiscretizedLogistic.log_prob / .sample (utils/discretized_logistic.py:35-85) and its autodiff
// (models/model03.py:141-147, models/model06.py).  Same branch structure as the mixture kernels (modl_math.cuh) but
// one logarithm per element is unavoidable here, so the result is assembled in the log domain:
//   normal : -|mid| + log((1-G)(1+G)) - log((A+G)(1+AG))
//   low    : -|mid| - ls + log(width) - 2 log(1+A)                  (utils/discretized_logistic.py:27-33)
//   edges  : -log(1+AG)   or   -|mid| - log(A+G)
// The log-scale is NOT clamped in this class (utils/discretized_logistic.py:38); exp(-ls) is saturated at 3e38 so
// that x == loc with a vanishing scale gives log 1 = 0 as the reference does instead of inf*0.
#include "modl_math.cuh"

namespace vaemdl {

struct DlArgs {
  const float* loc;
  const float* logscale;
  const void* x;
  float* lp_elem;
  double* partial;    // [total_warps][K] float64: one partial per (warp, image its run of tiles touches)
  unsigned* zero_me;  // nullable: a word the forward kernel clears for the fused finish kernel that follows it
  long long tw_base, tw_rem;  // warp w owns tiles [w*tw_base + min(w, tw_rem), + tw_base + (w < tw_rem))
  int K;
  double* ll_atomic;  // [n_img] float64 accumulators
  const float* g_image;
  const float* g_elem;
  float* dloc;
  float* dls;
  long long n_rows;       // n_img * rows_per_img
  long long rows_per_img; // D / CPT
  int x_batch;
  int x_u8;
  int C;       // channels of the [.., C] tensors
  int ld;      // channel-row stride of loc/logscale
  int ld_out;  // channel-row stride of dloc/dls
  float low, high, dx, width, ln_width;
};

struct DlOut {
  float lp, dloc, dls;
};

template <bool BWD>
__device__ __forceinline__ DlOut dl_elem(float x, float loc, float ls, const DlArgs& a) {
  const bool left = x <= a.low;    // utils/discretized_logistic.py:71-73
  const bool right = x >= a.high;  // :74-76
  const float inv = fminf(ex2a(-ls * kLog2e), 3.0e38f);
  const float mid = inv * (x - loc);
  const float am = fabsf(mid);
  const float A = ex2a(-am * kLog2e);
  const float h = inv * a.dx;
  float G, omG;
  exp_neg_h(h, G, omG);
  const float AG = A * G, ApG = A + G, opAG = 1.0f + AG, opA = 1.0f + A;
  const bool pos = mid >= 0.0f;
  const float rest_n = omG * (1.0f + G);
  const float den_n = ApG * opAG;
  const bool is_norm = A * rest_n > 1e-5f * den_n;  // prob > 1e-5 (:64)
  const bool edge = left || right;
  const bool one_over = (left == pos);
  // lp = -use_mid*|mid| + cst + ln2*(lg2(nu) - lg2(de))
  float nu = is_norm ? rest_n : 1.0f;
  float de = is_norm ? den_n : opA * opA;
  float cst = is_norm ? 0.0f : (a.ln_width - ls);
  float use_mid = am;
  if (edge) {
    nu = 1.0f;
    de = one_over ? opAG : ApG;
    cst = 0.0f;
    use_mid = one_over ? 0.0f : am;
  }
  DlOut o;
  o.lp = cst - use_mid + kLn2 * (lg2a(nu) - lg2a(de));
  if constexpr (BWD) {
    const float omA2 = (1.0f - A) * opA;
    const float sgn = pos ? 1.0f : -1.0f;
    const float hc = h_coth_h(h, G, omG);
    float den = is_norm ? den_n : opA * opA;
    float nm = is_norm ? -sgn * G * omA2 : -sgn * omA2;
    float nh = is_norm ? -h * A * rest_n : 0.0f;
    float c0 = is_norm ? hc : 0.0f;
    float dir = is_norm ? 0.0f : -1.0f;
    if (edge) {
      const float t = one_over ? AG : G;
      den = one_over ? opAG : ApG;
      nm = left ? t : -t;
      nh = h * t;
      c0 = 0.0f;
      dir = 0.0f;
    }
    const float rden = rcpa(den);
    const float Dm = nm * rden;
    o.dloc = -inv * Dm;
    o.dls = (dir - c0) - fmaf(mid, Dm, nh * rden);
  }
  return o;
}

__device__ __forceinline__ double dl_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// One thread per row of CPT channels (CPT = 3 for images, 1 for anything else); one warp = 32 consecutive rows.
template <int CPT, bool BWD>
__global__ void __launch_bounds__(256) dl_kernel(const DlArgs a) {
  const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (!BWD && a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
  // a run of consecutive 32-row tiles per warp: per-image sums stay in registers until the warp leaves the image
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);
  double acc0 = 0.0, acc1 = 0.0;
  const long long n_warp_first = (t_begin * 32) / a.rows_per_img;
  long long n_base = n_warp_first;
  for (long long t = t_begin; t < t_end; ++t) {
    const long long r_raw = t * 32 + lane;
    const bool active = r_raw < a.n_rows;
    const long long r = active ? r_raw : t * 32;
    const long long n = r / a.rows_per_img;
    const long long rr = r - n * a.rows_per_img;
    const long long xb = a.x_batch == 1 ? 0 : n % a.x_batch;
    const long long xo = (xb * a.rows_per_img + rr) * CPT;
    float g = 0.0f;
    if constexpr (BWD) {
      if (a.g_image) g = a.g_image[n];
    }
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const long long e = r * CPT + c;  // flat element index
      long long po, qo;                 // offsets into loc/logscale and into dloc/dls
      if (CPT == 3 && a.C == 3) {
        po = r * a.ld + c;
        qo = r * a.ld_out + c;
      } else {
        const long long row = e / a.C;
        const int col = static_cast<int>(e - row * a.C);
        po = row * a.ld + col;
        qo = row * a.ld_out + col;
      }
      float xv;
      if (a.x_u8)
        xv = __fdiv_rn(static_cast<float>(static_cast<const uint8_t*>(a.x)[xo + c]), 255.0f);
      else
        xv = static_cast<const float*>(a.x)[xo + c];
      const DlOut o = dl_elem<BWD>(xv, a.loc[po], a.logscale[po], a);
      if constexpr (!BWD) {
        if (a.lp_elem && active) a.lp_elem[e] = o.lp;
        acc += o.lp;
      } else {
        float ge = g;
        if (a.g_elem) ge += a.g_elem[e];
        if (active) {
          a.dloc[qo] = ge * o.dloc;
          a.dls[qo] = ge * o.dls;
        }
      }
    }
    if constexpr (!BWD) {
      const float val = active ? acc : 0.0f;
      if (a.partial) {
        // a tile holds rows of at most two images (rows_per_img >= 32 on this route)
        const long long n_first = __shfl_sync(kFull, n, 0);
        while (n_base < n_first) {
          const double done = dl_warp_sum(acc0);
          if (lane == 0) a.partial[gw * a.K + (n_base - n_warp_first)] = done;
          acc0 = acc1;
          acc1 = 0.0;
          ++n_base;
        }
        if (n == n_base)
          acc0 += static_cast<double>(val);
        else
          acc1 += static_cast<double>(val);
      } else if (a.ll_atomic && active) {
        atomicAdd(a.ll_atomic + n, static_cast<double>(val));
      }
    }
  }
  if constexpr (!BWD) {
    if (a.partial && t_begin < t_end) {
      const long long n_last = (t_end * 32 < a.n_rows ? t_end * 32 - 1 : a.n_rows - 1) / a.rows_per_img;
      const double d0 = dl_warp_sum(acc0), d1 = dl_warp_sum(acc1);
      if (lane == 0) {
        a.partial[gw * a.K + (n_base - n_warp_first)] = d0;
        if (n_base + 1 <= n_last) a.partial[gw * a.K + (n_base + 1 - n_warp_first)] = d1;
      }
    }
  }
}

__global__ void dl_cast_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]);
}

// sampler: clip(loc + exp(ls) * (log u - log(1-u)), low, high) in float64   (utils/discretized_logistic.py:80-85)
__global__ void dl_sample_kernel(const float* __restrict__ loc, const float* __restrict__ logscale, int C, int ld,
                                 const float* __restrict__ u, long long n_elem, float low, float high,
                                 float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n_elem; e += stride) {
    const long long row = e / C;
    const long long po = row * ld + (e - row * C);
    const double uu = static_cast<double>(u[e]);
    const double eps = log(uu) - log(1.0 - uu);
    double v = static_cast<double>(loc[po]) + exp(static_cast<double>(logscale[po])) * eps;
    v = fmin(fmax(v, static_cast<double>(low)), static_cast<double>(high));
    out[e] = static_cast<float>(v);
  }
}

static int dl_fill(DlArgs& a, const float* loc, const float* logscale, int C, int ld, const void* x, int x_dtype,
                   long long n_img, int x_batch, long long D, float low, float high, float levels, int& cpt) {
  if (!loc || !logscale || !x) return VAEMDL_EINVAL;
  if (C <= 0 || ld < C || n_img <= 0 || x_batch <= 0 || D <= 0) return VAEMDL_EINVAL;
  if (x_dtype != VAEMDL_X_F32 && x_dtype != VAEMDL_X_U8) return VAEMDL_EINVAL;
  if (!(levels > 1.0f) || !(high > low)) return VAEMDL_EINVAL;
  cpt = (C == 3 && D % 3 == 0) ? 3 : 1;
  a.loc = loc;
  a.logscale = logscale;
  a.x = x;
  a.rows_per_img = D / cpt;
  a.n_rows = n_img * a.rows_per_img;
  a.x_batch = x_batch;
  a.x_u8 = x_dtype == VAEMDL_X_U8;
  a.C = C;
  a.ld = ld;
  a.ld_out = C;
  a.low = low;
  a.high = high;
  const double width = (static_cast<double>(high) - static_cast<double>(low)) / (static_cast<double>(levels) - 1.0);
  a.width = static_cast<float>(width);      // utils/discretized_logistic.py:18
  a.dx = static_cast<float>(width / 2.0);   // :21
  a.ln_width = static_cast<float>(log(width));
  return VAEMDL_OK;
}

template <bool BWD>
static int dl_launch(DlArgs a, int cpt, cudaStream_t st, PartialGeom* geom = nullptr) {
  const DeviceInfo& di = device_info();
  const long long n_tiles = (a.n_rows + 31) / 32;
  long long blocks = (n_tiles + 7) / 8;
  long long cap = static_cast<long long>(di.sm_count) * 8;
  if (cap * 8 > kMaxGridWarps) cap = kMaxGridWarps / 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const long long total_warps = blocks * 8;
  a.tw_base = n_tiles / total_warps;
  a.tw_rem = n_tiles % total_warps;
  a.K = static_cast<int>(((a.tw_base + (a.tw_rem ? 1 : 0)) * 32 + a.rows_per_img - 1) / a.rows_per_img + 1);
  if (geom) *geom = PartialGeom{a.partial, a.tw_base, a.tw_rem, a.K, 32, a.rows_per_img};
  if (cpt == 3)
    dl_kernel<3, BWD><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a);
  else
    dl_kernel<1, BWD><<<static_cast<unsigned>(blocks), 256, 0, st>>>(a);
  return cuda_rc(cudaGetLastError());
}

}  // namespace vaemdl

using namespace vaemdl;

extern "C" size_t vaemdl_dlogistic_workspace_bytes(long long n_img, long long D) {
  if (n_img <= 0 || D <= 0) return 0;
  // per-(warp, image) partials (also covers the atomic route's one accumulator per image) + the finish kernel's
  // block sums / per-image scratch + the arrival counter
  return partial_elems(n_img) * sizeof(double) + static_cast<size_t>(n_img) * sizeof(double) + 256;
}

namespace vaemdl {
static int dl_fwd_impl(const float* loc, const float* logscale, int C, int ld, const void* x, int x_dtype,
                       long long n_img, int x_batch, long long D, float low, float high, float levels, float* lp_elem,
                       float* ll_image, double* ll_image_f64, const IwaeOut& iw, void* workspace, size_t workspace_bytes,
                       cudaStream_t st) {
  DlArgs a{};
  int cpt = 1;
  int rc = dl_fill(a, loc, logscale, C, ld, x, x_dtype, n_img, x_batch, D, low, high, levels, cpt);
  if (rc) return rc;
  const bool iwae = iw.S > 0;
  if (iwae && static_cast<long long>(iw.S) * iw.B != n_img) return VAEMDL_EINVAL;
  const bool want_ll = ll_image || ll_image_f64 || iwae;
  if (!lp_elem && !want_ll) return VAEMDL_EINVAL;
  a.lp_elem = lp_elem;
  const bool use_partials = want_ll && a.rows_per_img >= 32;
  char* ws = static_cast<char*>(workspace);
  const size_t tail_off = partial_elems(n_img) * sizeof(double);
  unsigned* counter = nullptr;
  if (want_ll) {
    if (!workspace || workspace_bytes < vaemdl_dlogistic_workspace_bytes(n_img, D)) return VAEMDL_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 7u) return VAEMDL_EALIGN;
    counter = reinterpret_cast<unsigned*>(ws + tail_off + static_cast<size_t>(n_img) * sizeof(double));
    if (use_partials) {
      a.partial = reinterpret_cast<double*>(ws);
      if (iwae && iw.elbo) a.zero_me = counter;
    } else {
      a.ll_atomic = ll_image_f64 ? ll_image_f64 : reinterpret_cast<double*>(ws);
      cudaError_t e = cudaMemsetAsync(a.ll_atomic, 0, sizeof(double) * n_img, st);
      if (e != cudaSuccess) return cuda_rc(e);
    }
  }
  PartialGeom geom{};
  rc = dl_launch<false>(a, cpt, st, &geom);
  if (rc) return rc;
  if (use_partials)
    return finish_partials(geom, n_img, ll_image, ll_image_f64, iw, reinterpret_cast<double*>(ws + tail_off), counter, st);
  if (!want_ll) return VAEMDL_OK;
  if (ll_image) {
    dl_cast_kernel<<<static_cast<unsigned>((n_img + 255) / 256), 256, 0, st>>>(a.ll_atomic, ll_image, n_img);
    rc = cuda_rc(cudaGetLastError());
  }
  if (rc || !iwae) return rc;
  return vaemdl_iwae_tail(nullptr, a.ll_atomic, iw.extra, iw.S, iw.B, iw.B_total, iw.log_w, iw.lme_b, iw.elbo, iw.g_ll, st);
}
}  // namespace vaemdl

extern "C" int vaemdl_dlogistic_fwd(const float* loc, const float* logscale, int C, int ld, const void* x, int x_dtype,
                                    long long n_img, int x_batch, long long D, float low, float high, float levels,
                                    float* lp_elem, float* ll_image, double* ll_image_f64, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  return dl_fwd_impl(loc, logscale, C, ld, x, x_dtype, n_img, x_batch, D, low, high, levels, lp_elem, ll_image,
                     ll_image_f64, IwaeOut{}, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int vaemdl_dlogistic_iwae_fwd(const float* loc, const float* logscale, int C, int ld, const void* x,
                                         int x_dtype, int S, long long B, long long B_total, int x_batch, long long D,
                                         float low, float high, float levels, const float* extra, float* ll_image,
                                         double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  if (S <= 0 || B <= 0 || B_total < 0) return VAEMDL_EINVAL;
  if (elbo && !lme_b) return VAEMDL_EINVAL;
  IwaeOut iw;
  iw.S = S;
  iw.B = B;
  iw.B_total = B_total;
  iw.extra = extra;
  iw.log_w = log_w;
  iw.lme_b = lme_b;
  iw.elbo = elbo;
  iw.g_ll = g_ll;
  return dl_fwd_impl(loc, logscale, C, ld, x, x_dtype, static_cast<long long>(S) * B, x_batch, D, low, high, levels,
                     nullptr, ll_image, ll_image_f64, iw, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int vaemdl_dlogistic_bwd(const float* loc, const float* logscale, int C, int ld, const void* x, int x_dtype,
                                    long long n_img, int x_batch, long long D, float low, float high, float levels,
                                    const float* g_image, const float* g_elem, float* dloc, float* dlogscale, int ld_out,
                                    void* stream) {
  DlArgs a{};
  int cpt = 1;
  int rc = dl_fill(a, loc, logscale, C, ld, x, x_dtype, n_img, x_batch, D, low, high, levels, cpt);
  if (rc) return rc;
  if (!dloc || !dlogscale || (!g_image && !g_elem) || ld_out < C) return VAEMDL_EINVAL;
  a.g_image = g_image;
  a.g_elem = g_elem;
  a.dloc = dloc;
  a.dls = dlogscale;
  a.ld_out = ld_out;
  return dl_launch<true>(a, cpt, static_cast<cudaStream_t>(stream));
}

extern "C" int vaemdl_dlogistic_sample(const float* loc, const float* logscale, int C, int ld, const float* u,
                                       long long n_elem, float low, float high, float* x_out, void* stream) {
  if (!loc || !logscale || !u || !x_out || C <= 0 || ld < C || n_elem <= 0) return VAEMDL_EINVAL;
  const DeviceInfo& di = device_info();
  long long blocks = (n_elem + 255) / 256;
  const long long cap = static_cast<long long>(di.sm_count) * 16;
  if (blocks > cap) blocks = cap;
  dl_sample_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(loc, logscale, C, ld, u,
                                                                                               n_elem, low, high, x_out);
  return cuda_rc(cudaGetLastError());
}
