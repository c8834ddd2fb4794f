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
#include <cooperative_groups.h>

#include <cstdlib>
#include <mutex>

#include "modl_math.cuh"
#include "packed.cuh"

namespace vaemdl {

struct DlArgs {
  const float* loc;
  const float* logscale;
  const void* x;
  float* lp_elem;
  double* partial;    // [n_img][K] float64: one partial per (image, warp whose run of tiles touches it)
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
  int small;   // row indices fit 32 bits
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
  // runs numbered CTA-minor: the (one tile longer) first tw_rem runs spread evenly over the SMs
  const long long gw = static_cast<long long>(threadIdx.x >> 5) * gridDim.x + blockIdx.x;
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
        xv = u8_to_unit(static_cast<const uint8_t*>(a.x)[xo + c]);
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
          if (lane == 0) a.partial[partial_slot(n_base, gw, a.rows_per_img, 32, a.tw_base, a.tw_rem, a.K, a.small)] = done;
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
        a.partial[partial_slot(n_base, gw, a.rows_per_img, 32, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
        if (n_base + 1 <= n_last) a.partial[partial_slot(n_base + 1, gw, a.rows_per_img, 32, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
      }
    }
  }
}

// ---- fast path for image tensors (C == 3, an even number of pixels per image): two pixels per lane -----------------------
// The lo / hi halves of Blackwell's packed FP32 pipe (FFMA2 / FMUL2 / FADD2) hold the same channel of two
// neighbouring pixels, the un-split [..,6] conv output of models/model03.py:88-91 arrives as three 128-bit loads per
// lane, and (image, row) indices advance incrementally (no integer division in the loop).
struct DlOut2 {
  f2 lp, dloc, dls;
};
__device__ __forceinline__ f2 lg2_2(f2 a) { return pk(lg2a(lo(a)), lg2a(hi(a))); }
__device__ __forceinline__ f2 abs_2(f2 a) { return pk(fabsf(lo(a)), fabsf(hi(a))); }

template <bool BWD>
__device__ __forceinline__ DlOut2 dl_elem2(f2 x, f2 loc, f2 ls, const DlArgs& a) {
  const bool ll_ = lo(x) <= a.low, lh_ = hi(x) <= a.low;     // utils/discretized_logistic.py:71-73
  const bool rl_ = lo(x) >= a.high, rh_ = hi(x) >= a.high;   // :74-76
  const f2 inv = min_2(ex2_2(ls * (-kLog2e)), 3.0e38f);
  const f2 mid = inv * (x - loc);
  const f2 am = abs_2(mid);
  const f2 A = ex2_2(am * (-kLog2e));
  const f2 h = inv * a.dx;
  f2 q = fma2(h, -1.0f / 720.0f, 1.0f / 120.0f);
  q = fma2(h, q, -1.0f / 24.0f);
  q = fma2(h, q, 1.0f / 6.0f);
  q = fma2(h, q, -0.5f);
  q = fma2(h, q, 1.0f);
  f2 omG = h * q;
  f2 G = sp(1.0f) - omG;
  const bool nl = lo(h) >= kHSmall, nh_ = hi(h) >= kHSmall;
  const bool narrow = __any_sync(kFull, nl || nh_);
  if (narrow) {
    const f2 Ge = ex2_2(h * (-kLog2e));
    G = sel_2(nl, nh_, Ge, G);
    omG = sel_2(nl, nh_, sp(1.0f) - Ge, omG);
  }
  const f2 AG = A * G, ApG = A + G, opAG = AG + 1.0f, opA = A + 1.0f;
  const bool pl = lo(mid) >= 0.0f, ph = hi(mid) >= 0.0f;
  const f2 rest_n = omG * (G + 1.0f);
  const f2 den_n = ApG * opAG;
  const f2 lhs = A * rest_n, thr = den_n * 1e-5f;
  const bool il = lo(lhs) > lo(thr), ih = hi(lhs) > hi(thr);  // prob > 1e-5 (:64)
  const bool el = ll_ || rl_, eh = lh_ || rh_;
  const bool ol = (ll_ == pl), oh = (lh_ == ph);
  // 0/1 masks on the packed pipe instead of selects wherever one side is zero (a select of a packed value costs the
  // compiler four predicated moves; these kernels are issue-bound).  Masked quantities are finite.
  const f2 mi = pk(il ? 1.0f : 0.0f, ih ? 1.0f : 0.0f);  // 1 in the CDF-difference branch
  const f2 nmi = sp(1.0f) - mi;                          // 1 in the low-probability branch
  f2 nu = sel_2(il, ih, rest_n, sp(1.0f));
  f2 de = sel_2(il, ih, den_n, opA * opA);
  f2 cst = nmi * (sp(a.ln_width) - ls);
  f2 use_mid = am;
  if (el || eh) {
    const f2 de_e = sel_2(ol, oh, opAG, ApG);
    const f2 um_e = sel_2(ol, oh, sp(0.0f), am);
    nu = sel_2(el, eh, sp(1.0f), nu);
    de = sel_2(el, eh, de_e, de);
    cst = sel_2(el, eh, sp(0.0f), cst);
    use_mid = sel_2(el, eh, um_e, use_mid);
  }
  DlOut2 o;
  o.lp = fma2(lg2_2(nu) - lg2_2(de), kLn2, cst - use_mid);
  if constexpr (BWD) {
    const f2 omA2 = (sp(1.0f) - A) * opA;
    f2 hc;
    {
      const f2 h2 = h * h;
      hc = fma2(h2, 2.0f / 945.0f, -1.0f / 45.0f);
      hc = fma2(h2, hc, 1.0f / 3.0f);
      hc = fma2(h2, hc, 1.0f);
      if (narrow) {
        const f2 e = h * fma2(G, G, 1.0f) * rcp_2(rest_n);
        hc = sel_2(nl, nh_, e, hc);
      }
    }
    f2 den = de;  // the same denominators serve the derivatives
    // G in the CDF-difference branch, 1 in the low-probability branch (1 - mi * omG is the expression G was formed with
    // unless a lane took the exact exp(-h))
    const f2 gsel = narrow ? sel_2(il, ih, G, sp(1.0f)) : fma2(mi, omG * -1.0f, 1.0f);
    f2 nm = neg_sign_of_2(gsel * omA2, mid);
    f2 nh = mi * ((h * -1.0f) * lhs);
    f2 c0 = mi * hc;
    f2 dir = mi + -1.0f;
    if (el || eh) {
      const f2 t = sel_2(ol, oh, AG, G);
      const f2 nm_e = pk(ll_ ? lo(t) : -lo(t), lh_ ? hi(t) : -hi(t));
      nm = sel_2(el, eh, nm_e, nm);
      nh = sel_2(el, eh, h * t, nh);
      c0 = sel_2(el, eh, sp(0.0f), c0);
      dir = sel_2(el, eh, sp(0.0f), dir);
    }
    const f2 rden = rcp_2(den);
    const f2 Dm = nm * rden;
    o.dloc = (inv * -1.0f) * Dm;
    o.dls = (dir - c0) - fma2(mid, Dm, nh * rden);
  }
  return o;
}

// this lane's pixel pair (rows r0, r0 + 1): location / log-scale / observation of the three channels, pixel A in the lo
// halves, pixel B in the hi halves
template <bool IL>
__device__ __forceinline__ void dl_pair_load(const DlArgs& a, long long r0, long long xb, long long rpi, long long rr,
                                             f2 loc[3], f2 ls[3], f2 xv[3]) {
  if constexpr (IL) {
    const float4* p = reinterpret_cast<const float4*>(a.loc + r0 * 6);
    const float4 u0 = p[0], u1 = p[1], u2 = p[2];  // [locA(3) lsA(3) locB(3) lsB(3)]
    loc[0] = pk(u0.x, u1.z);
    loc[1] = pk(u0.y, u1.w);
    loc[2] = pk(u0.z, u2.x);
    ls[0] = pk(u0.w, u2.y);
    ls[1] = pk(u1.x, u2.z);
    ls[2] = pk(u1.y, u2.w);
  } else {
    const float2* p = reinterpret_cast<const float2*>(a.loc + r0 * 3);
    const float2* q = reinterpret_cast<const float2*>(a.logscale + r0 * 3);
    const float2 u0 = p[0], u1 = p[1], u2 = p[2], v0 = q[0], v1 = q[1], v2 = q[2];  // [A0 A1 | A2 B0 | B1 B2]
    loc[0] = pk(u0.x, u1.y);
    loc[1] = pk(u0.y, u2.x);
    loc[2] = pk(u1.x, u2.y);
    ls[0] = pk(v0.x, v1.y);
    ls[1] = pk(v0.y, v2.x);
    ls[2] = pk(v1.x, v2.y);
  }
  const long long xo = (xb * rpi + rr) * 3;
  if (a.x_u8) {
    const uint8_t* xp = static_cast<const uint8_t*>(a.x) + xo;  // 6 bytes, 2-byte aligned (xo is a multiple of 6)
    const ushort3 w = *reinterpret_cast<const ushort3*>(xp);
    xv[0] = pk(u8_to_unit(w.x & 0xff), u8_to_unit(w.y >> 8));
    xv[1] = pk(u8_to_unit(w.x >> 8), u8_to_unit(w.z & 0xff));
    xv[2] = pk(u8_to_unit(w.y & 0xff), u8_to_unit(w.z >> 8));
  } else {
    const float2* xp = reinterpret_cast<const float2*>(static_cast<const float*>(a.x) + xo);
    const float2 w0 = xp[0], w1 = xp[1], w2 = xp[2];
    xv[0] = pk(w0.x, w1.y);
    xv[1] = pk(w0.y, w2.x);
    xv[2] = pk(w1.x, w2.y);
  }
}

template <bool IL>
__device__ __forceinline__ void dl_pair_store(const DlArgs& a, long long r0, const f2 dl[3], const f2 ds[3]) {
  if constexpr (IL) {
    float4* out = reinterpret_cast<float4*>(a.dloc + r0 * 6);
    out[0] = make_float4(lo(dl[0]), lo(dl[1]), lo(dl[2]), lo(ds[0]));
    out[1] = make_float4(lo(ds[1]), lo(ds[2]), hi(dl[0]), hi(dl[1]));
    out[2] = make_float4(hi(dl[2]), hi(ds[0]), hi(ds[1]), hi(ds[2]));
  } else {
    float2* o1 = reinterpret_cast<float2*>(a.dloc + r0 * 3);
    float2* o2 = reinterpret_cast<float2*>(a.dls + r0 * 3);
    o1[0] = make_float2(lo(dl[0]), lo(dl[1]));
    o1[1] = make_float2(lo(dl[2]), hi(dl[0]));
    o1[2] = make_float2(hi(dl[1]), hi(dl[2]));
    o2[0] = make_float2(lo(ds[0]), lo(ds[1]));
    o2[1] = make_float2(lo(ds[2]), hi(ds[0]));
    o2[2] = make_float2(hi(ds[1]), hi(ds[2]));
  }
}

// IL: loc/logscale are the two halves of one [.., 6] tensor (ld = 6, logscale = loc + 3) and so are dloc/dls.
template <bool BWD, bool IL>
__global__ void __launch_bounds__(256, 3) dl_pair_kernel(const DlArgs a) {
  // runs numbered CTA-minor: the (one tile longer) first tw_rem runs spread evenly over the SMs
  const long long gw = static_cast<long long>(threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  const int lane = threadIdx.x & 31;
  if (!BWD && a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
  if constexpr (!BWD) pdl_trigger();  // the finish kernel's launch may overlap this kernel (it waits for it to complete)
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);
  if (t_begin >= t_end) return;
  const long long rpi = a.rows_per_img;
  // image of the warp's first row: the only integer divisions of the kernel (32-bit when the problem allows it)
  long long n_warp_first, xb;
  if (a.small) {
    n_warp_first = static_cast<unsigned>(t_begin * 64) / static_cast<unsigned>(rpi);
    xb = a.x_batch == 1 ? 0 : static_cast<unsigned>(n_warp_first) % static_cast<unsigned>(a.x_batch);
  } else {
    n_warp_first = (t_begin * 64) / rpi;
    xb = a.x_batch == 1 ? 0 : n_warp_first % a.x_batch;
  }
  // this lane's pixel pair: rows r0, r0 + 1 of image n (rows_per_img is even: a pair never straddles two images)
  long long r0 = t_begin * 64 + 2 * lane;
  long long n = n_warp_first;
  long long rr = r0 - n * rpi;
  while (rr >= rpi) {
    rr -= rpi;
    ++n;
    if (a.x_batch != 1 && ++xb == a.x_batch) xb = 0;
  }
  double acc0 = 0.0, acc1 = 0.0;
  long long n_base = n_warp_first;
  bool any1 = false;  // some row of the current tile belongs to image n_base + 1 (warp-uniform)
  // the loads of tile t + 1 are issued before tile t is evaluated: a tile is only 48 bytes per lane, so without this
  // every warp sat on its own load latency (26 % + 15 % of the stall samples on the first uses of the loaded values)
  bool active = r0 < a.n_rows;
  f2 loc[3], ls[3], xv[3];
  if (active) {
    dl_pair_load<IL>(a, r0, xb, rpi, rr, loc, ls, xv);
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) loc[c] = ls[c] = xv[c] = sp(0.0f);
  }
  for (long long t = t_begin; t < t_end; ++t) {
    // next tile: 64 rows further on
    long long r0n = r0 + 64, rrn = rr + 64, nn = n, xbn = xb;
    while (rrn >= rpi) {
      rrn -= rpi;
      ++nn;
      if (a.x_batch != 1 && ++xbn == a.x_batch) xbn = 0;
    }
    bool nact = false;
    f2 nloc[3], nls[3], nxv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) nloc[c] = nls[c] = nxv[c] = sp(0.0f);
    if (t + 1 < t_end) {
      nact = r0n < a.n_rows;
      if (nact) dl_pair_load<IL>(a, r0n, xbn, rpi, rrn, nloc, nls, nxv);
    }
    DlOut2 o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = dl_elem2<BWD>(xv[c], loc[c], ls[c], a);  // every lane runs it (warp votes inside)
    if constexpr (!BWD) {
      if (a.lp_elem && active) {
        float2* out = reinterpret_cast<float2*>(a.lp_elem + r0 * 3);
        out[0] = make_float2(lo(o[0].lp), lo(o[1].lp));
        out[1] = make_float2(lo(o[2].lp), hi(o[0].lp));
        out[2] = make_float2(hi(o[1].lp), hi(o[2].lp));
      }
      const f2 s2 = o[0].lp + o[1].lp + o[2].lp;
      const float val = active ? lo(s2) + hi(s2) : 0.0f;
      if (a.partial) {
        const long long n_first = __shfl_sync(kFull, n, 0);
        while (n_base < n_first) {
          const double done = dl_warp_sum(acc0);
          if (lane == 0) a.partial[partial_slot(n_base, gw, rpi, 64, a.tw_base, a.tw_rem, a.K, a.small)] = done;
          acc0 = acc1;
          acc1 = 0.0;
          ++n_base;
        }
        if (n == n_base)
          acc0 += static_cast<double>(val);
        else
          acc1 += static_cast<double>(val);
        any1 = __any_sync(kFull, active && n != n_base);
      } else if (a.ll_atomic && active) {
        atomicAdd(a.ll_atomic + n, static_cast<double>(val));
      }
    } else {
      // The upstream gradient is only needed here, after the tile's derivatives have been formed: a gradient kernel
      // launched programmatically behind the finish kernel loads and evaluates its first tile while that kernel still runs.
      if (t == t_begin) pdl_wait();
    }
    if (BWD && active) {
      const float g = a.g_image ? a.g_image[n] : 0.0f;
      f2 ge[3] = {sp(g), sp(g), sp(g)};
      if (a.g_elem) {
        const float2* gp = reinterpret_cast<const float2*>(a.g_elem + r0 * 3);
        const float2 w0 = gp[0], w1 = gp[1], w2 = gp[2];
        ge[0] = ge[0] + pk(w0.x, w1.y);
        ge[1] = ge[1] + pk(w0.y, w2.x);
        ge[2] = ge[2] + pk(w1.x, w2.y);
      }
      f2 dl[3], ds[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dl[c] = ge[c] * o[c].dloc;
        ds[c] = ge[c] * o[c].dls;
      }
      dl_pair_store<IL>(a, r0, dl, ds);
    }
    r0 = r0n;
    rr = rrn;
    n = nn;
    xb = xbn;
    active = nact;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      loc[c] = nloc[c];
      ls[c] = nls[c];
      xv[c] = nxv[c];
    }
  }
  if constexpr (!BWD) {
    if (a.partial) {
      const double d0 = dl_warp_sum(acc0);
      if (lane == 0) a.partial[partial_slot(n_base, gw, rpi, 64, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
      if (any1) {
        const double d1 = dl_warp_sum(acc1);
        if (lane == 0) a.partial[partial_slot(n_base + 1, gw, rpi, 64, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
      }
    }
  }
}

// ---- one step in one launch (models 03 / 04 / 06: BASELINE configs[1]) ---------------------------------------------------
// Forward, IWAE finish and gradient of the plain discretized logistic in ONE cooperative kernel.  The shapes this serves
// are small (S x B x 32 x 32 x 3 with a few hundred images: a handful of 64-pixel tiles per resident warp), so every warp
// parks the unscaled derivatives of its tiles in shared memory across the two grid barriers: the parameters are read
// ONCE, the logistic terms are evaluated ONCE, and after the barriers the gradient is one multiply per element.
// Same formulas and summation order as dl_pair_kernel<false> + finish_kernel + dl_pair_kernel<true>; the forward value comes
// out of the gradient instantiation of dl_elem2 here, so the two routes agree to float32 round-off (not bit for bit).
constexpr int kStepWarps = 24;  // warps per CTA of the one-launch step: one 768-thread CTA per SM (a third of the CTAs of
                                // 256-thread blocks at the two grid barriers, the same number of warps)
constexpr int kStepKeep = 13;  // floats a lane parks per tile: 6 derivative pairs + the image index of its pixel pair
struct DlStepArgs {
  DlArgs a;
  StepFinish f;
  int T;  // tiles a warp can park (its run is at most T tiles long)
};

template <bool IL>
__global__ void __launch_bounds__(kStepWarps * 32, 1) dl_step_kernel(const DlStepArgs sa) {
  extern __shared__ float dl_keep[];  // [warps][T][kStepKeep][32]
  const DlArgs& a = sa.a;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = static_cast<long long>(warp) * gridDim.x + blockIdx.x;
  const long long total_warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);   // t_end - t_begin <= T
  const long long rpi = a.rows_per_img;
  float* keep = dl_keep + static_cast<size_t>(warp) * sa.T * (kStepKeep * 32) + lane;
  if (t_begin < t_end) {
    long long n_warp_first, xb;
    if (a.small) {
      n_warp_first = static_cast<unsigned>(t_begin * 64) / static_cast<unsigned>(rpi);
      xb = a.x_batch == 1 ? 0 : static_cast<unsigned>(n_warp_first) % static_cast<unsigned>(a.x_batch);
    } else {
      n_warp_first = (t_begin * 64) / rpi;
      xb = a.x_batch == 1 ? 0 : n_warp_first % a.x_batch;
    }
    long long r0 = t_begin * 64 + 2 * lane;
    long long n = n_warp_first;
    long long rr = r0 - n * rpi;
    while (rr >= rpi) {
      rr -= rpi;
      ++n;
      if (a.x_batch != 1 && ++xb == a.x_batch) xb = 0;
    }
    double acc0 = 0.0, acc1 = 0.0;
    long long n_base = n_warp_first;
    bool any1 = false;
    for (long long t = t_begin; t < t_end; ++t) {
      const bool active = r0 < a.n_rows;
      f2 loc[3], ls[3], xv[3];
      if (active) {
        dl_pair_load<IL>(a, r0, xb, rpi, rr, loc, ls, xv);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) loc[c] = ls[c] = xv[c] = sp(0.0f);
      }
      float* kp = keep + (t - t_begin) * (kStepKeep * 32);
      f2 s2 = sp(0.0f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const DlOut2 o = dl_elem2<true>(xv[c], loc[c], ls[c], a);
        kp[(4 * c + 0) * 32] = lo(o.dloc);
        kp[(4 * c + 1) * 32] = hi(o.dloc);
        kp[(4 * c + 2) * 32] = lo(o.dls);
        kp[(4 * c + 3) * 32] = hi(o.dls);
        s2 = c == 0 ? o.lp : s2 + o.lp;
      }
      kp[12 * 32] = __int_as_float(active ? static_cast<int>(n) : -1);  // (S * B images: far below 2^31)
      const float val = active ? lo(s2) + hi(s2) : 0.0f;
      const long long n_first = __shfl_sync(kFull, n, 0);
      while (n_base < n_first) {
        const double done = dl_warp_sum(acc0);
        if (lane == 0) a.partial[partial_slot(n_base, gw, rpi, 64, a.tw_base, a.tw_rem, a.K, a.small)] = done;
        acc0 = acc1;
        acc1 = 0.0;
        ++n_base;
      }
      if (n == n_base)
        acc0 += static_cast<double>(val);
      else
        acc1 += static_cast<double>(val);
      any1 = __any_sync(kFull, active && n != n_base);
      r0 += 64;
      rr += 64;
      while (rr >= rpi) {
        rr -= rpi;
        ++n;
        if (a.x_batch != 1 && ++xb == a.x_batch) xb = 0;
      }
    }
    const double d0 = dl_warp_sum(acc0);
    if (lane == 0) a.partial[partial_slot(n_base, gw, rpi, 64, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
    if (any1) {
      const double d1 = dl_warp_sum(acc1);
      if (lane == 0) a.partial[partial_slot(n_base + 1, gw, rpi, 64, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
    }
  }
  __threadfence();
  grid.sync();
  step_finish(sa.f, gw, total_warps, lane);
  __threadfence();
  grid.sync();
  if (sa.f.elbo && gw == total_warps - 1) {  // batch mean, fixed order
    double t = 0.0;
    for (long long b = lane; b < sa.f.B; b += 32) t += sa.f.lme64[b];
    t = dl_warp_sum(t);
    if (lane == 0) {
      const float e = static_cast<float>(t / static_cast<double>(sa.f.b_norm));  // models/loss.py:37
      sa.f.elbo[0] = e;
      peer_publish(sa.f.peer, e);
    }
  }
  for (long long t = t_begin; t < t_end; ++t) {
    const float* kp = keep + (t - t_begin) * (kStepKeep * 32);
    const int n = __float_as_int(kp[12 * 32]);
    if (n >= 0) {
      const f2 ge = sp(sa.f.g_ll[n]);
      f2 dl[3], ds[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dl[c] = ge * pk(kp[(4 * c + 0) * 32], kp[(4 * c + 1) * 32]);
        ds[c] = ge * pk(kp[(4 * c + 2) * 32], kp[(4 * c + 3) * 32]);
      }
      dl_pair_store<IL>(a, t * 64 + 2 * lane, dl, ds);
    }
  }
}

__global__ void dl_cast_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]);
}

// sampler: clip(loc + exp(ls) * (log u - log(1-u)), low, high)   (utils/discretized_logistic.py:80-85)
// float32 with the accurate logf / log1pf / expf (the reference's own precision): whenever the result is not clipped,
// |exp(ls) * eps| <= |high - low| + |loc|, so one-ulp errors in eps and exp(ls) stay below ~3e-7 absolute -- float64
// transcendentals made this kernel FP64-pipe-bound at a quarter of the HBM roofline (models/model06.py:166 draws x on
// every forward pass).  Four elements per thread and iteration keep enough loads in flight.
__global__ void dl_sample_kernel(const float* __restrict__ loc, const float* __restrict__ logscale, int C, int ld,
                                 const float* __restrict__ u, long long n_elem, float low, float high,
                                 float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long e0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (long long base = e0; base < n_elem; base += 4 * stride) {
    float uu[4], lc[4], ls[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long e = base + k * stride;
      uu[k] = 0.5f;
      lc[k] = ls[k] = 0.0f;
      if (e < n_elem) {
        const long long row = e / C;
        const long long po = row * ld + (e - row * C);
        uu[k] = u[e];
        lc[k] = loc[po];
        ls[k] = logscale[po];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long e = base + k * stride;
      const float eps = logf(uu[k]) - log1pf(-uu[k]);
      const float v = fminf(fmaxf(fmaf(expf(ls[k]), eps, lc[k]), low), high);
      if (e < n_elem) out[e] = v;
    }
  }
}

// image tensors (C = 3): one thread per pixel, no index division; loc / logscale rows are ld floats apart
__global__ void dl_sample3_kernel(const float* __restrict__ loc, const float* __restrict__ logscale, int ld,
                                  const float* __restrict__ u, long long n_rows, float low, float high,
                                  float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
    float uu[3], lc[3], ls[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uu[c] = u[r * 3 + c];
      lc[c] = loc[r * ld + c];
      ls[c] = logscale[r * ld + c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float eps = logf(uu[c]) - log1pf(-uu[c]);
      out[r * 3 + c] = fminf(fmaxf(fmaf(expf(ls[c]), eps, lc[c]), low), high);
    }
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
  a.small = a.n_rows < (1ll << 31) - 64;
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

// which kernel serves these arguments: 0 = generic (one row of CPT channels per lane, 32-row tiles),
// 1 = pixel pairs on the un-split [..,6] layout, 2 = pixel pairs on separate dense [..,3] tensors (64-row tiles)
static int dl_kind(const DlArgs& a, int cpt, bool bwd) {
  auto al = [](const void* p, uintptr_t m) { return (reinterpret_cast<uintptr_t>(p) & (m - 1)) == 0; };
  if (cpt != 3 || a.C != 3 || (a.rows_per_img & 1)) return 0;
  if (!a.x_u8 && !al(a.x, 8)) return 0;
  if (a.lp_elem && !al(a.lp_elem, 8)) return 0;
  if (a.g_elem && !al(a.g_elem, 8)) return 0;
  if (a.ld == 6 && a.logscale == a.loc + 3 && al(a.loc, 16)) {
    if (!bwd || (a.ld_out == 6 && a.dls == a.dloc + 3 && al(a.dloc, 16))) return 1;
  }
  if (a.ld == 3 && al(a.loc, 8) && al(a.logscale, 8)) {
    if (!bwd || (a.ld_out == 3 && al(a.dloc, 8) && al(a.dls, 8))) return 2;
  }
  return 0;
}
static int dl_tile_rows(int kind) { return kind ? 64 : 32; }

template <bool BWD>
static int dl_launch(DlArgs a, int cpt, int kind, cudaStream_t st, PartialGeom* geom = nullptr) {
  const DeviceInfo& di = device_info();
  const int TR = dl_tile_rows(kind);
  const long long n_tiles = (a.n_rows + TR - 1) / TR;
  long long blocks = (n_tiles + 7) / 8;
  // a warp's start-up and flush (index divisions, warp reductions) cost about as much as one tile: give every warp a
  // few tiles rather than spreading a small problem over as many warps as possible
  static const int per_sm = [] {
    const char* e = getenv("VAEMDL_DL_BLOCKS_PER_SM");
    return e ? atoi(e) : 0;
  }();
  // pixel-pair kernels: 80 registers -> 3 resident CTAs per SM, and a fully resident grid measured best (162 vs 197 us
  // forward at 16 x 256 x 64 x 64 with 8 CTAs per SM, i.e. 2.7 waves)
  long long cap = static_cast<long long>(di.sm_count) * (per_sm > 0 ? per_sm : (kind ? 3 : 8));
  if (cap * 8 > kMaxGridWarps) cap = kMaxGridWarps / 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const long long total_warps = blocks * 8;
  a.tw_base = n_tiles / total_warps;
  a.tw_rem = n_tiles % total_warps;
  a.K = partial_K(a.rows_per_img, TR, a.tw_base);
  if (a.partial && !partials_fit(a.n_rows / a.rows_per_img, a.K)) return VAEMDL_EWORKSPACE;  // (before anything is enqueued)
  if (geom) *geom = PartialGeom{a.partial, a.tw_base, a.tw_rem, a.K, TR, a.rows_per_img};
  const unsigned grid = static_cast<unsigned>(blocks);
  if (BWD && kind) {  // programmatic launch: the kernel's prologue and first tile overlap the finish kernel (see dl_pair_kernel)
    void (*kern)(DlArgs) = kind == 1 ? dl_pair_kernel<BWD, true> : dl_pair_kernel<BWD, false>;
    return cuda_rc(launch_pdl(kern, grid, 256u, 0, st, a));
  }
  if (kind == 1)
    dl_pair_kernel<BWD, true><<<grid, 256, 0, st>>>(a);
  else if (kind == 2)
    dl_pair_kernel<BWD, false><<<grid, 256, 0, st>>>(a);
  else if (cpt == 3)
    dl_kernel<3, BWD><<<grid, 256, 0, st>>>(a);
  else
    dl_kernel<1, BWD><<<grid, 256, 0, st>>>(a);
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
  const int kind = dl_kind(a, cpt, false);
  const bool use_partials = want_ll && a.rows_per_img >= dl_tile_rows(kind);
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
  rc = dl_launch<false>(a, cpt, kind, st, &geom);
  if (rc) return rc;
  if (use_partials)
    return finish_partials(geom, n_img, ll_image, ll_image_f64, iw, reinterpret_cast<double*>(ws + tail_off), counter, st);
  if (!want_ll) return VAEMDL_OK;
  if (ll_image) {
    dl_cast_kernel<<<static_cast<unsigned>((n_img + 255) / 256), 256, 0, st>>>(a.ll_atomic, ll_image, n_img);
    rc = cuda_rc(cudaGetLastError());
  }
  if (rc || !iwae) return rc;
  rc = vaemdl_iwae_tail(nullptr, a.ll_atomic, iw.extra, iw.S, iw.B, iw.B_total, iw.log_w, iw.lme_b, iw.elbo, iw.g_ll, st);
  if (rc || !iw.elbo) return rc;
  return peer_push(iw.peer, iw.elbo, st);
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
  iw.peer = take_peer();
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
  return dl_launch<true>(a, cpt, dl_kind(a, cpt, true), static_cast<cudaStream_t>(stream));
}

namespace vaemdl {
constexpr int kStepMaxT = 12;
// smallest number of parked tiles per warp T for which a grid of resident CTAs covers n_tiles; 0 = not eligible
static int dl_step_plan(const DlArgs& a, int kind, int S, long long n_tiles, long long* blocks_out) {
  // VAEMDL_FUSED=1: the one-launch cooperative step.  Default: three launches -- since the gradient kernel evaluates its
  // first tile while the finish kernel runs they are ahead (BASELINE configs[1]: 25.1 vs 26.7 us, 23.0 vs 24.4 us as graphs).
  const char* e = getenv("VAEMDL_FUSED");
  if (!(e && e[0] == '1')) return 0;
  if (S > 32 || (kind != 1 && kind != 2) || a.rows_per_img < 64) return 0;
  static int coop = -1, occ[2][kStepMaxT + 1];
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (coop < 0) {
      int dev = 0, v = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev);
      const int max_smem = device_info().max_smem_optin;
      cudaFuncSetAttribute(dl_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
      cudaFuncSetAttribute(dl_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
      for (int T = 1; T <= kStepMaxT; ++T) {
        const size_t smem = static_cast<size_t>(T) * kStepWarps * kStepKeep * 32 * 4;
        occ[0][T] = occ[1][T] = 0;
        if (smem > static_cast<size_t>(max_smem)) continue;  // more parked tiles than shared memory holds
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0][T], dl_step_kernel<true>, kStepWarps * 32, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1][T], dl_step_kernel<false>, kStepWarps * 32, smem);
      }
      coop = v;
    }
  }
  if (!coop) return 0;
  for (int T = 1; T <= kStepMaxT; ++T) {
    long long blocks = static_cast<long long>(device_info().sm_count) * occ[kind == 1 ? 0 : 1][T];  // every CTA resident
    if (blocks * kStepWarps > kMaxGridWarps) blocks = kMaxGridWarps / kStepWarps;
    if (blocks < 1) continue;
    const long long need = (n_tiles + kStepWarps - 1) / kStepWarps;
    if (blocks > need) blocks = need;
    const long long total_warps = blocks * kStepWarps;
    if ((n_tiles + total_warps - 1) / total_warps <= T) {
      *blocks_out = blocks;
      return T;
    }
  }
  return 0;
}
}  // namespace vaemdl

/* One IWAE step of the plain discretized logistic in one call (models/model03.py:139-148, models/model06.py): one
 * cooperative launch for the small image shapes, forward + finish + gradient (3 launches) otherwise. */
extern "C" int vaemdl_dlogistic_iwae_step(const float* loc, const float* logscale, int C, int ld, const void* x,
                                          int x_dtype, int S, long long B, long long B_total, int x_batch, long long D,
                                          float low, float high, float levels, const float* extra, float* ll_image,
                                          double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                                          float* dloc, float* dlogscale, int ld_out, void* workspace,
                                          size_t workspace_bytes, void* stream, int* launches) {
  if (S <= 0 || B <= 0 || B_total < 0 || !lme_b || !g_ll || !dloc || !dlogscale || ld_out < C) return VAEMDL_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n_img = static_cast<long long>(S) * B;
  DlArgs a{};
  int cpt = 1;
  int rc = dl_fill(a, loc, logscale, C, ld, x, x_dtype, n_img, x_batch, D, low, high, levels, cpt);
  if (rc) return rc;
  a.dloc = dloc;
  a.dls = dlogscale;
  a.ld_out = ld_out;
  const int kind = dl_kind(a, cpt, true);
  const long long n_tiles = (a.n_rows + 63) / 64;
  long long blocks = 0;
  const int T = dl_step_plan(a, kind, S, n_tiles, &blocks);
  if (T == 0) {
    if (launches) *launches = 3;
    rc = vaemdl_dlogistic_iwae_fwd(loc, logscale, C, ld, x, x_dtype, S, B, B_total, x_batch, D, low, high, levels, extra,
                                   ll_image, ll_image_f64, log_w, lme_b, elbo, g_ll, workspace, workspace_bytes, stream);
    if (rc) return rc;
    return vaemdl_dlogistic_bwd(loc, logscale, C, ld, x, x_dtype, n_img, x_batch, D, low, high, levels, g_ll, nullptr, dloc,
                                dlogscale, ld_out, stream);
  }
  if (!workspace || workspace_bytes < vaemdl_dlogistic_workspace_bytes(n_img, D)) return VAEMDL_EWORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 7u) return VAEMDL_EALIGN;
  char* ws = static_cast<char*>(workspace);
  a.partial = reinterpret_cast<double*>(ws);
  const long long total_warps = blocks * kStepWarps;
  a.tw_base = n_tiles / total_warps;
  a.tw_rem = n_tiles % total_warps;
  a.K = partial_K(a.rows_per_img, 64, a.tw_base);
  if (!partials_fit(n_img, a.K)) return VAEMDL_EWORKSPACE;
  DlStepArgs sa{};
  sa.a = a;
  sa.T = T;
  sa.f.geom = PartialGeom{a.partial, a.tw_base, a.tw_rem, a.K, 64, a.rows_per_img};
  sa.f.extra = extra;
  sa.f.ll = ll_image;
  sa.f.ll64 = ll_image_f64;
  sa.f.log_w = log_w;
  sa.f.lme_b = lme_b;
  sa.f.elbo = elbo;
  sa.f.g_ll = g_ll;
  sa.f.lme64 = reinterpret_cast<double*>(ws + partial_elems(n_img) * sizeof(double));
  sa.f.B = B;
  sa.f.S = S;
  sa.f.b_norm = static_cast<float>(B_total > 0 ? B_total : B);
  sa.f.small = n_img * a.rows_per_img < (1ll << 31);
  sa.f.peer = elbo ? take_peer() : PeerOut{};
  if (launches) *launches = 1;
  void* args[] = {&sa};
  void* kern = kind == 1 ? reinterpret_cast<void*>(dl_step_kernel<true>) : reinterpret_cast<void*>(dl_step_kernel<false>);
  rc = cuda_rc(cudaLaunchCooperativeKernel(kern, dim3(static_cast<unsigned>(blocks)), dim3(kStepWarps * 32), args,
                                           static_cast<size_t>(T) * kStepWarps * kStepKeep * 32 * 4, st));
  if (rc == static_cast<int>(cudaErrorCooperativeLaunchTooLarge) || rc == static_cast<int>(cudaErrorLaunchOutOfResources)) {
    // the grid cannot be co-resident right now: nothing was enqueued, run the three ordinary launches instead
    cudaGetLastError();
    give_peer(sa.f.peer);
    if (launches) *launches = 3;
    rc = vaemdl_dlogistic_iwae_fwd(loc, logscale, C, ld, x, x_dtype, S, B, B_total, x_batch, D, low, high, levels, extra,
                                   ll_image, ll_image_f64, log_w, lme_b, elbo, g_ll, workspace, workspace_bytes, stream);
    if (rc) return rc;
    return vaemdl_dlogistic_bwd(loc, logscale, C, ld, x, x_dtype, n_img, x_batch, D, low, high, levels, g_ll, nullptr, dloc,
                                dlogscale, ld_out, stream);
  }
  return rc;
}

extern "C" int vaemdl_dlogistic_sample(const float* loc, const float* logscale, int C, int ld, const float* u,
                                       long long n_elem, float low, float high, float* x_out, void* stream) {
  if (!loc || !logscale || !u || !x_out || C <= 0 || ld < C || n_elem <= 0) return VAEMDL_EINVAL;
  const DeviceInfo& di = device_info();
  if (C == 3 && n_elem % 3 == 0) {
    const long long n_rows = n_elem / 3;
    long long rb = (n_rows + 255) / 256;
    const long long rcap = static_cast<long long>(di.sm_count) * 32;
    if (rb > rcap) rb = rcap;
    dl_sample3_kernel<<<static_cast<unsigned>(rb), 256, 0, static_cast<cudaStream_t>(stream)>>>(loc, logscale, ld, u, n_rows,
                                                                                              low, high, x_out);
    return cuda_rc(cudaGetLastError());
  }
  long long blocks = (n_elem + 4 * 256 - 1) / (4 * 256);
  const long long cap = static_cast<long long>(di.sm_count) * 16;
  if (blocks < 1) blocks = 1;
  if (blocks > cap) blocks = cap;
  dl_sample_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(loc, logscale, C, ld, u,
                                                                                               n_elem, low, high, x_out);
  return cuda_rc(cudaGetLastError());
}
