// modl_math.cuh -- per-sub-pixel discretized-logistic arithmetic shared by the MoDL and plain-DL kernels.
//
// Reference formulas (utils/mdl.py:165-207 == utils/mdl_openai.py:111-150 == utils/discretized_logistic.py:35-78):
//   inv = exp(-ls); mid = inv*(x-loc); p = mid + h; q = mid - h; h = inv*dx
//   left  edge : log sigmoid(p)            right edge : log sigmoid(-q)
//   normal     : log(sigmoid(p)-sigmoid(q))   if that difference > 1e-5
//   low-prob   : mid - ls - 2 softplus(mid) + log(width)          otherwise
//
// Kernel formulation.  With A = exp(-|mid|) <= 1 and G = exp(-h) <= 1 every branch is a ratio num/den of
// cancellation-free positive terms (the identity sigmoid(m+h)-sigmoid(m-h) = sinh h / (cosh m + cosh h)):
//   normal   : A (1-G^2) / ((A+G)(1+AG))
//   low-prob : A inv width / (1+A)^2
//   left     : mid>=0 ? 1/(1+AG) : A/(A+G)          right : mid>=0 ? A/(A+G) : 1/(1+AG)
// so one pixel-mixture costs 2-3 MUFU per sub-pixel (exp(-ls), exp(-|mid|), and exp(-h) only when h is not
// small) and the three sub-pixels share ONE reciprocal; no logarithm is taken per sub-pixel at all -- the mixture
// sum is accumulated in the linear domain and only the per-pixel result goes through lg2 (see modl_kernels.cu).
// The derivatives needed by the backward kernel are ratios with the same denominators.
#pragma once
#include "common.cuh"

namespace vaemdl {

constexpr float kHSmall = 0.15f;  // below this, exp(-h) and h*coth(h) come from short polynomials (FMA pipe, no MUFU)

struct SubF {  // forward: f = num / den
  float num, den;
};

struct SubB {  // backward: adds the derivative numerators (all share den)
  float num, den;
  float nm;   // d log f / d mid  = nm / den
  float nh;   // h * d log f / d h = c0 + nh / den
  float c0;
  float dir;  // direct d log f / d ls term (low-prob branch: -1)
  float inv, mid;
};

// exp(-h) and 1-exp(-h) for h >= 0.  Polynomial (relative error < 1.2e-7 for h < kHSmall) unless the warp has a lane
// with a narrow scale; the MUFU path is taken warp-uniformly so wide-scale data never touches the SFU for this term.
__device__ __forceinline__ void exp_neg_h(float h, float& G, float& omG) {
  // 1-exp(-h) = h*(1 - h/2 + h^2/6 - h^3/24 + h^4/120 - h^5/720)
  float q = fmaf(h, -1.0f / 720.0f, 1.0f / 120.0f);
  q = fmaf(h, q, -1.0f / 24.0f);
  q = fmaf(h, q, 1.0f / 6.0f);
  q = fmaf(h, q, -0.5f);
  q = fmaf(h, q, 1.0f);
  omG = h * q;
  G = 1.0f - omG;
  if (__any_sync(kFull, h >= kHSmall)) {
    const float Ge = ex2a(-h * kLog2e);
    if (h >= kHSmall) {
      G = Ge;
      omG = 1.0f - Ge;
    }
  }
}

// h*coth(h) = h (1+G^2)/(1-G^2)
__device__ __forceinline__ float h_coth_h(float h, float G, float omG) {
  const float h2 = h * h;
  // 1 + h^2/3 - h^4/45 + 2 h^6/945
  float r = fmaf(h2, 2.0f / 945.0f, -1.0f / 45.0f);
  r = fmaf(h2, r, 1.0f / 3.0f);
  r = fmaf(h2, r, 1.0f);
  if (__any_sync(kFull, h >= kHSmall)) {
    const float e = h * fmaf(G, G, 1.0f) * rcpa(omG * (1.0f + G));
    if (h >= kHSmall) r = e;
  }
  return r;
}

// One sub-pixel.  ls is the (already clamped) log-scale, dx the half bin width, width = 2*dx.
template <bool BWD>
__device__ __forceinline__ void subpix(float x, bool left, bool right, float loc, float ls, float dx, float width,
                                       typename std::conditional<BWD, SubB, SubF>::type& o) {
  const float inv = ex2a(-ls * kLog2e);
  const float mid = inv * (x - loc);
  const float A = ex2a(-fabsf(mid) * kLog2e);
  const float h = inv * dx;
  float G, omG;
  exp_neg_h(h, G, omG);
  const float AG = A * G;
  const float ApG = A + G;
  const float opAG = 1.0f + AG;
  const float opA = 1.0f + A;
  const bool pos = mid >= 0.0f;
  // normal branch
  const float num_n = A * omG * (1.0f + G);
  const float den_n = ApG * opAG;
  const bool is_norm = num_n > 1e-5f * den_n;  // sigmoid(p)-sigmoid(q) > 1e-5   (utils/mdl.py:193)
  // low-probability branch
  const float num_l = A * inv * width;
  const float den_l = opA * opA;
  float num = is_norm ? num_n : num_l;
  float den = is_norm ? den_n : den_l;
  const bool edge = left || right;
  // (left & mid>=0) or (right & mid<0): 1/(1+AG);  otherwise A/(A+G)
  const bool one_over = (left == pos);
  if (edge) {
    num = one_over ? 1.0f : A;
    den = one_over ? opAG : ApG;
  }
  o.num = num;
  o.den = den;
  if constexpr (BWD) {
    const float omA2 = (1.0f - A) * opA;
    const float sgn = pos ? 1.0f : -1.0f;
    float nm = is_norm ? -sgn * G * omA2 : -sgn * omA2;
    float nh = is_norm ? -h * num_n : 0.0f;
    const float hc = h_coth_h(h, G, omG);  // evaluated by ALL lanes (warp-uniform vote inside)
    float c0 = is_norm ? hc : 0.0f;
    float dir = is_norm ? 0.0f : -1.0f;
    if (edge) {
      // derivative of log sigmoid(p) is sigmoid(-p); of log sigmoid(-q) is -sigmoid(q) (w.r.t. q = mid - h)
      const float t = one_over ? AG : G;  // numerator of sigmoid(-p) resp. sigmoid(q)
      nm = left ? t : -t;
      nh = h * t;
      c0 = 0.0f;
      dir = 0.0f;
    }
    o.nm = nm;
    o.nh = nh;
    o.c0 = c0;
    o.dir = dir;
    o.inv = inv;
    o.mid = mid;
  }
}

// tanh of the three autoregressive coefficients with 3 ex2 + ONE shared reciprocal.
__device__ __forceinline__ void tanh3(float k0, float k1, float k2, float& t0, float& t1, float& t2) {
  const float c = 2.0f * kLog2e;
  const float E0 = ex2a(fminf(fmaxf(k0, -10.0f), 10.0f) * c);
  const float E1 = ex2a(fminf(fmaxf(k1, -10.0f), 10.0f) * c);
  const float E2 = ex2a(fminf(fmaxf(k2, -10.0f), 10.0f) * c);
  const float d0 = E0 + 1.0f, d1 = E1 + 1.0f, d2 = E2 + 1.0f;
  const float d01 = d0 * d1;
  const float R = rcpa(d01 * d2);
  t0 = (E0 - 1.0f) * (d1 * d2) * R;
  t1 = (E1 - 1.0f) * (d0 * d2) * R;
  t2 = (E2 - 1.0f) * d01 * R;
}

// ---------------------------------------------------------------------------------------------------------------
// Log-domain evaluation of one sub-pixel, used only by the rare per-pixel fallback when the linear-domain mixture
// sum leaves the float32 range.  Accurate libm-style functions, not MUFU approximations.
static __device__ __noinline__ float subpix_logf(float x, bool left, bool right, float loc, float ls, float dx, float width) {
  const float inv = expf(-ls);
  const float mid = inv * (x - loc);
  const float am = fabsf(mid);
  const float A = expf(-am);
  const float h = inv * dx;
  const float G = expf(-h);
  const float omG = -expm1f(-h);
  const bool pos = mid >= 0.0f;
  const float AG = A * G, ApG = A + G, opAG = 1.0f + AG;
  if (left || right) {
    const bool one_over = (left == pos);
    return one_over ? -log1pf(AG) : (-am - logf(ApG));
  }
  const float num_n = A * omG * (1.0f + G);
  const float den_n = ApG * opAG;
  if (num_n > 1e-5f * den_n) return -am + logf(omG * (1.0f + G)) - logf(den_n);
  return -am - ls + logf(width) - 2.0f * log1pf(A);
}

}  // namespace vaemdl
