#pragma once
// modl_kernels.cuh -- mixture-of-discretized-logistics log-likelihood (forward) and parameter gradient (backward):
// kernel templates, launchers and the host-side implementation shared by modl_kernels.cu (means chained on the observed
// x, utils/mdl.py / utils/mdl_openai*.py) and modl_plain.cu (means chained on the means, utils/mdl_plain.py).
//
// Replaces utils/mdl.py:56-207, utils/mdl_openai.py:83-157, utils/mdl_openai_iwae.py:33-67 and the autodiff of
// models/model05.py:141-145 (see include/vaemdl.h).
//
// Data movement.  The parameter tensor is a flat stream of rows (one 40*M-byte row per pixel-sample).  Every warp
// owns a run of consecutive tiles and runs its own pipeline: a 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) brings a
// tile of PPT rows into the warp's shared-memory slot and signals an mbarrier; each lane reads its own row (or its
// MC-mixture chunk of the row) with 64-bit shared loads inside a rolled loop over component pairs.  The backward
// kernel overwrites the row in place with the unscaled gradients, rescales them once the pixel's mixture sum is
// known and hands the tile back with a bulk store.  Per-image sums stay in registers (float64) along the run and leave
// the warp once per image (finish.cu adds them up).  There is no CTA-wide synchronisation.
//
// Work split.  n_mix 10 / 20 / 30: M = MC * LPP, LPP lanes share a pixel, each owning MC components, two components
// per packed register (modl_tile_kernel: 10x1, 10x2, 10x3).  n_mix 1..9: one lane owns two pixels, the same component
// of both per packed register (modl_pp_kernel).  Any other M runs on a plain one-thread-per-pixel kernel (correct,
// not tuned).  Template parameter AR selects what the green / blue means are chained on (pair_eval).
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "modl_math.cuh"
#include "packed.cuh"

namespace vaemdl {

constexpr float kDx = 1.0f / 255.0f;     // half bin width in [-1,1] units   (utils/mdl.py:47-50)
constexpr float kWidth = 2.0f / 255.0f;  // bin width                         (utils/mdl.py:47)
constexpr float kTinySum = 1e-30f;       // below this the linear-domain mixture sum is re-done in the log domain

struct ModlArgs {
  const float* params;
  const void* x;
  float* lp_pixel;       // nullable
  double* partial;       // [n_img][K] float64 partial sums, one per (image, warp whose tile run touches it) (nullable)
  double* ll_atomic;     // [n_img] pre-zeroed float64 accumulators, used instead of `partial` when tiles would span >2 images
  const float* g_image;  // nullable
  const float* g_pixel;  // nullable
  float* dparams;
  unsigned* zero_me;  // nullable: a word the forward kernel clears for the fused finish kernel that follows it
  long long n_px;  // n_img * H * W
  long long num_tiles;
  long long tw_base, tw_rem;  // warp w owns tiles [w*tw_base + min(w, tw_rem), +tw_base + (w < tw_rem)): consecutive tiles
  int K;                      // partial slots per image: max number of warp runs one image can intersect
  int small;                  // n_px fits 32 bits
  int reverse;                // backward only: walk the run from its last tile to its first (the tiles the forward kernel
                              // read last are the ones still in L2)
  int keep_tiles;             // forward: the last keep_tiles tiles of a run are loaded with an L2 evict_last hint
  int bwd_hint;               // backward: L2 evict_first hint on the parameter loads (bit 0) / gradient stores (bit 1)
  int plain;                  // 1: channel means chained on the component's own means (utils/mdl_plain.py:160-162),
                              // 0: on the observed x (utils/mdl.py:139-145); the fast kernels take this as template AR
  int HW;
  int x_batch;
  int x_u8;
  int x_unit;       // apply x*2-1
  int edge_openai;  // < -0.999 / > 0.999 instead of <= -1 / >= 1
  int M;
  int bf16;    // parameters (and the gradient) are bfloat16 in global memory; all arithmetic stays float32
  int spread;  // 1: run r belongs to warp (r / #CTAs) of CTA (r % #CTAs), 0: to warp (r % warps) of CTA (r / warps)
  // run-time tile geometry (modl_rt_kernel: any n_mix without its own instantiation)
  int rt_MC, rt_LPP, rt_PPT, rt_rot, rt_warp_f;
};

// Which run of consecutive tiles a warp owns.  The first tw_rem runs are one tile longer than the rest; numbering the
// runs CTA-minor spreads those evenly over the SMs (CTA-major puts all of them on the first tw_rem / warps SMs, which
// then finish a whole tile after the others: 4.3 tiles per warp = 14 % of the kernel at BASELINE configs[0]).
__device__ __forceinline__ long long run_index(const ModlArgs& a, int warp, int nwarps) {
  return a.spread ? static_cast<long long>(warp) * gridDim.x + blockIdx.x : static_cast<long long>(blockIdx.x) * nwarps + warp;
}

// ---- bfloat16 parameters (the decoder's conv output arrives in bf16) ------------------------------------------
// The tile travels as bf16 (half the DRAM bytes) and is widened to float32 IN PLACE in the warp's slot before the compute
// loop, so nothing downstream changes: the bulk copy lands n bf16 values at byte offset 2n of the slot (= its upper half
// for a full tile); element i is read at 2n + 2i and written at 4i, front to back, a chunk of 256 elements per step
// (reads of a step happen before its writes; a step's writes end where the next step's reads begin, at the latest).
// The backward kernel narrows the float32 gradient tile the same way (round to nearest even) and stores n * 2 bytes.
__device__ __forceinline__ float bf16_bits_to_f32(unsigned short b) { return __uint_as_float(static_cast<unsigned>(b) << 16); }
__device__ __forceinline__ unsigned short f32_to_bf16_bits(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
__device__ __forceinline__ void widen_bf16_inplace(float* slot, int n, int lane) {  // n % 8 == 0
  const uint2* src = reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(slot) + 2 * n);  // 4 bf16 per uint2
  float4* dst = reinterpret_cast<float4*>(slot);
  const int nq = n >> 2;
  for (int base = 0; base < nq; base += 64) {
    const int i0 = base + lane, i1 = base + 32 + lane;
    uint2 v0 = make_uint2(0u, 0u), v1 = make_uint2(0u, 0u);
    if (i0 < nq) v0 = src[i0];
    if (i1 < nq) v1 = src[i1];
    __syncwarp();
    if (i0 < nq)
      dst[i0] = make_float4(__uint_as_float(v0.x << 16), __uint_as_float(v0.x & 0xffff0000u), __uint_as_float(v0.y << 16),
                            __uint_as_float(v0.y & 0xffff0000u));
    if (i1 < nq)
      dst[i1] = make_float4(__uint_as_float(v1.x << 16), __uint_as_float(v1.x & 0xffff0000u), __uint_as_float(v1.y << 16),
                            __uint_as_float(v1.y & 0xffff0000u));
    __syncwarp();
  }
}
__device__ __forceinline__ void narrow_bf16_inplace(float* slot, int n, int lane) {  // n % 8 == 0
  const float4* src = reinterpret_cast<const float4*>(slot);
  uint2* dst = reinterpret_cast<uint2*>(slot);
  const int nq = n >> 2;
  for (int base = 0; base < nq; base += 64) {
    const int i0 = base + lane, i1 = base + 32 + lane;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (i0 < nq) v0 = src[i0];
    if (i1 < nq) v1 = src[i1];
    __syncwarp();
    if (i0 < nq) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(v0.x, v0.y), b = __floats2bfloat162_rn(v0.z, v0.w);
      dst[i0] = make_uint2(*reinterpret_cast<const unsigned*>(&a), *reinterpret_cast<const unsigned*>(&b));
    }
    if (i1 < nq) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(v1.x, v1.y), b = __floats2bfloat162_rn(v1.z, v1.w);
      dst[i1] = make_uint2(*reinterpret_cast<const unsigned*>(&a), *reinterpret_cast<const unsigned*>(&b));
    }
    __syncwarp();
  }
}
// parameter `idx` of a row in GLOBAL memory (the rare log-domain fallback reads the row where it lies)
__device__ __forceinline__ float ld_param(const float* row, int idx, bool bf16) {
  return bf16 ? bf16_bits_to_f32(reinterpret_cast<const unsigned short*>(row)[idx]) : row[idx];
}
// row `i` of the parameter tensor in global memory
__device__ __forceinline__ const float* param_row(const ModlArgs& a, long long i, int rowf) {
  return a.bf16 ? reinterpret_cast<const float*>(reinterpret_cast<const unsigned short*>(a.params) + i * rowf)
                : a.params + i * rowf;
}

struct Pixel {  // one pixel: both halves of a packed register see the same observation
  float x[3];
  bool left[3], right[3];
};
struct PixelPair {  // two pixels: lo half = pixel A, hi half = pixel B (the pixel-pair kernel for small n_mix)
  f2 x[3];
  bool ll[3], lh[3], rl[3], rh[3];
};
struct EdgeFlags {  // x at the lowest / highest bin, per half
  bool ll, lh, rl, rh;
};
__device__ __forceinline__ f2 px_x(const Pixel& p, int c) { return sp(p.x[c]); }
__device__ __forceinline__ f2 px_x(const PixelPair& p, int c) { return p.x[c]; }
__device__ __forceinline__ EdgeFlags px_edge(const Pixel& p, int c) {
  return EdgeFlags{p.left[c], p.left[c], p.right[c], p.right[c]};
}
__device__ __forceinline__ EdgeFlags px_edge(const PixelPair& p, int c) {
  return EdgeFlags{p.ll[c], p.lh[c], p.rl[c], p.rh[c]};
}

// pixel `pix` of image n (image n is scored against x[n % x_batch], include/vaemdl.h)
__device__ __forceinline__ void load_pixel(const ModlArgs& a, long long n, int pix, Pixel& px) {
  const long long xb = a.x_batch == 1 ? 0 : (n < a.x_batch ? n : n % a.x_batch);
  const long long xo = (xb * a.HW + pix) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v;
    if (a.x_u8) {
      v = u8_to_unit(static_cast<const uint8_t*>(a.x)[xo + c]);  // utils/data.py:15-16
    } else {
      v = static_cast<const float*>(a.x)[xo + c];
    }
    if (a.x_unit) v = __fmaf_rn(v, 2.0f, -1.0f);  // utils/mdl.py:65
    px.x[c] = v;
    px.left[c] = a.edge_openai ? (v < -0.999f) : (v <= -1.0f);
    px.right[c] = a.edge_openai ? (v > 0.999f) : (v >= 1.0f);
  }
}

// ---- rare fallback: one mixture's  logit + sum_c log f_c  in the log domain, straight from global memory -------
static __device__ __noinline__ float modl_logt(const float* __restrict__ row, int M, int m, const Pixel& px, bool plain,
                                               bool bf16 = false) {
  const float k0 = tanhf(ld_param(row, M + 2 * M + m, bf16));
  const float k1 = tanhf(ld_param(row, M + 3 * M + 2 * M + m, bf16));
  const float k2 = tanhf(ld_param(row, M + 6 * M + 2 * M + m, bf16));
  float t = ld_param(row, m, bf16);
  float a0 = px.x[0], a1 = px.x[1];  // what the green / blue means are chained on
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float loc = ld_param(row, M + c * 3 * M + m, bf16);
    if (c == 1) loc = loc + k0 * a0;
    if (c == 2) loc = loc + k1 * a0 + k2 * a1;
    if (plain) {  // utils/mdl_plain.py:160-162
      if (c == 0) a0 = loc;
      if (c == 1) a1 = loc;
    }
    const float ls = fmaxf(ld_param(row, M + c * 3 * M + M + m, bf16), -7.0f);
    t += subpix_logf(px.x[c], px.left[c], px.right[c], loc, ls, kDx, kWidth);
  }
  return t;
}
// log sum_m exp(logit_m + sum_c log f)  and  log sum_m exp(logit_m)
static __device__ __noinline__ void modl_pixel_logdomain(const float* __restrict__ row, int M, const Pixel& px, bool plain,
                                                  float& lse_t, float& lse_l, bool bf16 = false) {
  float mt = -INFINITY, ml = -INFINITY;
  for (int m = 0; m < M; ++m) {
    mt = fmaxf(mt, modl_logt(row, M, m, px, plain, bf16));
    ml = fmaxf(ml, ld_param(row, m, bf16));
  }
  float st = 0.f, sl = 0.f;
  for (int m = 0; m < M; ++m) {
    st += expf(modl_logt(row, M, m, px, plain, bf16) - mt);
    sl += expf(ld_param(row, m, bf16) - ml);
  }
  lse_t = mt + logf(st);
  lse_l = ml + logf(sl);
}

// ---- group (LPP lanes of one pixel) all-reduce with a fixed summation order --------------------------------------
template <int LPP>
__device__ __forceinline__ float group_sum(float v, int lane) {
  if constexpr (LPP == 1) {
    return v;
  } else {
    const int base = lane - (lane % LPP);
    float s = __shfl_sync(kFull, v, base);
#pragma unroll
    for (int j = 1; j < LPP; ++j) s += __shfl_sync(kFull, v, (base + j) & 31);
    return s;
  }
}
template <int LPP>
__device__ __forceinline__ float group_max(float v, int lane) {
  if constexpr (LPP == 1) {
    return v;
  } else {
    const int base = lane - (lane % LPP);
    float s = __shfl_sync(kFull, v, base);
#pragma unroll
    for (int j = 1; j < LPP; ++j) s = fmaxf(s, __shfl_sync(kFull, v, (base + j) & 31));
    return s;
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// log2 with the exponent split off: lg2.approx is only accurate to 2^-22 RELATIVE outside (0.5, 2); on the mantissa
// its error is 2^-22 absolute, which keeps the per-pixel log-prob good to ~2e-7 instead of ~5e-6.
__device__ __forceinline__ float lg2_split(float v) {
  const int bits = __float_as_int(v);
  const int e = (bits >> 23) - 127;
  const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
  return static_cast<float>(e) + lg2a(m);
}

// ---- one mixture component, scalar (any-M kernel) ------------------------------------------------------------------------------------------
// forward: returns P = prod_c f_c (linear domain)
__device__ __forceinline__ float mix_fwd(const Pixel& px, const float mu[3], const float s[3], const float kap[3],
                                         bool plain) {
  float k0, k1, k2;
  tanh3(kap[0], kap[1], kap[2], k0, k1, k2);  // utils/mdl.py:110
  const float loc0 = mu[0];
  const float a0 = plain ? loc0 : px.x[0];
  const float loc1 = fmaf(k0, a0, mu[1]);                            // utils/mdl.py:140 | utils/mdl_plain.py:161
  const float a1 = plain ? loc1 : px.x[1];
  const float loc2 = fmaf(k2, a1, fmaf(k1, a0, mu[2]));              // utils/mdl.py:141-145 | utils/mdl_plain.py:162
  SubF f0, f1, f2;
  subpix<false>(px.x[0], px.left[0], px.right[0], loc0, fmaxf(s[0], -7.0f), kDx, kWidth, f0);
  subpix<false>(px.x[1], px.left[1], px.right[1], loc1, fmaxf(s[1], -7.0f), kDx, kWidth, f1);
  subpix<false>(px.x[2], px.left[2], px.right[2], loc2, fmaxf(s[2], -7.0f), kDx, kWidth, f2);
  return (f0.num * f1.num * f2.num) * rcpa(f0.den * f1.den * f2.den);
}

// backward: returns P and the nine d log P / d(param) values u = {dmuR dsR dkR dmuG dsG dkG dmuB dsB dkB}
__device__ __forceinline__ float mix_bwd(const Pixel& px, const float mu[3], const float s[3], const float kap[3],
                                         float u[9], bool plain) {
  float k[3];
  tanh3(kap[0], kap[1], kap[2], k[0], k[1], k[2]);
  float loc[3];
  loc[0] = mu[0];
  const float a0 = plain ? loc[0] : px.x[0];
  loc[1] = fmaf(k[0], a0, mu[1]);
  const float a1 = plain ? loc[1] : px.x[1];
  loc[2] = fmaf(k[2], a1, fmaf(k[1], a0, mu[2]));
  SubB f[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) subpix<true>(px.x[c], px.left[c], px.right[c], loc[c], fmaxf(s[c], -7.0f), kDx, kWidth, f[c]);
  const float d01 = f[0].den * f[1].den;
  const float R = rcpa(d01 * f[2].den);
  float rd[3];
  rd[0] = f[1].den * f[2].den * R;
  rd[1] = f[0].den * f[2].den * R;
  rd[2] = d01 * R;
  float dloc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float Dm = f[c].nm * rd[c];
    dloc[c] = -f[c].inv * Dm;
    float dls = (f[c].dir - f[c].c0) - fmaf(f[c].mid, Dm, f[c].nh * rd[c]);
    if (!(s[c] >= -7.0f)) dls = 0.0f;  // tf.maximum(logscale, -7) routes the gradient to logscale iff logscale >= -7
    u[3 * c + 0] = dloc[c];
    u[3 * c + 1] = dls;
  }
  // coefficients: loc_g = mu_g + k0 x_r ; loc_b = mu_b + k1 x_r + k2 x_g ; d tanh = 1 - tanh^2
  if (plain) {  // the chain runs through the means: total derivatives w.r.t. loc_g, loc_r pick up the downstream terms
    dloc[1] = fmaf(k[2], dloc[2], dloc[1]);
    dloc[0] = fmaf(k[0], dloc[1], fmaf(k[1], dloc[2], dloc[0]));
    u[0] = dloc[0];
    u[3] = dloc[1];
  }
  u[2] = dloc[1] * a0 * fmaf(-k[0], k[0], 1.0f);
  u[5] = dloc[2] * a0 * fmaf(-k[1], k[1], 1.0f);
  u[8] = dloc[2] * a1 * fmaf(-k[2], k[2], 1.0f);
  return (f[0].num * f[1].num * f[2].num) * R;
}

// ---- packed (two mixture components per instruction) sub-pixel arithmetic --------------------------------------------
// Same formulas as subpix<> in modl_math.cuh; `lo` half = component m, `hi` half = component m+1.
// NARROW is a warp-uniform property of the pair (some lane has a log-scale below kLsNarrow, i.e. h >= kHSmall): only
// then are exp(-h) and h*coth(h) evaluated on the MUFU pipe, otherwise short polynomials on the packed FMA pipe.
constexpr float kLsNarrow = -3.6441f;  // -log(kHSmall * 255) rounded towards 0: ls <= this  <=>  h = exp(-ls)/255 >= kHSmall (conservatively)

struct Sub2 {
  f2 num, den;               // f = num / den
  f2 nm, nh, c0, dir;        // backward numerators (see SubB)
  f2 inv, mid;
};

template <bool NARROW, bool BWD>
__device__ __forceinline__ void subpix2(f2 x, EdgeFlags e, f2 loc, f2 s_raw, Sub2& o) {
  const f2 ls = max_2(s_raw, -7.0f);                                  // utils/mdl.py:109
  const f2 inv = ex2_2(ls * (-kLog2e));
  const f2 mid = inv * (x - loc);
  const f2 A = ex2_negabs_2(mid * kLog2e);
  const f2 h = inv * kDx;
  f2 q = fma2(h, -1.0f / 720.0f, 1.0f / 120.0f);
  q = fma2(h, q, -1.0f / 24.0f);
  q = fma2(h, q, 1.0f / 6.0f);
  q = fma2(h, q, -0.5f);
  q = fma2(h, q, 1.0f);
  f2 omG = h * q;
  f2 G = sp(1.0f) - omG;
  bool nl = false, nh_ = false;
  if constexpr (NARROW) {
    const f2 Ge = ex2_2(h * (-kLog2e));
    const f2 omGe = sp(1.0f) - Ge;
    nl = lo(h) >= kHSmall;
    nh_ = hi(h) >= kHSmall;
    G = sel_2(nl, nh_, Ge, G);
    omG = sel_2(nl, nh_, omGe, omG);
  }
  const f2 AG = A * G;
  const f2 ApG = A + G;
  const f2 opAG = AG + 1.0f;
  const f2 opA = A + 1.0f;
  const f2 opG = G + 1.0f;
  const f2 rest_n = omG * opG;
  const f2 num_n = A * rest_n;
  const f2 den_n = ApG * opAG;
  const f2 thr = den_n * 1e-5f;
  const bool il = lo(num_n) > lo(thr), ih = hi(num_n) > hi(thr);    // sigmoid(p)-sigmoid(q) > 1e-5 (utils/mdl.py:193)
  const f2 num_l = (A * inv) * kWidth;
  const f2 den_l = opA * opA;
  f2 num = sel_2(il, ih, num_n, num_l);
  f2 den = sel_2(il, ih, den_n, den_l);
  const bool el = e.ll || e.rl, eh = e.lh || e.rh;
  const bool ool = (e.ll == (lo(mid) >= 0.0f)), ooh = (e.lh == (hi(mid) >= 0.0f));  // 1/(1+AG) vs A/(A+G)
  if (el || eh) {
    num = sel_2(el, eh, sel_2(ool, ooh, sp(1.0f), A), num);
    den = sel_2(el, eh, sel_2(ool, ooh, opAG, ApG), den);
  }
  o.num = num;
  o.den = den;
  if constexpr (BWD) {
    const f2 omA2 = (sp(1.0f) - A) * opA;
    f2 nm = neg_sign_of_2(sel_2(il, ih, G * omA2, omA2), mid);
    f2 nh = sel_2(il, ih, (h * -1.0f) * num_n, sp(0.0f));
    const f2 h2 = h * h;
    f2 hc = fma2(h2, 2.0f / 945.0f, -1.0f / 45.0f);
    hc = fma2(h2, hc, 1.0f / 3.0f);
    hc = fma2(h2, hc, 1.0f);
    if constexpr (NARROW) {
      const f2 e = h * fma2(G, G, 1.0f) * rcp_2(rest_n);
      hc = sel_2(nl, nh_, e, hc);
    }
    f2 c0 = sel_2(il, ih, hc, sp(0.0f));
    f2 dir = sel_2(il, ih, sp(0.0f), sp(-1.0f));
    if (el || eh) {
      const f2 t = sel_2(ool, ooh, AG, G);
      nm = sel_2(el, eh, pk(e.ll ? lo(t) : -lo(t), e.lh ? hi(t) : -hi(t)), nm);
      nh = sel_2(el, eh, h * t, nh);
      c0 = sel_2(el, eh, sp(0.0f), c0);
      dir = sel_2(el, eh, sp(0.0f), dir);
    }
    o.nm = nm;
    o.nh = nh;
    o.c0 = c0;
    o.dir = dir;
    o.inv = inv;
    o.mid = mid;
  }
}

// tanh of three coefficient pairs: 6 ex2 + 2 rcp
__device__ __forceinline__ void tanh3_2(const f2 kp[3], f2 k[3]) {
  const float c = 2.0f * kLog2e;
  const f2 E0 = min_2(ex2_2(kp[0] * c), 1073741824.0f);
  const f2 E1 = min_2(ex2_2(kp[1] * c), 1073741824.0f);
  const f2 E2 = min_2(ex2_2(kp[2] * c), 1073741824.0f);
  const f2 d0 = E0 + 1.0f, d1 = E1 + 1.0f, d2 = E2 + 1.0f;
  const f2 d01 = d0 * d1;
  const f2 R = rcp_2(d01 * d2);
  k[0] = (E0 + -1.0f) * (d1 * d2) * R;
  k[1] = (E1 + -1.0f) * (d0 * d2) * R;
  k[2] = (E2 + -1.0f) * d01 * R;
}

// One pair of mixture components.  Returns P = prod_c f_c (linear domain); BWD also the nine d log P / d param pairs.
// AR = 0: green / blue means chained on the OBSERVED x (utils/mdl.py:139-145, PixelCNN++);
// AR = 1: chained on the component's own means (utils/mdl_plain.py:160-162, no conditioning on x).
template <bool NARROW, bool BWD, typename PX, int AR>
__device__ __forceinline__ f2 pair_eval(const PX& px, const f2 mu[3], const f2 s[3], const f2 kp[3], f2 u[9]) {
  f2 k[3];
  tanh3_2(kp, k);
  f2 loc[3];
  loc[0] = mu[0];
  const f2 x0 = AR ? loc[0] : px_x(px, 0);
  loc[1] = fma2(k[0], x0, mu[1]);                               // utils/mdl.py:140 | utils/mdl_plain.py:161
  const f2 x1 = AR ? loc[1] : px_x(px, 1);
  loc[2] = fma2(k[2], x1, fma2(k[1], x0, mu[2]));               // utils/mdl.py:141-145 | utils/mdl_plain.py:162
  Sub2 f[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) subpix2<NARROW, BWD>(px_x(px, c), px_edge(px, c), loc[c], s[c], f[c]);
  const f2 d01 = f[0].den * f[1].den;
  const f2 R = rcp_2(d01 * f[2].den);
  const f2 P = (f[0].num * f[1].num) * (f[2].num * R);
  if constexpr (BWD) {
    f2 rd[3];
    rd[0] = f[1].den * f[2].den * R;
    rd[1] = f[0].den * f[2].den * R;
    rd[2] = d01 * R;
    f2 dloc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const f2 Dm = f[c].nm * rd[c];
      dloc[c] = (f[c].inv * -1.0f) * Dm;
      f2 dls = (f[c].dir - f[c].c0) - fma2(f[c].mid, Dm, f[c].nh * rd[c]);
      // tf.maximum(logscale, -7): the gradient reaches logscale iff logscale >= -7
      dls = sel_2(lo(s[c]) >= -7.0f, hi(s[c]) >= -7.0f, dls, sp(0.0f));
      u[3 * c + 0] = dloc[c];
      u[3 * c + 1] = dls;
    }
    if constexpr (AR != 0) {  // total derivatives through the chain of means
      dloc[1] = fma2(k[2], dloc[2], dloc[1]);
      dloc[0] = fma2(k[0], dloc[1], fma2(k[1], dloc[2], dloc[0]));
      u[0] = dloc[0];
      u[3] = dloc[1];
    }
    u[2] = (dloc[1] * x0) * fma2(k[0] * -1.0f, k[0], 1.0f);
    u[5] = (dloc[2] * x0) * fma2(k[1] * -1.0f, k[1], 1.0f);
    u[8] = (dloc[2] * x1) * fma2(k[2] * -1.0f, k[2], 1.0f);
  }
  return P;
}

// ---- the tiled kernel -------------------------------------------------------------------------------------------------
// M = MC * LPP mixtures; LPP lanes share a pixel, each owning MC consecutive components, processed two at a time.
template <int MC, int LPP>
struct Tile {
  static constexpr int M = MC * LPP;
  static constexpr int PPT = 32 / LPP;  // pixels per warp tile
  static constexpr int ROWF = 10 * M;
  static constexpr int TILE_F = PPT * ROWF;
  static constexpr int TILE_B = TILE_F * 4;
  static constexpr int AUX_F = PPT * M;  // backward: W*P per (pixel, component)
  static constexpr int NPAIR = (MC + 1) / 2;
  static constexpr bool ALIGNED = (M % 2 == 0) && (MC % 2 == 0);  // component pairs sit on 8-byte boundaries
  static_assert(TILE_B % 16 == 0, "bulk copies need 16-byte multiples");
};

// a pair of consecutive floats at row[off], row[off+1]; `single`: only row[off] exists (odd MC), both halves get it
template <bool ALIGNED>
__device__ __forceinline__ f2 ld_pair(const float* row, int off, bool single) {
  if constexpr (ALIGNED) {
    const float2 t = *reinterpret_cast<const float2*>(row + off);
    return pk(t.x, t.y);
  } else {
    const float a = row[off];
    const float b = single ? a : row[off + 1];
    return pk(a, b);
  }
}
template <bool ALIGNED>
__device__ __forceinline__ void st_pair(float* row, int off, bool single, f2 v) {
  if constexpr (ALIGNED) {
    *reinterpret_cast<float2*>(row + off) = make_float2(lo(v), hi(v));
  } else {
    row[off] = lo(v);
    if (!single) row[off + 1] = hi(v);
  }
}

struct PixRaw {
  unsigned v[3];
};
__device__ __forceinline__ PixRaw load_pixel_raw(const ModlArgs& a, long long n, int pix) {
  const long long xb = a.x_batch == 1 ? 0 : (n < a.x_batch ? n : n % a.x_batch);
  const long long xo = (xb * a.HW + pix) * 3;
  PixRaw r;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    r.v[c] = a.x_u8 ? static_cast<unsigned>(static_cast<const uint8_t*>(a.x)[xo + c])
                    : __float_as_uint(static_cast<const float*>(a.x)[xo + c]);
  return r;
}
__device__ __forceinline__ void decode_pixel(const ModlArgs& a, const PixRaw& r, Pixel& px) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = a.x_u8 ? u8_to_unit(r.v[c]) : __uint_as_float(r.v[c]);  // utils/data.py:15-16
    if (a.x_unit) v = __fmaf_rn(v, 2.0f, -1.0f);                                                  // utils/mdl.py:65
    px.x[c] = v;
    px.left[c] = a.edge_openai ? (v < -0.999f) : (v <= -1.0f);
    px.right[c] = a.edge_openai ? (v > 0.999f) : (v >= 1.0f);
  }
}

// FUSED: the forward and the backward pass of one step run inside ONE cooperative kernel (modl_step_kernel): both use
// the backward shared-memory layout, the mbarrier is initialised once and its phase carries over, the forward pass
// leaves its last tile in the slot and the (reversed) backward pass starts on it without loading anything.
// PD = 1: bfloat16 parameters / gradient in global memory (widened / narrowed in place in the slot, see widen_bf16_inplace)
template <int MC, int LPP, bool BWD, int NSLOT, int AR, bool FUSED, int PD = 0>
__device__ __forceinline__ void tile_body(const ModlArgs& a, unsigned char* smem_raw) {
  using T = Tile<MC, LPP>;
  constexpr int M = T::M, PPT = T::PPT, ROWF = T::ROWF, TILE_F = T::TILE_F, NPAIR = T::NPAIR;
  constexpr bool AL = T::ALIGNED;
  constexpr int WARP_F = NSLOT * TILE_F + ((BWD || FUSED) ? T::AUX_F : 0);
  static_assert(!FUSED || NSLOT == 1, "the fused step keeps one slot per warp");
  static_assert(PD == 0 || (NSLOT == 1 && !FUSED), "bf16 parameters: one slot per warp, three-launch step");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* slots = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * WARP_F;
  float* aux = slots + NSLOT * TILE_F;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(nwarps) * WARP_F * 4) + warp * NSLOT;

  if constexpr (!(FUSED && BWD)) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) mbar_init(&bars[s], 1);
      fence_barrier_init();
    }
    __syncwarp();
  }
  if constexpr (!FUSED) {
    if (!BWD && a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
    if constexpr (BWD)
      pdl_wait();     // launched programmatically behind the finish kernel: g_image (and, in general, the parameters) must be complete
    else
      pdl_trigger();  // let the finish kernel's launch overlap this kernel's tail
  }

  const long long gw = run_index(a, warp, nwarps);
  const bool lane_used = (lane / LPP) < PPT;
  const int p = lane_used ? (lane / LPP) : 0;  // idle lanes (LPP=3: lanes 30,31) shadow pixel 0
  const int sub = lane % LPP;
  const int m0 = sub * MC;

  // this warp's run of CONSECUTIVE tiles (balanced split of the tile range over all warps of the grid): per-image
  // sums then accumulate in registers across tiles and leave the warp once per image instead of once per tile
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);

  auto tile_rows = [&](long long t) -> int {
    const long long rem = a.n_px - t * PPT;
    return rem < PPT ? static_cast<int>(rem) : PPT;
  };
  // bring tile t into slot s (bulk copy when the byte count allows it, plain loads for a ragged tail tile)
  const long long t_cnt = t_end - t_begin;
  const bool rev = BWD && a.reverse;
  const long long t_first = rev ? t_end - 1 : t_begin;
  const long long t_dir = rev ? -1 : 1;
  const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();
  auto issue = [&](long long t, int s) {
    const int rows = tile_rows(t);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * (PD ? 2u : 4u);
    const char* src = reinterpret_cast<const char*>(a.params) + t * TILE_F * (PD ? 2 : 4);
    float* slot_f = slots + s * TILE_F;
    char* dst = reinterpret_cast<char*>(slot_f) + (PD ? bytes : 0u);  // bf16 lands behind the room its float32 image needs
    if ((bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars[s], bytes);
        if (BWD) {
          if (a.bwd_hint & 1)
            bulk_g2s_hint(dst, src, bytes, &bars[s], pol_first);
          else
            bulk_g2s(dst, src, bytes, &bars[s]);
        } else {
          if (a.keep_tiles > 0)
            bulk_g2s_hint(dst, src, bytes, &bars[s], (t_end - t) <= a.keep_tiles ? pol_last : pol_first);
          else
            bulk_g2s(dst, src, bytes, &bars[s]);
        }
      }
    } else {
      for (int i = lane; i < rows * ROWF; i += 32)
        slot_f[i] = PD ? bf16_bits_to_f32(reinterpret_cast<const unsigned short*>(src)[i]) : reinterpret_cast<const float*>(src)[i];
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(&bars[s], 0);
    }
  };

  // forward: every slot is in flight from the start; backward: slots are refilled one tile ahead (see below);
  // fused backward: the first tile (the forward pass's last) is already in the slot
  if constexpr (!(FUSED && BWD)) {
#pragma unroll
    for (int s = 0; s < (BWD ? 1 : NSLOT); ++s) {
      if (s < t_cnt) issue(t_first + s * t_dir, s);
    }
  }
  // loads this warp's barrier has completed before this pass (fused backward: the whole forward pass but the resident tile)
  const uint32_t phase0 = (FUSED && BWD) ? static_cast<uint32_t>(t_cnt - 1) : 0u;

  // (image, pixel-in-image) of this lane's pixel-sample, advanced incrementally: one 64-bit division per kernel
  const long long step_n = PPT / a.HW;
  const int step_pix = static_cast<int>(PPT - step_n * a.HW);
  long long n_own = (t_first * PPT + p) / a.HW;
  int pix_own = static_cast<int>((t_first * PPT + p) - n_own * a.HW);
  // float64 running sums of the image the warp is in (acc0, image n_base) and of the next one (acc1), per lane
  double acc0 = 0.0, acc1 = 0.0;
  const long long n_warp_first = (t_begin * PPT) / a.HW;
  long long n_base = n_warp_first;

  // software prefetch of the (L2-resident) pixel and upstream-gradient values one tile ahead
  auto fetch = [&](long long t, long long n_lane, int pix_lane, long long& n_out, long long& nfirst_out, PixRaw& raw,
                   float& g_out) {
    const int rows = tile_rows(t);
    const long long n_first = __shfl_sync(kFull, n_lane, 0);
    const int pix_first = __shfl_sync(kFull, pix_lane, 0);
    const bool in = p < rows;  // lanes past a ragged last tile shadow the tile's first pixel
    const long long n = in ? n_lane : n_first;
    const int pix = in ? pix_lane : pix_first;
    raw = load_pixel_raw(a, n, pix);
    g_out = 0.0f;
    if constexpr (BWD) {
      if (a.g_image) g_out = a.g_image[n];
      if (a.g_pixel) g_out += a.g_pixel[n * a.HW + pix];
    }
    n_out = n;
    nfirst_out = n_first;
  };

  long long n_cur = 0, nfirst_cur = 0;
  PixRaw raw_cur{};
  float g_cur = 0.0f;
  if (t_cnt > 0) fetch(t_first, n_own, pix_own, n_cur, nfirst_cur, raw_cur, g_cur);

  for (long long it = 0; it < t_cnt; ++it) {
    const long long t = t_first + it * t_dir;
    const int s = static_cast<int>(it % NSLOT);
    const uint32_t parity = (phase0 + static_cast<uint32_t>(it / NSLOT)) & 1u;
    const int rows = tile_rows(t);
    const int pp = p < rows ? p : 0;
    const bool active = lane_used && (p < rows);
    const long long i = t * PPT + pp;  // this lane's pixel-sample
    const long long n = n_cur, n_first = nfirst_cur;
    const float g = g_cur;
    Pixel px;
    decode_pixel(a, raw_cur, px);
    // advance the index and prefetch the next tile's pixel / upstream gradient
    if (!rev) {
      n_own += step_n;
      pix_own += step_pix;
      if (pix_own >= a.HW) {
        pix_own -= a.HW;
        ++n_own;
      }
    } else {
      n_own -= step_n;
      pix_own -= step_pix;
      if (pix_own < 0) {
        pix_own += a.HW;
        --n_own;
      }
    }
    if (it + 1 < t_cnt) fetch(t + t_dir, n_own, pix_own, n_cur, nfirst_cur, raw_cur, g_cur);

    float* slot = slots + s * TILE_F;
    float* rowp = slot + pp * ROWF;
    float* auxp = aux + pp * M;
    if (!(FUSED && BWD && it == 0)) mbar_wait(&bars[s], parity);
    if constexpr (PD != 0) {
      if (((rows * ROWF * 2) & 15) == 0) widen_bf16_inplace(slot, rows * ROWF, lane);  // (a ragged tile was widened by its loads)
    }
    if constexpr (BWD && NSLOT > 1) {
      // the other slot's gradient tile was handed to the TMA engine at the end of the previous iteration: once its
      // shared-memory reads are done, refill that slot with this warp's next tile (lands while this tile computes)
      if (it + 1 < t_cnt) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        issue(t + t_dir, s ^ 1);
      }
    }

    // W_m = exp(logit_m - max logit)
    float lmax = rowp[m0];
#pragma unroll
    for (int m = 1; m < MC; ++m) lmax = fmaxf(lmax, rowp[m0 + m]);
    lmax = group_max<LPP>(lmax, lane);

    f2 sumW2 = sp(0.0f), sumWP2 = sp(0.0f);
#pragma unroll 1
    for (int pr = 0; pr < NPAIR; ++pr) {
      const int m = m0 + 2 * pr;
      const bool single = (MC % 2 == 1) && (pr == NPAIR - 1);
      f2 lg = ld_pair<AL>(rowp, m, single);
      if (single) lg = pk(lo(lg), -INFINITY);  // the padding half gets zero weight
      f2 mu[3], sc[3], kp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mu[c] = ld_pair<AL>(rowp, (1 + 3 * c) * M + m, single);
        sc[c] = ld_pair<AL>(rowp, (2 + 3 * c) * M + m, single);
        kp[c] = ld_pair<AL>(rowp, (3 + 3 * c) * M + m, single);
      }
      const float smin = fminf(fminf(fminf(lo(sc[0]), hi(sc[0])), fminf(lo(sc[1]), hi(sc[1]))), fminf(lo(sc[2]), hi(sc[2])));
      const bool narrow = __any_sync(kFull, smin < kLsNarrow);
      const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
      f2 u[9];
      f2 P;
      if (narrow)
        P = pair_eval<true, BWD, Pixel, AR>(px, mu, sc, kp, u);
      else
        P = pair_eval<false, BWD, Pixel, AR>(px, mu, sc, kp, u);
      sumW2 = sumW2 + W;
      sumWP2 = fma2(W, P, sumWP2);
      if constexpr (BWD) {
        // unscaled gradients overwrite the component's parameters in place; W*P goes to the aux strip.
        // Only lanes that own a real pixel write: shadow lanes (ragged tile, or lanes 30/31 when 3 lanes share a
        // pixel) would otherwise race with the owner of pixel 0.
        if (active) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          st_pair<AL>(rowp, (1 + 3 * c) * M + m, single, u[3 * c + 0]);
          st_pair<AL>(rowp, (2 + 3 * c) * M + m, single, u[3 * c + 1]);
          st_pair<AL>(rowp, (3 + 3 * c) * M + m, single, u[3 * c + 2]);
        }
        st_pair<AL>(auxp, m, single, W * P);
        }
      }
    }
    const float S = group_sum<LPP>(lo(sumWP2) + hi(sumWP2), lane);
    const float SW = group_sum<LPP>(lo(sumW2) + hi(sumW2), lane);
    const bool tiny = !(S > kTinySum);  // also catches NaN
    const float* grow = param_row(a, i, ROWF);

    if constexpr (!BWD) {
      if constexpr (PD != 0) fence_async_smem();  // the widening wrote the slot through the generic proxy
      __syncwarp();
      {  // every lane has read its row: re-arm the slot for this warp's tile NSLOT iterations ahead
        if (it + NSLOT < t_cnt) issue(t + NSLOT * t_dir, s);
      }
      float lp = (lg2_split(S) - lg2_split(SW)) * kLn2;  // utils/mdl.py:78-89 in one step
      if (tiny) {
        float lt, ll;
        modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, PD != 0);
        lp = lt - ll;
      }
      const bool owner = active && sub == 0;
      if (a.lp_pixel && owner) a.lp_pixel[i] = lp;
      const float val = owner ? lp : 0.0f;
      if (a.partial) {
        // float64 from here on: the per-image sums (~ -2e4 nats) feed a softmax over importance samples.
        // A tile holds pixels of at most two images (HW >= PPT on this route): n_first and n_first + 1.
        while (n_base < n_first) {  // the warp has left image n_base: its sum leaves the registers (warp-uniform)
          const double done = warp_sum(acc0);
          if (lane == 0) a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = done;
          acc0 = acc1;
          acc1 = 0.0;
          ++n_base;
        }
        if (n == n_base)
          acc0 += static_cast<double>(val);
        else
          acc1 += static_cast<double>(val);
      } else if (a.ll_atomic) {
        if (owner) atomicAdd(a.ll_atomic + n, static_cast<double>(val));
      }
    } else {
      const float rS = rcpa(S), rSW = rcpa(SW);
      float lt = 0.f, ll = 0.f;
      if (tiny) modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, PD != 0);
#pragma unroll 1
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int m = m0 + 2 * pr;
        const bool single = (MC % 2 == 1) && (pr == NPAIR - 1);
        f2 lg = ld_pair<AL>(rowp, m, single);
        const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
        const f2 wp = ld_pair<AL>(auxp, m, single);
        f2 r = wp * rS;     // posterior responsibility of the component
        f2 pi = W * rSW;    // softmax(logits)
        if (tiny) {
          r = pk(expf(modl_logt(grow, M, m, px, a.plain != 0, PD != 0) - lt),
                 single ? 0.0f : expf(modl_logt(grow, M, m + 1, px, a.plain != 0, PD != 0) - lt));
          pi = pk(expf(ld_param(grow, m, PD != 0) - ll), single ? 0.0f : expf(ld_param(grow, m + 1, PD != 0) - ll));
        }
        const f2 gr = r * g;
        if (active) {
          st_pair<AL>(rowp, m, single, (r - pi) * g);
#pragma unroll
          for (int j = 1; j < 10; ++j) st_pair<AL>(rowp, j * M + m, single, ld_pair<AL>(rowp, j * M + m, single) * gr);
        }
      }
      // hand the gradient tile to the TMA engine
      const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * (PD ? 2u : 4u);
      char* dst = reinterpret_cast<char*>(a.dparams) + t * TILE_F * (PD ? 2 : 4);
      if ((bytes & 15u) == 0) {
        if constexpr (PD != 0) {
          __syncwarp();
          narrow_bf16_inplace(slot, rows * ROWF, lane);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (a.bwd_hint & 2)
            bulk_s2g_hint(dst, slot, bytes, pol_first);
          else
            bulk_s2g(dst, slot, bytes);
          bulk_commit();
        }
      } else {
        __syncwarp();
        for (int q = lane; q < rows * ROWF; q += 32) {
          if (PD)
            reinterpret_cast<unsigned short*>(dst)[q] = f32_to_bf16_bits(slot[q]);
          else
            reinterpret_cast<float*>(dst)[q] = slot[q];
        }
        __syncwarp();
      }
      if constexpr (NSLOT == 1) {
        if (it + 1 < t_cnt) {
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
          issue(t + t_dir, 0);
        }
      }
    }
  }
  if constexpr (BWD) {
    if (lane == 0) bulk_wait_all<0>();
  } else {
    if (a.partial && t_begin < t_end) {
      const long long n_last = (t_end * PPT < a.n_px ? t_end * PPT - 1 : a.n_px - 1) / a.HW;  // image of the warp's last pixel-sample
      const double d0 = warp_sum(acc0), d1 = warp_sum(acc1);
      if (lane == 0) {
        a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
        if (n_base + 1 <= n_last) a.partial[partial_slot(n_base + 1, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
      }
    }
  }
}

template <int MC, int LPP, bool BWD, int NSLOT, int MAXT, int AR, int PD = 0>
__global__ void __launch_bounds__(MAXT, 1) modl_tile_kernel(const ModlArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  tile_body<MC, LPP, BWD, NSLOT, AR, false, PD>(a, smem_raw);
}

// ---- one step in one launch: forward -> grid barrier -> per-image sums, log-mean-exp, softmax weights -> grid barrier ->
// backward.  For training shapes small enough that launch boundaries and pipeline ramps dominate (BASELINE configs[0]:
// 131 MB of parameters, ~4 tiles per warp): no launch gaps, one ramp instead of three, and each warp's last forward tile
// is still in shared memory when its reversed backward run starts.  Cooperative launch (all CTAs co-resident).
struct StepArgs {
  ModlArgs a;
  StepFinish f;
};

template <int MC, int LPP, int AR>
__global__ void __launch_bounds__(512, 1) modl_step_kernel(const StepArgs sa) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int lane = threadIdx.x & 31;
  const long long nwarps = blockDim.x >> 5;
  const long long gw = static_cast<long long>(blockIdx.x) * nwarps + (threadIdx.x >> 5);
  const long long total_warps = static_cast<long long>(gridDim.x) * nwarps;
  tile_body<MC, LPP, false, 1, AR, true>(sa.a, smem_raw);
  __threadfence();
  grid.sync();
  step_finish(sa.f, gw, total_warps, lane);
  __threadfence();
  grid.sync();
  if (sa.f.elbo && gw == total_warps - 1) {  // batch mean, fixed order (the last warp owns the shortest run)
    double t = 0.0;
    for (long long b = lane; b < sa.f.B; b += 32) t += sa.f.lme64[b];
    t = warp_sum(t);
    if (lane == 0) sa.f.elbo[0] = static_cast<float>(t / static_cast<double>(sa.f.b_norm));  // models/loss.py:37
  }
  tile_body<MC, LPP, true, 1, AR, true>(sa.a, smem_raw);
}

static __global__ void cast_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]);
}

// ---- the pixel-pair kernel (small n_mix, e.g. the reference's own default n_mix = 5) ---------------------------------------
// Same pipeline as modl_tile_kernel (per-warp TMA bulk loads, runs of consecutive tiles, in-place gradient staging), but
// the two halves of a packed register hold the SAME mixture component of TWO pixels: lane l owns rows l and l + 32 of a
// 64-row tile.  Nothing is wasted on an odd component count, the tile is as large as the n_mix = 10 one (12.8 KB at
// n_mix = 5), and the two rows of a lane sit 32 rows apart so that the scalar shared-memory loads spread over the banks.
template <int M>
struct TilePP {
  static constexpr int PPT = 64;
  static constexpr int ROWF = 10 * M;
  static constexpr int TILE_F = PPT * ROWF;
  static constexpr int TILE_B = TILE_F * 4;
  static constexpr int AUX_F = PPT * M;
};

__device__ __forceinline__ Pixel half_pixel(const PixelPair& pp, bool hi_half) {
  Pixel px;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    px.x[c] = hi_half ? hi(pp.x[c]) : lo(pp.x[c]);
    px.left[c] = hi_half ? pp.lh[c] : pp.ll[c];
    px.right[c] = hi_half ? pp.rh[c] : pp.rl[c];
  }
  return px;
}

template <int M, bool BWD, int MAXT, int AR>
__global__ void __launch_bounds__(MAXT, 1) modl_pp_kernel(const ModlArgs a) {
  using T = TilePP<M>;
  constexpr int PPT = T::PPT, ROWF = T::ROWF, TILE_F = T::TILE_F;
  constexpr int WARP_F = TILE_F + (BWD ? T::AUX_F : 0);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* slot = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * WARP_F;
  float* aux = slot + TILE_F;
  // Even n_mix: the row stride 10*M words shares a large power of two with the 32 banks (16-way conflicts at M = 8), so
  // every group of ROTB lanes walks the components in its own rotated order (<= 2-way for every M; the order of the
  // per-pixel sum then depends on the lane, the result stays reproducible run to run).
  constexpr int ROTB = (M % 2 != 0 || M < 2) ? 0 : (M == 8 ? 4 : 8);
  const int rot0 = ROTB ? (lane / (ROTB ? ROTB : 1)) % M : 0;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(nwarps) * WARP_F * 4) + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (!BWD && a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
  if constexpr (BWD)
    pdl_wait();
  else
    pdl_trigger();

  const long long gw = run_index(a, warp, nwarps);
  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);
  const long long t_cnt = t_end - t_begin;
  const bool rev = BWD && a.reverse;
  const long long t_first = rev ? t_end - 1 : t_begin;
  const long long t_dir = rev ? -1 : 1;
  const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();

  auto tile_rows = [&](long long t) -> int {
    const long long rem = a.n_px - t * PPT;
    return rem < PPT ? static_cast<int>(rem) : PPT;
  };
  auto issue = [&](long long t) {
    const int rows = tile_rows(t);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * 4u;
    const float* src = a.params + t * TILE_F;
    if ((bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        if (BWD) {
          if (a.bwd_hint & 1)
            bulk_g2s_hint(slot, src, bytes, bar, pol_first);
          else
            bulk_g2s(slot, src, bytes, bar);
        } else {
          if (a.keep_tiles > 0)
            bulk_g2s_hint(slot, src, bytes, bar, (t_end - t) <= a.keep_tiles ? pol_last : pol_first);
          else
            bulk_g2s(slot, src, bytes, bar);
        }
      }
    } else {
      for (int i = lane; i < rows * ROWF; i += 32) slot[i] = src[i];
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar, 0);
    }
  };
  if (t_cnt > 0) issue(t_first);

  // (image, pixel-in-image) of this lane's FIRST pixel-sample (row `lane` of the tile), advanced incrementally
  const long long step_n = PPT / a.HW;
  const int step_pix = static_cast<int>(PPT - step_n * a.HW);
  long long n_own = (t_first * PPT + lane) / a.HW;
  int pix_own = static_cast<int>((t_first * PPT + lane) - n_own * a.HW);
  double acc0 = 0.0, acc1 = 0.0;
  const long long n_warp_first = (t_begin * PPT) / a.HW;
  long long n_base = n_warp_first;

  struct Fetched {
    long long nA, nB, n_first;
    int pixA, pixB;
    PixRaw rawA, rawB;
    float gA, gB;
  };
  auto fetch = [&](long long t, long long n_lane, int pix_lane, Fetched& f) {
    const int rows = tile_rows(t);
    const long long n_first = __shfl_sync(kFull, n_lane, 0);
    const int pix_first = __shfl_sync(kFull, pix_lane, 0);
    long long nB = n_lane;
    int pixB = pix_lane + 32;  // the lane's second row, 32 rows further on
    while (pixB >= a.HW) {
      pixB -= a.HW;
      ++nB;
    }
    const bool inA = lane < rows, inB = lane + 32 < rows;  // rows past a ragged last tile shadow the tile's first pixel
    f.nA = inA ? n_lane : n_first;
    f.pixA = inA ? pix_lane : pix_first;
    f.nB = inB ? nB : n_first;
    f.pixB = inB ? pixB : pix_first;
    f.n_first = n_first;
    f.rawA = load_pixel_raw(a, f.nA, f.pixA);
    f.rawB = load_pixel_raw(a, f.nB, f.pixB);
    f.gA = f.gB = 0.0f;
    if constexpr (BWD) {
      if (a.g_image) {
        f.gA = a.g_image[f.nA];
        f.gB = a.g_image[f.nB];
      }
      if (a.g_pixel) {
        f.gA += a.g_pixel[f.nA * a.HW + f.pixA];
        f.gB += a.g_pixel[f.nB * a.HW + f.pixB];
      }
    }
  };
  Fetched cur{};
  if (t_cnt > 0) fetch(t_first, n_own, pix_own, cur);

  for (long long it = 0; it < t_cnt; ++it) {
    const long long t = t_first + it * t_dir;
    const uint32_t parity = static_cast<uint32_t>(it & 1);
    const int rows = tile_rows(t);
    const bool actA = lane < rows, actB = lane + 32 < rows;
    const int ppA = actA ? lane : 0, ppB = actB ? lane + 32 : 0;
    const long long iA = t * PPT + ppA, iB = t * PPT + ppB;
    const long long nA = cur.nA, nB = cur.nB, n_first = cur.n_first;
    const f2 g2 = pk(cur.gA, cur.gB);
    PixelPair px;
    {
      Pixel pa, pb;
      decode_pixel(a, cur.rawA, pa);
      decode_pixel(a, cur.rawB, pb);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        px.x[c] = pk(pa.x[c], pb.x[c]);
        px.ll[c] = pa.left[c];
        px.lh[c] = pb.left[c];
        px.rl[c] = pa.right[c];
        px.rh[c] = pb.right[c];
      }
    }
    if (!rev) {
      n_own += step_n;
      pix_own += step_pix;
      if (pix_own >= a.HW) {
        pix_own -= a.HW;
        ++n_own;
      }
    } else {
      n_own -= step_n;
      pix_own -= step_pix;
      if (pix_own < 0) {
        pix_own += a.HW;
        --n_own;
      }
    }
    if (it + 1 < t_cnt) fetch(t + t_dir, n_own, pix_own, cur);

    float* rowA = slot + ppA * ROWF;
    float* rowB = slot + ppB * ROWF;
    float* auxA = aux + ppA * M;
    float* auxB = aux + ppB * M;
    mbar_wait(bar, parity);

    f2 lmax = pk(rowA[rot0], rowB[rot0]);
#pragma unroll
    for (int mi = 1; mi < M; ++mi) {
      const int m = (mi + rot0 >= M) ? mi + rot0 - M : mi + rot0;
      lmax = pk(fmaxf(lo(lmax), rowA[m]), fmaxf(hi(lmax), rowB[m]));
    }

    f2 sumW = sp(0.0f), sumWP = sp(0.0f);
#pragma unroll 1
    for (int mi = 0; mi < M; ++mi) {
      const int m = (mi + rot0 >= M) ? mi + rot0 - M : mi + rot0;
      const f2 lg = pk(rowA[m], rowB[m]);
      f2 mu[3], sc[3], kp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mu[c] = pk(rowA[(1 + 3 * c) * M + m], rowB[(1 + 3 * c) * M + m]);
        sc[c] = pk(rowA[(2 + 3 * c) * M + m], rowB[(2 + 3 * c) * M + m]);
        kp[c] = pk(rowA[(3 + 3 * c) * M + m], rowB[(3 + 3 * c) * M + m]);
      }
      const float smin = fminf(fminf(fminf(lo(sc[0]), hi(sc[0])), fminf(lo(sc[1]), hi(sc[1]))), fminf(lo(sc[2]), hi(sc[2])));
      const bool narrow = __any_sync(kFull, smin < kLsNarrow);
      const f2 W = ex2_2((lg - lmax) * kLog2e);
      f2 u[9];
      f2 P;
      if (narrow)
        P = pair_eval<true, BWD, PixelPair, AR>(px, mu, sc, kp, u);
      else
        P = pair_eval<false, BWD, PixelPair, AR>(px, mu, sc, kp, u);
      sumW = sumW + W;
      sumWP = fma2(W, P, sumWP);
      if constexpr (BWD) {
        // unscaled gradients overwrite the component's parameters in place; W*P goes to the aux strip (owners only)
        const f2 wp = W * P;
        if (actA) {
#pragma unroll
          for (int j = 0; j < 9; ++j) rowA[(1 + j) * M + m] = lo(u[j]);
          auxA[m] = lo(wp);
        }
        if (actB) {
#pragma unroll
          for (int j = 0; j < 9; ++j) rowB[(1 + j) * M + m] = hi(u[j]);
          auxB[m] = hi(wp);
        }
      }
    }
    const bool tinyA = !(lo(sumWP) > kTinySum), tinyB = !(hi(sumWP) > kTinySum);  // also catches NaN
    const float* growA = a.params + iA * ROWF;
    const float* growB = a.params + iB * ROWF;

    if constexpr (!BWD) {
      __syncwarp();
      if (it + 1 < t_cnt) issue(t + t_dir);  // every lane has read its rows: re-arm the slot with the warp's next tile
      float lpA = (lg2_split(lo(sumWP)) - lg2_split(lo(sumW))) * kLn2;  // utils/mdl.py:78-89 in one step
      float lpB = (lg2_split(hi(sumWP)) - lg2_split(hi(sumW))) * kLn2;
      if (tinyA) {
        float lt, ll;
        modl_pixel_logdomain(growA, M, half_pixel(px, false), a.plain != 0, lt, ll);
        lpA = lt - ll;
      }
      if (tinyB) {
        float lt, ll;
        modl_pixel_logdomain(growB, M, half_pixel(px, true), a.plain != 0, lt, ll);
        lpB = lt - ll;
      }
      if (a.lp_pixel) {
        if (actA) a.lp_pixel[iA] = lpA;
        if (actB) a.lp_pixel[iB] = lpB;
      }
      const float valA = actA ? lpA : 0.0f, valB = actB ? lpB : 0.0f;
      if (a.partial) {
        // a tile holds pixels of at most two images (HW >= 64 on this route): n_first and n_first + 1
        while (n_base < n_first) {
          const double done = warp_sum(acc0);
          if (lane == 0) a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = done;
          acc0 = acc1;
          acc1 = 0.0;
          ++n_base;
        }
        if (nA == n_base)
          acc0 += static_cast<double>(valA);
        else
          acc1 += static_cast<double>(valA);
        if (nB == n_base)
          acc0 += static_cast<double>(valB);
        else
          acc1 += static_cast<double>(valB);
      } else if (a.ll_atomic) {
        if (actA) atomicAdd(a.ll_atomic + nA, static_cast<double>(valA));
        if (actB) atomicAdd(a.ll_atomic + nB, static_cast<double>(valB));
      }
    } else {
      const f2 rS = rcp_2(sumWP), rSW = rcp_2(sumW);
      float ltA = 0.f, llA = 0.f, ltB = 0.f, llB = 0.f;
      Pixel pxa, pxb;
      if (tinyA || tinyB) {
        pxa = half_pixel(px, false);
        pxb = half_pixel(px, true);
        if (tinyA) modl_pixel_logdomain(growA, M, pxa, a.plain != 0, ltA, llA);
        if (tinyB) modl_pixel_logdomain(growB, M, pxb, a.plain != 0, ltB, llB);
      }
#pragma unroll 1
      for (int mi = 0; mi < M; ++mi) {
        const int m = (mi + rot0 >= M) ? mi + rot0 - M : mi + rot0;
        const f2 lg = pk(rowA[m], rowB[m]);
        const f2 W = ex2_2((lg - lmax) * kLog2e);
        const f2 wp = pk(auxA[m], auxB[m]);
        f2 r = wp * rS;     // posterior responsibility of the component
        f2 pi = W * rSW;    // softmax(logits)
        if (tinyA || tinyB) {
          float rA = lo(r), rB = hi(r), piA = lo(pi), piB = hi(pi);
          if (tinyA) {
            rA = expf(modl_logt(growA, M, m, pxa, a.plain != 0) - ltA);
            piA = expf(growA[m] - llA);
          }
          if (tinyB) {
            rB = expf(modl_logt(growB, M, m, pxb, a.plain != 0) - ltB);
            piB = expf(growB[m] - llB);
          }
          r = pk(rA, rB);
          pi = pk(piA, piB);
        }
        const f2 gr = r * g2;
        const f2 dl = (r - pi) * g2;
        if (actA) {
          rowA[m] = lo(dl);
#pragma unroll
          for (int j = 1; j < 10; ++j) rowA[j * M + m] *= lo(gr);
        }
        if (actB) {
          rowB[m] = hi(dl);
#pragma unroll
          for (int j = 1; j < 10; ++j) rowB[j * M + m] *= hi(gr);
        }
      }
      // hand the gradient tile to the TMA engine
      const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * 4u;
      float* dst = a.dparams + t * TILE_F;
      if ((bytes & 15u) == 0) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (a.bwd_hint & 2)
            bulk_s2g_hint(dst, slot, bytes, pol_first);
          else
            bulk_s2g(dst, slot, bytes);
          bulk_commit();
        }
      } else {
        __syncwarp();
        for (int q = lane; q < rows * ROWF; q += 32) dst[q] = slot[q];
        __syncwarp();
      }
      if (it + 1 < t_cnt) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        issue(t + t_dir);
      }
    }
  }
  if constexpr (BWD) {
    if (lane == 0) bulk_wait_all<0>();
  } else {
    if (a.partial && t_cnt > 0) {
      const long long n_last = (t_end * PPT < a.n_px ? t_end * PPT - 1 : a.n_px - 1) / a.HW;
      const double d0 = warp_sum(acc0), d1 = warp_sum(acc1);
      if (lane == 0) {
        a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
        if (n_base + 1 <= n_last) a.partial[partial_slot(n_base + 1, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
      }
    }
  }
}

// ---- the run-time tiled kernel (any n_mix) --------------------------------------------------------------------------------
// Same pipeline as modl_tile_kernel with one slot per warp, but n_mix and the split of a pixel over lanes are run-time
// values chosen by the host (rt_plan): LPP lanes share a pixel, lane `sub` of the group owns components
// [sub*MC, min(M, (sub+1)*MC)), two per packed register; PPT = 32 / LPP pixels per tile (one fewer when PPT * M would be
// odd: bulk copies need 16-byte multiples).  A lane walks its component pairs in an order rotated by rot * pixel so that
// the 32 scalar shared-memory loads of one instruction spread over the banks whatever 10 * M is modulo 32.
__device__ __forceinline__ float group_sum_rt(float v, int base, int LPP) {
  float s = __shfl_sync(kFull, v, base);
  for (int j = 1; j < LPP; ++j) s += __shfl_sync(kFull, v, (base + j) & 31);
  return s;
}
__device__ __forceinline__ float group_max_rt(float v, int base, int LPP) {
  float s = __shfl_sync(kFull, v, base);
  for (int j = 1; j < LPP; ++j) s = fmaxf(s, __shfl_sync(kFull, v, (base + j) & 31));
  return s;
}

// AL: n_mix and MC even -> every component pair sits on an 8-byte boundary and is moved with 64-bit shared accesses
template <bool BWD, int AR, bool AL, int PD = 0>
__global__ void __launch_bounds__(512, 1) modl_rt_kernel(const ModlArgs a) {
  const int M = a.M, MC = a.rt_MC, LPP = a.rt_LPP, PPT = a.rt_PPT;
  const int ROWF = 10 * M, TILE_F = PPT * ROWF, NPAIR = (MC + 1) >> 1;
  const int WARP_F = a.rt_warp_f;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* slot = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * WARP_F;
  float* aux = slot + TILE_F;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(nwarps) * WARP_F * 4) + warp;
  if (lane == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (!BWD && a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
  if constexpr (BWD)
    pdl_wait();
  else
    pdl_trigger();

  const long long gw = run_index(a, warp, nwarps);
  const int p_raw = lane / LPP;
  const bool lane_used = p_raw < PPT;
  const int p = lane_used ? p_raw : 0;  // idle lanes shadow pixel 0 (they never write)
  const int sub = lane - p_raw * LPP;
  const int gbase = lane - sub;         // first lane of this pixel's group
  const int m0 = sub * MC;
  const int m_end = m0 + MC < M ? m0 + MC : M;
  const int m_safe = m0 < M ? m0 : 0;   // a group's last lane may own nothing (M = 9 over 4 lanes)
  int rot = (a.rt_rot * p) % NPAIR;

  const long long t_begin = gw * a.tw_base + (gw < a.tw_rem ? gw : a.tw_rem);
  const long long t_end = t_begin + a.tw_base + (gw < a.tw_rem ? 1 : 0);
  const long long t_cnt = t_end - t_begin;
  const bool rev = BWD && a.reverse;
  const long long t_first = rev ? t_end - 1 : t_begin;
  const long long t_dir = rev ? -1 : 1;
  const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();

  auto tile_rows = [&](long long t) -> int {
    const long long rem = a.n_px - t * PPT;
    return rem < PPT ? static_cast<int>(rem) : PPT;
  };
  auto issue = [&](long long t) {
    const int rows = tile_rows(t);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * (PD ? 2u : 4u);
    const char* src = reinterpret_cast<const char*>(a.params) + t * TILE_F * (PD ? 2 : 4);
    char* land = reinterpret_cast<char*>(slot) + (PD ? bytes : 0u);  // bf16 lands behind the room its float32 image needs
    if ((bytes & 15u) == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, bytes);
        if (BWD) {
          if (a.bwd_hint & 1)
            bulk_g2s_hint(land, src, bytes, bar, pol_first);
          else
            bulk_g2s(land, src, bytes, bar);
        } else {
          if (a.keep_tiles > 0)
            bulk_g2s_hint(land, src, bytes, bar, (t_end - t) <= a.keep_tiles ? pol_last : pol_first);
          else
            bulk_g2s(land, src, bytes, bar);
        }
      }
    } else {
      for (int i = lane; i < rows * ROWF; i += 32)
        slot[i] = PD ? bf16_bits_to_f32(reinterpret_cast<const unsigned short*>(src)[i]) : reinterpret_cast<const float*>(src)[i];
      __syncwarp();
      if (lane == 0) mbar_arrive_expect_tx(bar, 0);
    }
  };
  if (t_cnt > 0) issue(t_first);

  const long long step_n = PPT / a.HW;
  const int step_pix = static_cast<int>(PPT - step_n * a.HW);
  long long n_own = (t_first * PPT + p) / a.HW;
  int pix_own = static_cast<int>((t_first * PPT + p) - n_own * a.HW);
  double acc0 = 0.0, acc1 = 0.0;
  long long n_base = (t_begin * PPT) / a.HW;

  auto fetch = [&](long long t, long long n_lane, int pix_lane, long long& n_out, long long& nfirst_out, PixRaw& raw,
                   float& g_out) {
    const int rows = tile_rows(t);
    const long long n_first = __shfl_sync(kFull, n_lane, 0);
    const int pix_first = __shfl_sync(kFull, pix_lane, 0);
    const bool in = p < rows;
    const long long n = in ? n_lane : n_first;
    const int pix = in ? pix_lane : pix_first;
    raw = load_pixel_raw(a, n, pix);
    g_out = 0.0f;
    if constexpr (BWD) {
      if (a.g_image) g_out = a.g_image[n];
      if (a.g_pixel) g_out += a.g_pixel[n * a.HW + pix];
    }
    n_out = n;
    nfirst_out = n_first;
  };
  long long n_cur = 0, nfirst_cur = 0;
  PixRaw raw_cur{};
  float g_cur = 0.0f;
  if (t_cnt > 0) fetch(t_first, n_own, pix_own, n_cur, nfirst_cur, raw_cur, g_cur);

  for (long long it = 0; it < t_cnt; ++it) {
    const long long t = t_first + it * t_dir;
    const uint32_t parity = static_cast<uint32_t>(it & 1);
    const int rows = tile_rows(t);
    const int pp = p < rows ? p : 0;
    const bool active = lane_used && (p < rows);
    const long long i = t * PPT + pp;
    const long long n = n_cur, n_first = nfirst_cur;
    const float g = g_cur;
    Pixel px;
    decode_pixel(a, raw_cur, px);
    if (!rev) {
      n_own += step_n;
      pix_own += step_pix;
      if (pix_own >= a.HW) {
        pix_own -= a.HW;
        ++n_own;
      }
    } else {
      n_own -= step_n;
      pix_own -= step_pix;
      if (pix_own < 0) {
        pix_own += a.HW;
        --n_own;
      }
    }
    if (it + 1 < t_cnt) fetch(t + t_dir, n_own, pix_own, n_cur, nfirst_cur, raw_cur, g_cur);

    float* rowp = slot + pp * ROWF;
    float* auxp = aux + pp * M;
    mbar_wait(bar, parity);
    if constexpr (PD != 0) {
      if (((rows * ROWF * 2) & 15) == 0) widen_bf16_inplace(slot, rows * ROWF, lane);
    }

    float lmax = -INFINITY;
    for (int m = m0; m < m_end; ++m) lmax = fmaxf(lmax, rowp[m]);
    lmax = group_max_rt(lmax, gbase, LPP);

    f2 sumW2 = sp(0.0f), sumWP2 = sp(0.0f);
#pragma unroll 1
    for (int pr = 0; pr < NPAIR; ++pr) {
      const int prr = pr + rot >= NPAIR ? pr + rot - NPAIR : pr + rot;
      const int m = m0 + 2 * prr;
      const bool vlo = m < m_end, vhi = m + 1 < m_end;
      const int ml = vlo ? m : m_safe, mh = vhi ? m + 1 : ml;
      f2 lg = ld_pair<AL>(rowp, ml, !vhi);
      lg = pk(vlo ? lo(lg) : -INFINITY, vhi ? hi(lg) : -INFINITY);  // padding halves get zero weight
      f2 mu[3], sc[3], kp[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* q = rowp + (1 + 3 * c) * M;
        mu[c] = ld_pair<AL>(q, ml, !vhi);
        sc[c] = ld_pair<AL>(q, M + ml, !vhi);
        kp[c] = ld_pair<AL>(q, 2 * M + ml, !vhi);
      }
      const float smin = fminf(fminf(fminf(lo(sc[0]), hi(sc[0])), fminf(lo(sc[1]), hi(sc[1]))), fminf(lo(sc[2]), hi(sc[2])));
      const bool narrow = __any_sync(kFull, smin < kLsNarrow);
      const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
      f2 u[9];
      f2 P;
      if (narrow)
        P = pair_eval<true, BWD, Pixel, AR>(px, mu, sc, kp, u);
      else
        P = pair_eval<false, BWD, Pixel, AR>(px, mu, sc, kp, u);
      if (!vlo) P = sp(0.0f);  // a padding pair was evaluated on clamped (backward: possibly overwritten) values
      sumW2 = sumW2 + W;
      sumWP2 = fma2(W, P, sumWP2);
      if constexpr (BWD) {
        const f2 wp = W * P;
        if (active && vlo) {  // (AL: a pair is valid or padding as a whole)
#pragma unroll
          for (int j = 0; j < 9; ++j) st_pair<AL>(rowp, (1 + j) * M + ml, !vhi, u[j]);
          st_pair<AL>(auxp, ml, !vhi, wp);
        }
      }
    }
    const float S = group_sum_rt(lo(sumWP2) + hi(sumWP2), gbase, LPP);
    const float SW = group_sum_rt(lo(sumW2) + hi(sumW2), gbase, LPP);
    const bool tiny = !(S > kTinySum);  // also catches NaN
    const float* grow = param_row(a, i, ROWF);

    if constexpr (!BWD) {
      if constexpr (PD != 0) fence_async_smem();
      __syncwarp();
      if (it + 1 < t_cnt) issue(t + t_dir);  // every lane has read its row: re-arm the slot with the warp's next tile
      float lp = (lg2_split(S) - lg2_split(SW)) * kLn2;
      if (tiny) {
        float lt, ll;
        modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, PD != 0);
        lp = lt - ll;
      }
      const bool owner = active && sub == 0;
      if (a.lp_pixel && owner) a.lp_pixel[i] = lp;
      const float val = owner ? lp : 0.0f;
      if (a.partial) {
        while (n_base < n_first) {
          const double done = warp_sum(acc0);
          if (lane == 0) a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = done;
          acc0 = acc1;
          acc1 = 0.0;
          ++n_base;
        }
        if (n == n_base)
          acc0 += static_cast<double>(val);
        else
          acc1 += static_cast<double>(val);
      } else if (a.ll_atomic) {
        if (owner) atomicAdd(a.ll_atomic + n, static_cast<double>(val));
      }
    } else {
      const float rS = rcpa(S), rSW = rcpa(SW);
      float lt = 0.f, ll = 0.f;
      if (tiny) modl_pixel_logdomain(grow, M, px, a.plain != 0, lt, ll, PD != 0);
#pragma unroll 1
      for (int pr = 0; pr < NPAIR; ++pr) {
        const int prr = pr + rot >= NPAIR ? pr + rot - NPAIR : pr + rot;
        const int m = m0 + 2 * prr;
        const bool vlo = m < m_end, vhi = m + 1 < m_end;
        const int ml = vlo ? m : m_safe, mh = vhi ? m + 1 : ml;
        const f2 lg = ld_pair<AL>(rowp, ml, !vhi);
        const f2 W = ex2_2((lg - sp(lmax)) * kLog2e);
        const f2 wp = ld_pair<AL>(auxp, ml, !vhi);
        f2 r = wp * rS;     // posterior responsibility of the component
        f2 pi = W * rSW;    // softmax(logits)
        if (tiny) {
          r = pk(expf(modl_logt(grow, M, ml, px, a.plain != 0, PD != 0) - lt), expf(modl_logt(grow, M, mh, px, a.plain != 0, PD != 0) - lt));
          pi = pk(expf(ld_param(grow, ml, PD != 0) - ll), expf(ld_param(grow, mh, PD != 0) - ll));
        }
        const f2 gr = r * g;
        const f2 dl = (r - pi) * g;
        if (active && vlo) {
          st_pair<AL>(rowp, ml, !vhi, dl);
#pragma unroll
          for (int j = 1; j < 10; ++j) st_pair<AL>(rowp, j * M + ml, !vhi, ld_pair<AL>(rowp, j * M + ml, !vhi) * gr);
        }
      }
      const uint32_t bytes = static_cast<uint32_t>(rows) * ROWF * (PD ? 2u : 4u);
      char* dst = reinterpret_cast<char*>(a.dparams) + t * TILE_F * (PD ? 2 : 4);
      if ((bytes & 15u) == 0) {
        if constexpr (PD != 0) {
          __syncwarp();
          narrow_bf16_inplace(slot, rows * ROWF, lane);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (a.bwd_hint & 2)
            bulk_s2g_hint(dst, slot, bytes, pol_first);
          else
            bulk_s2g(dst, slot, bytes);
          bulk_commit();
        }
      } else {
        __syncwarp();
        for (int q = lane; q < rows * ROWF; q += 32) {
          if (PD)
            reinterpret_cast<unsigned short*>(dst)[q] = f32_to_bf16_bits(slot[q]);
          else
            reinterpret_cast<float*>(dst)[q] = slot[q];
        }
        __syncwarp();
      }
      if (it + 1 < t_cnt) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        issue(t + t_dir);
      }
    }
  }
  if constexpr (BWD) {
    if (lane == 0) bulk_wait_all<0>();
  } else {
    if (a.partial && t_cnt > 0) {
      const long long n_last = (t_end * PPT < a.n_px ? t_end * PPT - 1 : a.n_px - 1) / a.HW;
      const double d0 = warp_sum(acc0), d1 = warp_sum(acc1);
      if (lane == 0) {
        a.partial[partial_slot(n_base, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d0;
        if (n_base + 1 <= n_last) a.partial[partial_slot(n_base + 1, gw, a.HW, PPT, a.tw_base, a.tw_rem, a.K, a.small)] = d1;
      }
    }
  }
}

// ---- any-M kernel: one thread per pixel-sample, parameters straight from global memory (correct, not tuned) --------------
template <bool BWD>
__global__ void __launch_bounds__(128) modl_generic_kernel(const ModlArgs a) {
  const int M = a.M;
  if (!BWD && a.zero_me && blockIdx.x == 0 && threadIdx.x == 0) *a.zero_me = 0u;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n_iter = (a.n_px + stride - 1) / stride;  // every lane runs the same trip count (warp votes inside)
  for (long long itn = 0; itn < n_iter; ++itn) {
    const long long i_raw = itn * stride + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool active = i_raw < a.n_px;
    const long long i = active ? i_raw : 0;
    Pixel px;
    const long long n = i / a.HW;
    load_pixel(a, n, static_cast<int>(i - n * a.HW), px);
    const float* row = a.params + i * 10 * M;
    float lmax = row[0];
    for (int m = 1; m < M; ++m) lmax = fmaxf(lmax, row[m]);
    float sumW = 0.f, sumWP = 0.f;
    for (int m = 0; m < M; ++m) {
      const float mu[3] = {row[M + m], row[4 * M + m], row[7 * M + m]};
      const float sc[3] = {row[2 * M + m], row[5 * M + m], row[8 * M + m]};
      const float kp[3] = {row[3 * M + m], row[6 * M + m], row[9 * M + m]};
      const float W = ex2a((row[m] - lmax) * kLog2e);
      sumW += W;
      sumWP = fmaf(W, mix_fwd(px, mu, sc, kp, a.plain != 0), sumWP);
    }
    const bool tiny = !(sumWP > kTinySum);
    float lt = 0.f, ll = 0.f;
    if (tiny) modl_pixel_logdomain(row, M, px, a.plain != 0, lt, ll);
    if constexpr (!BWD) {
      const float lp = tiny ? (lt - ll) : (lg2_split(sumWP) - lg2_split(sumW)) * kLn2;
      if (active) {
        if (a.lp_pixel) a.lp_pixel[i] = lp;
        if (a.ll_atomic) atomicAdd(a.ll_atomic + n, static_cast<double>(lp));
      }
    } else {
      float g = 0.f;
      if (a.g_image) g += a.g_image[n];
      if (a.g_pixel) g += a.g_pixel[i];
      const float rS = rcpa(sumWP), rSW = rcpa(sumW);
      float* orow = a.dparams + i * 10 * M;
      for (int m = 0; m < M; ++m) {
        const float mu[3] = {row[M + m], row[4 * M + m], row[7 * M + m]};
        const float sc[3] = {row[2 * M + m], row[5 * M + m], row[8 * M + m]};
        const float kp[3] = {row[3 * M + m], row[6 * M + m], row[9 * M + m]};
        const float W = ex2a((row[m] - lmax) * kLog2e);
        float u[9];
        const float P = mix_bwd(px, mu, sc, kp, u, a.plain != 0);
        float r = W * P * rS, pi = W * rSW;
        if (tiny) {
          r = expf(modl_logt(row, M, m, px, a.plain != 0) - lt);
          pi = expf(row[m] - ll);
        }
        if (active) {
          orow[m] = g * (r - pi);
          const float gr = g * r;
#pragma unroll
          for (int j = 0; j < 9; ++j) orow[(1 + j) * M + m] = gr * u[j];
        }
      }
    }
  }
}

// ---- launchers -------------------------------------------------------------------------------------------------------------
// Two shapes per kernel: (2 slots, <=8 warps, 255 registers) and (1 slot, <=16 warps, 128 registers).  VAEMDL_TUNE
// ("fwd=S:W,bwd=S:W", S slots, W warps per CTA) overrides the built-in choice; used by the tuning sweeps under tools/.
struct Shape {
  int slots, warps;
};
static Shape tune_shape(bool bwd, Shape dflt) {
  const char* env = getenv("VAEMDL_TUNE");
  if (!env) return dflt;
  const char* key = bwd ? "bwd=" : "fwd=";
  const char* p = strstr(env, key);
  if (!p) return dflt;
  int s = 0, w = 0;
  if (sscanf(p + 4, "%d:%d", &s, &w) == 2 && (s == 1 || s == 2) && w >= 1 && w <= 16) return Shape{s, w};
  return dflt;
}

// Small problems: with only a few tiles per warp the rounding of tiles/warp up to an integer costs more than a little
// occupancy does, so pick the warp count (>= 10) whose runs come out most even.  Large problems keep `max_warps`.
static int pick_warps(long long num_tiles, int sm_count, int max_warps) {
  if (getenv("VAEMDL_TUNE")) return max_warps;
  if (num_tiles >= static_cast<long long>(sm_count) * max_warps * 8) return max_warps;
  int best = max_warps;
  double best_score = -1.0;
  for (int w = max_warps; w >= 10 && w >= max_warps - 6; --w) {
    const double per = static_cast<double>(num_tiles) / (static_cast<double>(sm_count) * w);
    if (per <= 1.0) break;  // fewer tiles than warps: the grid shrinks instead
    const double longest = static_cast<double>((num_tiles + static_cast<long long>(sm_count) * w - 1) / (static_cast<long long>(sm_count) * w));
    const double score = per / longest * (0.8 + 0.2 * w / max_warps);
    if (score > best_score) {
      best_score = score;
      best = w;
    }
  }
  return best;
}

struct L2Opt {  // VAEMDL_L2="rev=0|1,keep=<MB>,hint=0|1": L2 reuse between the forward and the backward kernel of a step
  int rev = 1, keep_mb = 48, hint = 0;  // measured on B200: profiles/r01_l2_reuse.txt
  L2Opt() {
    const char* e = getenv("VAEMDL_L2");
    if (!e) return;
    const char* q;
    if ((q = strstr(e, "rev="))) rev = atoi(q + 4);
    if ((q = strstr(e, "keep="))) keep_mb = atoi(q + 5);
    if ((q = strstr(e, "hint="))) hint = atoi(q + 5);
  }
};
static void apply_l2_opt(ModlArgs& a, long long total_warps, long long tile_bytes) {
  static const L2Opt opt;
  a.reverse = opt.rev;
  a.bwd_hint = opt.hint;
  a.keep_tiles = static_cast<int>((static_cast<long long>(opt.keep_mb) << 20) / (total_warps * tile_bytes));
  if (opt.keep_mb > 0 && a.keep_tiles < 1) a.keep_tiles = 1;
}

struct TilePlan {  // how the forward grid split the tile range: what the per-image reduction needs to know
  long long total_warps = 0, tw_base = 0, tw_rem = 0;
  int K = 0, PPT = 0;
};

template <int MC, int LPP, bool BWD, int NSLOT, int MAXT, int AR, int PD = 0>
static int launch_tiled_shape(ModlArgs a, int warps, cudaStream_t st, TilePlan* plan) {
  using T = Tile<MC, LPP>;
  a.num_tiles = (a.n_px + T::PPT - 1) / T::PPT;
  const DeviceInfo& di = device_info();
  const size_t per_warp = (static_cast<size_t>(NSLOT) * T::TILE_F + (BWD ? T::AUX_F : 0)) * 4 + NSLOT * 8;
  if (warps > MAXT / 32) warps = MAXT / 32;
  while (warps > 1 && warps * per_warp > static_cast<size_t>(di.max_smem_optin)) --warps;
  warps = pick_warps(a.num_tiles, di.sm_count, warps);
  const size_t smem = warps * per_warp;
  if (smem > static_cast<size_t>(di.max_smem_optin)) return VAEMDL_EUNSUPPORTED;
  auto kern = modl_tile_kernel<MC, LPP, BWD, NSLOT, MAXT, AR, PD>;
  // the function attribute and the occupancy query cost several microseconds of host time: once per (device, shape)
  static std::mutex mu;
  static int c_dev = -1, c_warps = -1, c_ctas = 1;
  int ctas_per_sm;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || c_warps != warps) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      int n = 1;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, warps * 32, smem);
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_warps = warps;
      c_ctas = n < 1 ? 1 : n;
    }
    ctas_per_sm = c_ctas;
  }
  const long long need = (a.num_tiles + warps - 1) / warps;
  long long grid = static_cast<long long>(di.sm_count) * ctas_per_sm;  // persistent: every CTA resident
  if (grid > need) grid = need;
  if (grid * warps > kMaxGridWarps) grid = kMaxGridWarps / warps;
  if (grid < 1) grid = 1;
  const long long total_warps = grid * warps;
  a.tw_base = a.num_tiles / total_warps;
  a.tw_rem = a.num_tiles % total_warps;
  a.K = partial_K(a.HW, T::PPT, a.tw_base);
  apply_l2_opt(a, total_warps, T::TILE_B / (PD ? 2 : 1));
  if (plan) {
    plan->total_warps = total_warps;
    plan->tw_base = a.tw_base;
    plan->tw_rem = a.tw_rem;
    plan->K = a.K;
    plan->PPT = T::PPT;
  }
  if (BWD) return cuda_rc(launch_pdl(kern, static_cast<unsigned>(grid), static_cast<unsigned>(warps * 32), smem, st, a));
  kern<<<static_cast<unsigned>(grid), warps * 32, smem, st>>>(a);
  return cuda_rc(cudaGetLastError());
}

template <int MC, int LPP, bool BWD, int AR>
static int launch_tiled(ModlArgs a, cudaStream_t st, TilePlan* plan) {
  if (a.bf16) {  // bfloat16 parameters: x-conditioned class only, one slot per warp
    if constexpr (AR == 0)
      return launch_tiled_shape<MC, LPP, BWD, 1, 512, 0, 1>(a, tune_shape(BWD, Shape{1, 16}).warps, st, plan);
    else
      return VAEMDL_EUNSUPPORTED;
  }
  // 1 slot x 16 warps: measured best on B200 for every M (profiles/r01_tune_shapes.txt); latency is hidden by the 4
  // warps per scheduler rather than by a second slot per warp
  const Shape sh = tune_shape(BWD, Shape{1, 16});
  if (AR == 0 && sh.slots == 2) return launch_tiled_shape<MC, LPP, BWD, 2, 256, 0>(a, sh.warps, st, plan);  // tuning only
  return launch_tiled_shape<MC, LPP, BWD, 1, 512, AR>(a, sh.warps, st, plan);
}

// the fused one-launch step (modl_step_kernel): backward shared-memory footprint, cooperative launch
template <int MC, int LPP, int AR>
static int launch_step(ModlArgs a, StepFinish f, long long n_img, cudaStream_t st) {
  using T = Tile<MC, LPP>;
  a.num_tiles = (a.n_px + T::PPT - 1) / T::PPT;
  const DeviceInfo& di = device_info();
  const size_t per_warp = (static_cast<size_t>(T::TILE_F) + T::AUX_F) * 4 + 8;
  int warps = 16;
  while (warps > 1 && warps * per_warp > static_cast<size_t>(di.max_smem_optin)) --warps;
  warps = pick_warps(a.num_tiles, di.sm_count, warps);
  const size_t smem = warps * per_warp;
  auto kern = modl_step_kernel<MC, LPP, AR>;
  static std::mutex mu;
  static int c_dev = -1;
  static size_t c_smem = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || smem > c_smem) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_smem = smem;
    }
  }
  const long long need = (a.num_tiles + warps - 1) / warps;
  long long grid = di.sm_count;  // one CTA per SM: every CTA is resident, as the grid barrier requires
  if (grid > need) grid = need;
  if (grid * warps > kMaxGridWarps) grid = kMaxGridWarps / warps;
  if (grid < 1) grid = 1;
  const long long total_warps = grid * warps;
  a.tw_base = a.num_tiles / total_warps;
  a.tw_rem = a.num_tiles % total_warps;
  a.K = partial_K(a.HW, T::PPT, a.tw_base);
  if (static_cast<size_t>(n_img) * a.K > partial_elems(n_img)) return VAEMDL_EWORKSPACE;
  a.reverse = 1;
  a.keep_tiles = 0;
  a.bwd_hint = 0;
  f.geom = PartialGeom{a.partial, a.tw_base, a.tw_rem, a.K, T::PPT, a.HW};
  StepArgs sa{a, f};
  void* args[] = {&sa};
  return cuda_rc(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(static_cast<unsigned>(grid)),
                                             dim3(static_cast<unsigned>(warps * 32)), args, smem, st));
}

// pixel-pair kernel, n_mix = M (1 <= M <= 9): one slot per warp, as many warps as shared memory allows (<= 16)
template <int M, bool BWD, int AR>
static int launch_pp(ModlArgs a, cudaStream_t st, TilePlan* plan) {
  using T = TilePP<M>;
  a.num_tiles = (a.n_px + T::PPT - 1) / T::PPT;
  const DeviceInfo& di = device_info();
  const size_t per_warp = (static_cast<size_t>(T::TILE_F) + (BWD ? T::AUX_F : 0)) * 4 + 8;
  int warps = tune_shape(BWD, Shape{1, 16}).warps;
  while (warps > 1 && warps * per_warp > static_cast<size_t>(di.max_smem_optin)) --warps;
  warps = pick_warps(a.num_tiles, di.sm_count, warps);
  const size_t smem = warps * per_warp;
  auto kern = modl_pp_kernel<M, BWD, 512, AR>;
  static std::mutex mu;
  static int c_dev = -1, c_warps = -1, c_ctas = 1;
  int ctas_per_sm;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || c_warps != warps) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      int n = 1;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, warps * 32, smem);
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_warps = warps;
      c_ctas = n < 1 ? 1 : n;
    }
    ctas_per_sm = c_ctas;
  }
  const long long need = (a.num_tiles + warps - 1) / warps;
  long long grid = static_cast<long long>(di.sm_count) * ctas_per_sm;
  if (grid > need) grid = need;
  if (grid * warps > kMaxGridWarps) grid = kMaxGridWarps / warps;
  if (grid < 1) grid = 1;
  const long long total_warps = grid * warps;
  a.tw_base = a.num_tiles / total_warps;
  a.tw_rem = a.num_tiles % total_warps;
  a.K = partial_K(a.HW, T::PPT, a.tw_base);
  a.small = a.n_px < (1ll << 31) - 64;
  apply_l2_opt(a, total_warps, T::TILE_B);
  if (plan) {
    plan->total_warps = total_warps;
    plan->tw_base = a.tw_base;
    plan->tw_rem = a.tw_rem;
    plan->K = a.K;
    plan->PPT = T::PPT;
  }
  if (BWD) return cuda_rc(launch_pdl(kern, static_cast<unsigned>(grid), static_cast<unsigned>(warps * 32), smem, st, a));
  kern<<<static_cast<unsigned>(grid), warps * 32, smem, st>>>(a);
  return cuda_rc(cudaGetLastError());
}

// ---- run-time tile geometry for modl_rt_kernel -------------------------------------------------------------------------
// For a given n_mix pick (LPP, rot): score = lane efficiency / sqrt(mean bank-conflict degree of the scalar parameter loads)
// * sqrt(min(1, warps that fit / 16)).  Evaluated once per n_mix (a few thousand integer operations) and cached.
struct RtPlan {
  int MC = 0, LPP = 0, PPT = 0, rot = 0;
};
// score = lane efficiency * MC / (MC + 3) * sqrt(min(1, warps that fit / 16)) / (mean bank-conflict degree)^(1/4):
// measured on B200 (tools/rt_sweep.py), long component chunks per lane win over many lanes per pixel (the per-tile work of
// a lane -- logit max, group reductions, log, index bookkeeping -- is amortised over MC components), as long as 16 warps
// still fit in shared memory.  Evaluated once per (n_mix, direction) and cached.
static RtPlan rt_plan_compute(int M, bool bwd, bool bf16) {
  RtPlan best;
  double best_score = -1.0;
  for (int LPP = 1; LPP <= 16; ++LPP) {
    const int MC = (M + LPP - 1) / LPP;
    if (MC > 13) continue;
    int PPT = 32 / LPP;
    // tile bytes = PPT * 40 * M (float32) or PPT * 20 * M (bfloat16) must be a multiple of 16
    while (PPT >= 1 && ((PPT * M) & (bf16 ? 3 : 1))) --PPT;
    if (PPT < 1) continue;
    const int NP = (MC + 1) / 2;
    const double eff = static_cast<double>(M) / (2.0 * NP * LPP) * (static_cast<double>(PPT) * LPP / 32.0);
    const double per_warp = (PPT * 10.0 * M + (bwd ? PPT * M : 0)) * 4.0 + 24.0;
    double occ = 227.0 * 1024.0 / per_warp / 16.0;
    if (occ > 1.0) occ = 1.0;
    for (int rot = 0; rot < (NP > 1 ? 4 : 1); ++rot) {
      long long tot = 0, cnt = 0;
      for (int pr = 0; pr < NP; ++pr) {
        for (int half = 0; half < 2; ++half) {
          // the ten parameter planes of a row are k*M apart: evaluate the first one
          int per_bank_addr[32][32];
          int per_bank_n[32] = {0};
          for (int lane = 0; lane < 32; ++lane) {
            const int pq = lane / LPP, sub = lane % LPP;
            if (pq >= PPT) continue;
            const int m0 = sub * MC, m_end = m0 + MC < M ? m0 + MC : M;
            const int prr = (pr + rot * pq) % NP;
            int m = m0 + 2 * prr + half;
            if (m >= m_end) m = m0 < M ? m0 : 0;
            const int addr = pq * 10 * M + m;
            const int bnk = addr & 31;
            bool seen = false;
            for (int q = 0; q < per_bank_n[bnk]; ++q) seen = seen || per_bank_addr[bnk][q] == addr;
            if (!seen) per_bank_addr[bnk][per_bank_n[bnk]++] = addr;
          }
          int deg = 1;
          for (int bnk = 0; bnk < 32; ++bnk) deg = per_bank_n[bnk] > deg ? per_bank_n[bnk] : deg;
          tot += deg;
          ++cnt;
        }
      }
      const double conflict = static_cast<double>(tot) / static_cast<double>(cnt);
      const double score = eff * MC / (MC + 3.0) * sqrt(occ) / sqrt(sqrt(conflict));
      if (score > best_score) {
        best_score = score;
        best = RtPlan{MC, LPP, PPT, rot};
      }
    }
  }
  return best;
}
static RtPlan rt_plan(int M, bool bwd, bool bf16 = false) {
  const char* env = getenv("VAEMDL_RT");  // "LPP:rot" overrides the choice (tuning sweeps; re-read on every call)
  int l = 0, r = 0;
  if (env && sscanf(env, "%d:%d", &l, &r) == 2 && l >= 1 && l <= 32 && (M + l - 1) / l <= 16) {
    int ppt = 32 / l;
    while (ppt >= 1 && ((ppt * M) & (bf16 ? 3 : 1))) --ppt;
    if (ppt >= 1) return RtPlan{(M + l - 1) / l, l, ppt, r};
  }
  static std::mutex mu;
  static RtPlan cache[4][VAEMDL_MAX_MIX + 1];
  std::lock_guard<std::mutex> lock(mu);
  RtPlan& c = cache[(bwd ? 1 : 0) + (bf16 ? 2 : 0)][M];
  if (c.LPP == 0) c = rt_plan_compute(M, bwd, bf16);
  return c;
}

template <bool BWD, int AR, bool AL, int PD>
static int launch_rt_al(ModlArgs a, const RtPlan& rp, cudaStream_t st, TilePlan* plan);

template <bool BWD, int AR>
static int launch_rt(ModlArgs a, cudaStream_t st, TilePlan* plan) {
  const RtPlan rp = rt_plan(a.M, BWD, a.bf16 != 0);
  if (rp.LPP == 0) return VAEMDL_EUNSUPPORTED;
  static const bool no_al = getenv("VAEMDL_RT_NOAL") != nullptr;  // A/B: scalar shared accesses everywhere
  const bool al = (a.M % 2 == 0) && (rp.MC % 2 == 0 || rp.LPP == 1) && !no_al;
  if (a.bf16) {
    if constexpr (AR == 0)
      return al ? launch_rt_al<BWD, 0, true, 1>(a, rp, st, plan) : launch_rt_al<BWD, 0, false, 1>(a, rp, st, plan);
    else
      return VAEMDL_EUNSUPPORTED;
  }
  return al ? launch_rt_al<BWD, AR, true, 0>(a, rp, st, plan) : launch_rt_al<BWD, AR, false, 0>(a, rp, st, plan);
}

template <bool BWD, int AR, bool AL, int PD>
static int launch_rt_al(ModlArgs a, const RtPlan& rp, cudaStream_t st, TilePlan* plan) {
  a.rt_MC = rp.MC;
  a.rt_LPP = rp.LPP;
  a.rt_PPT = rp.PPT;
  a.rt_rot = rp.rot;
  const int tile_f = rp.PPT * 10 * a.M;
  a.rt_warp_f = (tile_f + (BWD ? rp.PPT * a.M : 0) + 3) & ~3;  // every warp's slot stays 16-byte aligned
  a.num_tiles = (a.n_px + rp.PPT - 1) / rp.PPT;
  const DeviceInfo& di = device_info();
  const size_t per_warp = static_cast<size_t>(a.rt_warp_f) * 4 + 8;
  int warps = tune_shape(BWD, Shape{1, 16}).warps;
  while (warps > 1 && warps * per_warp > static_cast<size_t>(di.max_smem_optin)) --warps;
  if (warps * per_warp > static_cast<size_t>(di.max_smem_optin)) return VAEMDL_EUNSUPPORTED;
  warps = pick_warps(a.num_tiles, di.sm_count, warps);
  const size_t smem = warps * per_warp;
  auto kern = modl_rt_kernel<BWD, AR, AL, PD>;
  static std::mutex mu;
  static int c_dev = -1;
  static size_t c_smem = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (c_dev != dev || smem > c_smem) {  // the attribute is a maximum: raise it when a larger footprint shows up
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_rc(e);
      c_dev = dev;
      c_smem = smem;
    }
  }
  const long long need = (a.num_tiles + warps - 1) / warps;
  long long grid = di.sm_count;  // persistent, one CTA per SM (__launch_bounds__(512, 1))
  if (grid > need) grid = need;
  if (grid * warps > kMaxGridWarps) grid = kMaxGridWarps / warps;
  if (grid < 1) grid = 1;
  const long long total_warps = grid * warps;
  a.tw_base = a.num_tiles / total_warps;
  a.tw_rem = a.num_tiles % total_warps;
  a.K = partial_K(a.HW, rp.PPT, a.tw_base);
  a.small = a.n_px < (1ll << 31) - 64;
  apply_l2_opt(a, total_warps, static_cast<long long>(tile_f) * (PD ? 2 : 4));
  if (plan) {
    plan->total_warps = total_warps;
    plan->tw_base = a.tw_base;
    plan->tw_rem = a.tw_rem;
    plan->K = a.K;
    plan->PPT = rp.PPT;
  }
  if (BWD) return cuda_rc(launch_pdl(kern, static_cast<unsigned>(grid), static_cast<unsigned>(warps * 32), smem, st, a));
  kern<<<static_cast<unsigned>(grid), warps * 32, smem, st>>>(a);
  return cuda_rc(cudaGetLastError());
}

// n_mix 1..9 run on the pixel-pair kernel.  n_mix = 5 also has a component-pair instantiation with 32-row tiles, which
// is a little faster while the problem is so small that a warp only sees a handful of tiles (measured: 112 vs 117 us
// per step at 5 x 128 x 32 x 32, 345 vs 314 us backward at 16 x 64 x 64 x 64).
static bool use_pixel_pairs(int M, long long n_px, bool bf16 = false) {
  const char* env = getenv("VAEMDL_PP");  // "0" / "1" force the choice for n_mix = 5 (A/B measurements, tests)
  if (M < 1 || M > 9 || bf16) return false;  // (bfloat16 parameters: tile<5,1> for n_mix = 5, the run-time kernel otherwise)
  if (M != 5) return true;
  if (env && (env[0] == '0' || env[0] == '1')) return env[0] == '1';
  return n_px >= 64ll * 148 * 16 * 6;
}

static int spread_runs() {
  const char* e = getenv("VAEMDL_SPREAD");  // "0": CTA-major run numbering (A/B)
  return !(e && e[0] == '0');
}

template <bool BWD, int AR>
static int launch_modl(ModlArgs a, cudaStream_t st, TilePlan* plan = nullptr) {
  a.plain = AR;
  a.spread = spread_runs();
  if (use_pixel_pairs(a.M, a.n_px, a.bf16 != 0)) {
    switch (a.M) {
      case 1: return launch_pp<1, BWD, AR>(a, st, plan);
      case 2: return launch_pp<2, BWD, AR>(a, st, plan);
      case 3: return launch_pp<3, BWD, AR>(a, st, plan);
      case 4: return launch_pp<4, BWD, AR>(a, st, plan);
      case 5: return launch_pp<5, BWD, AR>(a, st, plan);
      case 6: return launch_pp<6, BWD, AR>(a, st, plan);
      case 7: return launch_pp<7, BWD, AR>(a, st, plan);
      case 8: return launch_pp<8, BWD, AR>(a, st, plan);
      case 9: return launch_pp<9, BWD, AR>(a, st, plan);
    }
  }
  switch (a.M) {
    case 5:
      return launch_tiled<5, 1, BWD, AR>(a, st, plan);
    case 10:
      return launch_tiled<10, 1, BWD, AR>(a, st, plan);
    case 20:
      return launch_tiled<10, 2, BWD, AR>(a, st, plan);
    case 30:
      return launch_tiled<10, 3, BWD, AR>(a, st, plan);
    default: {
      static const bool force_generic = getenv("VAEMDL_GENERIC") != nullptr;  // A/B against the one-thread-per-pixel kernel
      if (!force_generic) return launch_rt<BWD, AR>(a, st, plan);
      if (a.bf16) return VAEMDL_EUNSUPPORTED;
      const DeviceInfo& di = device_info();
      long long blocks = (a.n_px + 127) / 128;
      const long long cap = static_cast<long long>(di.sm_count) * 8;
      if (blocks > cap) blocks = cap;
      modl_generic_kernel<BWD><<<static_cast<unsigned>(blocks), 128, 0, st>>>(a);
      return cuda_rc(cudaGetLastError());
    }
  }
}

static int tile_ppt(int M, long long n_px, bool bf16 = false) {
  if (use_pixel_pairs(M, n_px, bf16)) return 64;
  switch (M) {
    case 5:
    case 10:
      return 32;
    case 20:
      return 16;
    case 30:
      return 10;
    default:
      if (getenv("VAEMDL_GENERIC")) return 0;  // one-thread-per-pixel kernel: atomics
      return rt_plan(M, false, bf16).PPT;
  }
}

static int check_common(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, long long n_img,
                        int x_batch, int H, int W, int M) {
  if (!params || !x) return VAEMDL_EINVAL;
  if (n_img <= 0 || x_batch <= 0 || H <= 0 || W <= 0) return VAEMDL_EINVAL;
  if (x_dtype != VAEMDL_X_F32 && x_dtype != VAEMDL_X_U8) return VAEMDL_EINVAL;
  if (x_range != VAEMDL_RANGE_UNIT && x_range != VAEMDL_RANGE_SYM) return VAEMDL_EINVAL;
  if (x_dtype == VAEMDL_X_U8 && x_range != VAEMDL_RANGE_UNIT) return VAEMDL_EINVAL;
  if (edge_mode != VAEMDL_EDGE_MDL && edge_mode != VAEMDL_EDGE_OPENAI) return VAEMDL_EINVAL;
  if (M < 1 || M > VAEMDL_MAX_MIX) return VAEMDL_EUNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(params) & 15u) return VAEMDL_EALIGN;
  return VAEMDL_OK;
}

}  // namespace vaemdl


namespace vaemdl {
template <int AR>
static int modl_fwd_impl(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, long long n_img,
                         int x_batch, int H, int W, int M, float* lp_pixel, float* ll_image, double* ll_image_f64,
                         const IwaeOut& iw, void* workspace, size_t workspace_bytes, cudaStream_t st, int bf16 = 0) {
  int rc = check_common(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M);
  if (rc) return rc;
  const bool iwae = iw.S > 0;
  if (!lp_pixel && !ll_image && !ll_image_f64 && !iwae) return VAEMDL_EINVAL;
  if (iwae && static_cast<long long>(iw.S) * iw.B != n_img) return VAEMDL_EINVAL;
  const bool want_ll = ll_image || ll_image_f64 || iwae;
  ModlArgs a{};
  a.params = params;
  a.x = x;
  a.lp_pixel = lp_pixel;
  a.HW = H * W;
  a.n_px = n_img * a.HW;
  a.x_batch = x_batch;
  a.x_u8 = x_dtype == VAEMDL_X_U8;
  a.x_unit = x_range == VAEMDL_RANGE_UNIT;
  a.edge_openai = edge_mode == VAEMDL_EDGE_OPENAI;
  a.M = M;
  a.bf16 = bf16;
  const int ppt = tile_ppt(M, a.n_px, bf16 != 0);
  const bool use_partials = want_ll && ppt > 0 && a.HW >= ppt;
  char* ws = static_cast<char*>(workspace);
  size_t tail_off = 0;
  if (want_ll) {
    const size_t need = vaemdl_modl_workspace_bytes(n_img, H, W);
    if (!workspace || workspace_bytes < need) return VAEMDL_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) & 7u) return VAEMDL_EALIGN;
    tail_off = partial_elems(n_img) * sizeof(double);
    if (use_partials) {
      a.partial = reinterpret_cast<double*>(ws);
    } else {
      a.ll_atomic = ll_image_f64 ? ll_image_f64 : reinterpret_cast<double*>(ws);
      cudaError_t e = cudaMemsetAsync(a.ll_atomic, 0, sizeof(double) * n_img, st);
      if (e != cudaSuccess) return cuda_rc(e);
    }
  }
  unsigned* counter = want_ll ? reinterpret_cast<unsigned*>(ws + tail_off + static_cast<size_t>(n_img) * sizeof(double)) : nullptr;
  if (iwae && use_partials && iw.elbo) a.zero_me = counter;
  TilePlan plan;
  rc = launch_modl<false, AR>(a, st, &plan);
  if (rc) return rc;
  if (use_partials) {
    if (static_cast<size_t>(n_img) * plan.K > partial_elems(n_img)) return VAEMDL_EWORKSPACE;
    const PartialGeom geom{a.partial, plan.tw_base, plan.tw_rem, plan.K, plan.PPT, a.HW};
    return finish_partials(geom, n_img, ll_image, ll_image_f64, iw, reinterpret_cast<double*>(ws + tail_off), counter, st);
  }
  if (!want_ll) return VAEMDL_OK;
  // float64 atomics route (any-M kernel, images smaller than a tile)
  if (ll_image) {
    cast_f64_f32_kernel<<<static_cast<unsigned>((n_img + 255) / 256), 256, 0, st>>>(a.ll_atomic, ll_image, n_img);
    rc = cuda_rc(cudaGetLastError());
  }
  if (rc || !iwae) return rc;
  return vaemdl_iwae_tail(nullptr, a.ll_atomic, iw.extra, iw.S, iw.B, iw.B_total, iw.log_w, iw.lme_b, iw.elbo, iw.g_ll, st);
}
}  // namespace vaemdl

namespace vaemdl {
template <int AR>
static int modl_bwd_impl(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, long long n_img,
                         int x_batch, int H, int W, int M, const float* g_image, const float* g_pixel, float* dparams,
                         cudaStream_t st, int bf16 = 0) {
  int rc = check_common(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M);
  if (rc) return rc;
  if (!dparams || (!g_image && !g_pixel)) return VAEMDL_EINVAL;
  if (reinterpret_cast<uintptr_t>(dparams) & 15u) return VAEMDL_EALIGN;
  ModlArgs a{};
  a.params = params;
  a.x = x;
  a.g_image = g_image;
  a.g_pixel = g_pixel;
  a.dparams = dparams;
  a.HW = H * W;
  a.n_px = n_img * a.HW;
  a.x_batch = x_batch;
  a.x_u8 = x_dtype == VAEMDL_X_U8;
  a.x_unit = x_range == VAEMDL_RANGE_UNIT;
  a.edge_openai = edge_mode == VAEMDL_EDGE_OPENAI;
  a.M = M;
  a.bf16 = bf16;
  return launch_modl<true, AR>(a, st);
}

template <int AR>
static int modl_iwae_fwd_impl(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, int S,
                              long long B, long long B_total, int x_batch, int H, int W, int M, const float* extra,
                              float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                              void* workspace, size_t workspace_bytes, cudaStream_t st, int bf16 = 0) {
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
  return modl_fwd_impl<AR>(params, x, x_dtype, x_range, edge_mode, static_cast<long long>(S) * B, x_batch, H, W, M, nullptr,
                           ll_image, ll_image_f64, iw, workspace, workspace_bytes, st, bf16);
}
}  // namespace vaemdl

namespace vaemdl {
// VAEMDL_FUSED = "0": never, "1": whenever the shape is eligible, unset: eligible shapes with at most kFusedMaxTilesPerWarp
// tiles per warp (where launch boundaries and pipeline ramps are a visible share of the step)
constexpr long long kFusedMaxTilesPerWarp = 12;  // measured: 101.6 -> 96.6 us at 4.3 tiles per warp, a loss from ~25 on
static int fused_mode() {
  const char* e = getenv("VAEMDL_FUSED");
  if (!e) return -1;
  return e[0] == '0' ? 0 : 1;
}
static bool fused_eligible(int S, long long n_px, int HW, int M) {
  const int mode = fused_mode();
  if (mode == 0 || S > 32) return false;
  if (use_pixel_pairs(M, n_px)) return false;
  if (M != 5 && M != 10 && M != 20 && M != 30) return false;
  const int ppt = tile_ppt(M, n_px);
  if (HW < ppt) return false;
  static int coop = -1;
  if (coop < 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev);
    coop = v;
  }
  if (!coop) return false;
  if (mode == 1) return true;
  const long long tiles = (n_px + ppt - 1) / ppt;
  return tiles <= kFusedMaxTilesPerWarp * device_info().sm_count * 16;
}

// One IWAE step of the observation model: forward, per-image sums, log-mean-exp, elbo, softmax weights, parameter gradient.
// One cooperative launch when the shape is eligible, else forward + finish + backward (3 launches).
template <int AR>
static int modl_iwae_step_impl(const float* params, const void* x, int x_dtype, int x_range, int edge_mode, int S,
                               long long B, long long B_total, int x_batch, int H, int W, int M, const float* extra,
                               float* ll_image, double* ll_image_f64, float* log_w, float* lme_b, float* elbo, float* g_ll,
                               float* dparams, void* workspace, size_t workspace_bytes, cudaStream_t st, int* launches) {
  if (S <= 0 || B <= 0 || B_total < 0 || !lme_b || !g_ll || !dparams) return VAEMDL_EINVAL;
  const long long n_img = static_cast<long long>(S) * B;
  int rc = check_common(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M);
  if (rc) return rc;
  if (reinterpret_cast<uintptr_t>(dparams) & 15u) return VAEMDL_EALIGN;
  const int HW = H * W;
  const long long n_px = n_img * HW;
  if (!fused_eligible(S, n_px, HW, M)) {
    if (launches) *launches = 3;
    rc = modl_iwae_fwd_impl<AR>(params, x, x_dtype, x_range, edge_mode, S, B, B_total, x_batch, H, W, M, extra, ll_image,
                                ll_image_f64, log_w, lme_b, elbo, g_ll, workspace, workspace_bytes, st);
    if (rc) return rc;
    return modl_bwd_impl<AR>(params, x, x_dtype, x_range, edge_mode, n_img, x_batch, H, W, M, g_ll, nullptr, dparams, st);
  }
  const size_t need = vaemdl_modl_workspace_bytes(n_img, H, W);
  if (!workspace || workspace_bytes < need) return VAEMDL_EWORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 7u) return VAEMDL_EALIGN;
  char* ws = static_cast<char*>(workspace);
  ModlArgs a{};
  a.params = params;
  a.x = x;
  a.partial = reinterpret_cast<double*>(ws);
  a.g_image = g_ll;
  a.dparams = dparams;
  a.HW = HW;
  a.n_px = n_px;
  a.x_batch = x_batch;
  a.x_u8 = x_dtype == VAEMDL_X_U8;
  a.x_unit = x_range == VAEMDL_RANGE_UNIT;
  a.edge_openai = edge_mode == VAEMDL_EDGE_OPENAI;
  a.M = M;
  a.plain = AR;
  a.spread = spread_runs();
  StepFinish f{};
  f.extra = extra;
  f.ll = ll_image;
  f.ll64 = ll_image_f64;
  f.log_w = log_w;
  f.lme_b = lme_b;
  f.elbo = elbo;
  f.g_ll = g_ll;
  f.lme64 = reinterpret_cast<double*>(ws + partial_elems(n_img) * sizeof(double));
  f.B = B;
  f.S = S;
  f.b_norm = static_cast<float>(B_total > 0 ? B_total : B);
  f.small = n_px < (1ll << 31);
  if (launches) *launches = 1;
  switch (M) {
    case 5:
      return launch_step<5, 1, AR>(a, f, n_img, st);
    case 10:
      return launch_step<10, 1, AR>(a, f, n_img, st);
    case 20:
      return launch_step<10, 2, AR>(a, f, n_img, st);
    default:
      return launch_step<10, 3, AR>(a, f, n_img, st);
  }
}
}  // namespace vaemdl
