#pragma once
// modl_core.cuh -- kernel arguments, pixel decoding, bf16 widening / narrowing, log-domain fallback, scalar and packed
// sub-pixel arithmetic (pair_eval) shared by every MoDL kernel.
// Part of the MoDL kernel family; see modl_kernels.cuh for the overview.
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
  float2* pix_stats;  // [n_px] nullable: (mixture sum S, logit normaliser SW) of every pixel-sample.  The forward kernel
                      // writes them when given; a backward kernel instantiated with ST reads them and scales its gradients
                      // in the same pass that forms them (no second pass over the row, no aux strip)
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
  int tm_refill;              // tensor-memory gradient kernel: the component pair of the first pass before which the staging
                              // slot is refilled with the warp's next tile (modl_tm.cuh)
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
  int pair_rot;  // tiled kernels: walk the component pairs in a per-lane rotated order (shared-memory bank conflicts)
  // run-time tile geometry (modl_rt_kernel: any n_mix without its own instantiation)
  int rt_MC, rt_LPP, rt_PPT, rt_rot, rt_warp_f;
  // bin geometry of the un-conditioned pixel mixture (utils/mdl_plain.py:18, utils/discretized_logistic.py:10-21): the
  // x-conditioned classes (AR = 0) are fixed to 8-bit pixels on [-1, 1] and keep kDx / kWidth / kLsNarrow as compile-time
  // constants; the AR = 1 instantiations read these
  float low, high;   // edge tests x <= low / x >= high (AR = 0: -1 / 1)
  float dx, width;   // half bin width, bin width = (high - low) / (levels - 1)
  float ls_narrow;   // log-scales below this have h = exp(-ls) * dx >= kHSmall
};
template <int AR>
__device__ __forceinline__ float bin_dx(const ModlArgs& a) { return AR ? a.dx : kDx; }
template <int AR>
__device__ __forceinline__ float bin_width(const ModlArgs& a) { return AR ? a.width : kWidth; }

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
  float dx = kDx, width = kWidth;  // bin geometry (compile-time constants for the x-conditioned classes)
};
struct PixelPair {  // two pixels: lo half = pixel A, hi half = pixel B (the pixel-pair kernel for small n_mix)
  f2 x[3];
  bool ll[3], lh[3], rl[3], rh[3];
  float dx = kDx, width = kWidth;
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
    px.left[c] = a.edge_openai ? (v < -0.999f) : (v <= a.low);
    px.right[c] = a.edge_openai ? (v > 0.999f) : (v >= a.high);
  }
  px.dx = a.dx;
  px.width = a.width;
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
    t += subpix_logf(px.x[c], px.left[c], px.right[c], loc, ls, px.dx, px.width);
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
  subpix<false>(px.x[0], px.left[0], px.right[0], loc0, fmaxf(s[0], -7.0f), px.dx, px.width, f0);
  subpix<false>(px.x[1], px.left[1], px.right[1], loc1, fmaxf(s[1], -7.0f), px.dx, px.width, f1);
  subpix<false>(px.x[2], px.left[2], px.right[2], loc2, fmaxf(s[2], -7.0f), px.dx, px.width, f2);
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
  for (int c = 0; c < 3; ++c) subpix<true>(px.x[c], px.left[c], px.right[c], loc[c], fmaxf(s[c], -7.0f), px.dx, px.width, f[c]);
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
__device__ __forceinline__ void subpix2(f2 x, EdgeFlags e, f2 loc, f2 s_raw, Sub2& o, float dx, float width) {
  const f2 ls = max_2(s_raw, -7.0f);                                  // utils/mdl.py:109
  const f2 inv = ex2_2(ls * (-kLog2e));
  const f2 mid = inv * (x - loc);
  const f2 A = ex2_negabs_2(mid * kLog2e);
  const f2 h = inv * dx;
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
  const f2 num_l = (A * inv) * width;
  const f2 den_l = opA * opA;
  f2 num = sel_2(il, ih, num_n, num_l);
  f2 den = sel_2(il, ih, den_n, den_l);
  const bool el = e.ll || e.rl, eh = e.lh || e.rh;
  // (measured, tools/ab_lib.sh: working these two compares out only inside the edge branch removes 2-4 % of the instructions and
  //  makes both kernels 1.2-1.5 % SLOWER at the headline shape -- they stay loop-level code)
  const bool ool = (e.ll == (lo(mid) >= 0.0f)), ooh = (e.lh == (hi(mid) >= 0.0f));  // 1/(1+AG) vs A/(A+G)
  if (el || eh) {
    num = sel_2(el, eh, sel_2(ool, ooh, sp(1.0f), A), num);
    den = sel_2(el, eh, sel_2(ool, ooh, opAG, ApG), den);
  }
  o.num = num;
  o.den = den;
  if constexpr (BWD) {
    const f2 omA2 = (sp(1.0f) - A) * opA;
    // "value in the CDF-difference branch, 0 in the low-probability branch" as a multiplication by a 0/1 mask on the
    // packed pipe: a select of a packed value costs the compiler four predicated moves (they were a quarter of the
    // backward loop), the mask two compares.  Every masked quantity is finite (h <= e^7 / 255), so x * 0 is 0.
    const f2 mi = pk(il ? 1.0f : 0.0f, ih ? 1.0f : 0.0f);
    f2 gsel;  // G in the CDF-difference branch, 1 in the low-probability branch
    if constexpr (NARROW)
      gsel = sel_2(il, ih, G, sp(1.0f));
    else
      gsel = fma2(mi, omG * -1.0f, 1.0f);  // mi = 1: 1 - omG, the very expression G was formed with
    f2 nm = neg_sign_of_2(gsel * omA2, mid);
    f2 nh = mi * ((h * -1.0f) * num_n);
    const f2 h2 = h * h;
    f2 hc = fma2(h2, 2.0f / 945.0f, -1.0f / 45.0f);
    hc = fma2(h2, hc, 1.0f / 3.0f);
    hc = fma2(h2, hc, 1.0f);
    if constexpr (NARROW) {
      const f2 e = h * fma2(G, G, 1.0f) * rcp_2(rest_n);
      hc = sel_2(nl, nh_, e, hc);
    }
    f2 c0 = mi * hc;
    f2 dir = mi + -1.0f;
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
  for (int c = 0; c < 3; ++c) subpix2<NARROW, BWD>(px_x(px, c), px_edge(px, c), loc[c], s[c], f[c], px.dx, px.width);
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
      // tf.maximum(logscale, -7): the gradient reaches logscale iff logscale >= -7 (0/1 mask, see subpix2)
      dls = dls * pk(lo(s[c]) >= -7.0f ? 1.0f : 0.0f, hi(s[c]) >= -7.0f ? 1.0f : 0.0f);
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

}  // namespace vaemdl
