// packed.cuh -- two-wide float32 arithmetic on Blackwell's packed FP32 pipe (PTX add/mul/fma.f32x2 -> SASS FADD2 /
// FMUL2 / FFMA2: two FP32 results per issue slot).  The MoDL kernels process mixture components in pairs, one
// component per half, because the hot loop is bound by instruction issue, not by DRAM (profiles/r01_*).
#pragma once
#include "common.cuh"

namespace vaemdl {

struct f2 {
  unsigned long long v;
};

__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 sp(float a) { return pk(a, a); }
__device__ __forceinline__ float lo(f2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  return x;
}
__device__ __forceinline__ float hi(f2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  return y;
}
__device__ __forceinline__ f2 operator+(f2 a, f2 b) {
  f2 r;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 operator-(f2 a, f2 b) {
  f2 r;
  asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 operator*(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f2 operator+(f2 a, float b) { return a + sp(b); }
__device__ __forceinline__ f2 operator*(f2 a, float b) { return a * sp(b); }
__device__ __forceinline__ f2 fma2(f2 a, float b, float c) { return fma2(a, sp(b), sp(c)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, float c) { return fma2(a, b, sp(c)); }
__device__ __forceinline__ f2 fma2(f2 a, float b, f2 c) { return fma2(a, sp(b), c); }

// component-wise helpers on the scalar pipes (MUFU / ALU)
__device__ __forceinline__ f2 ex2_2(f2 a) { return pk(ex2a(lo(a)), ex2a(hi(a))); }
__device__ __forceinline__ f2 ex2_negabs_2(f2 a) { return pk(ex2a(-fabsf(lo(a))), ex2a(-fabsf(hi(a)))); }
__device__ __forceinline__ f2 rcp_2(f2 a) { return pk(rcpa(lo(a)), rcpa(hi(a))); }
__device__ __forceinline__ f2 max_2(f2 a, float c) { return pk(fmaxf(lo(a), c), fmaxf(hi(a), c)); }
__device__ __forceinline__ f2 min_2(f2 a, float c) { return pk(fminf(lo(a), c), fminf(hi(a), c)); }
__device__ __forceinline__ f2 sel_2(bool pl, bool ph, f2 a, f2 b) { return pk(pl ? lo(a) : lo(b), ph ? hi(a) : hi(b)); }
// copysign(|mag|, -ref) component-wise: flips mag (>= 0) to the opposite sign of ref
__device__ __forceinline__ f2 neg_sign_of_2(f2 mag, f2 ref) {
  const unsigned ml = __float_as_uint(lo(mag)), mh = __float_as_uint(hi(mag));
  const unsigned rl = __float_as_uint(lo(ref)), rh = __float_as_uint(hi(ref));
  return pk(__uint_as_float(ml | (~rl & 0x80000000u)), __uint_as_float(mh | (~rh & 0x80000000u)));
}

}  // namespace vaemdl
