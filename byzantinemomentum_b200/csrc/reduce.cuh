// reduce.cuh — warp-level helpers shared by K2 and K2'.
#pragma once

#include "common.cuh"

namespace bz {

typedef unsigned long long u64;

// Packed fp32 pairs (two coordinates per instruction: SASS FADD2 / FFMA2)
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 d;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// IEEE division of two lanes by a small positive integer m known at compile time (after unrolling),
// without the generic division's reciprocal, range check and slow-path call per lane:
//   y = RN(1/m),  q0 = RN(a*y),  r = a - q0*m (one FMA: exact),  q = RN(q0 + r*y)
// is the correctly rounded a/m for EVERY float a with 2^-100 <= |a| <= FLT_MAX and every m in 1..64:
// checked exhaustively, all 2^32 operands for each m (tools/divcheck.c; what fails outside that range:
// +-inf -> NaN, -0 -> +0, and for even m that are not powers of two the operands below 2^-122, whose
// remainder underflows).  Everything else — zeros, subnormals, inf, NaN — takes the generic division.
constexpr float kDivTiny = 7.888609052210118e-31f;      // 2^-100
__device__ __forceinline__ bool div_fast_ok(float a) {
  return fabsf(a) >= kDivTiny && fabsf(a) <= 3.402823466e+38f;
}
__device__ __forceinline__ float div_small(float a, float m) {
  const float y = 1.f / m;                    // m is a literal after unrolling: folded at compile time (IEEE division, no fast-math)
  const float q0 = __fmul_rn(a, y);
  const float r = __fmaf_rn(-q0, m, a);
  const float q = __fmaf_rn(r, y, q0);
  return div_fast_ok(a) ? q : __fdiv_rn(a, m);
}
__device__ __forceinline__ u64 div_small2(u64 a, float m) {
  const float y = 1.f / m;                    // m is a literal after unrolling: folded at compile time (IEEE division, no fast-math)
  const u64 y2 = pack2(y, y);
  const u64 q0 = mul2(a, y2);
  const u64 r = fma2(q0, pack2(-m, -m), a);
  u64 q = fma2(r, y2, q0);
  float lo, hi;
  unpack2(a, lo, hi);
  if (!(div_fast_ok(lo) && div_fast_ok(hi))) q = pack2(__fdiv_rn(lo, m), __fdiv_rn(hi, m));
  return q;
}
__device__ __forceinline__ float half_sum(u64 v) {
  return __fadd_rn(__uint_as_float((unsigned)(v & 0xffffffffull)), __uint_as_float((unsigned)(v >> 32)));
}


// Transposed warp reduction: on return, lane L holds sum over lanes of v[L] (v has 32 slots).
__device__ __forceinline__ float transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int j = 0; j < h; ++j) {
      const float send = up ? v[j] : v[j + h];
      const float keep = up ? v[j + h] : v[j];
      v[j] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, h));
    }
  }
  return v[0];
}

}  // namespace bz
