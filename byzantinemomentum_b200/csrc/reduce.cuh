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
