// reduce.cuh — warp-level helpers shared by K2 and K2'.
#pragma once

#include "common.cuh"

namespace bz {

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
