// common.cuh — shared device/host helpers of libbyzagg (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/byzagg.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libbyzagg is written for sm_100a (B200) only"
#endif

namespace bz {

constexpr int kMaxN = BZ_MAX_N;

// Host array of n device row pointers, passed BY VALUE as a kernel parameter (512 B in the
// constant bank): no torch.stack, no H2D copy of a pointer table.
struct RowTable {
  const float* p[kMaxN];
};

// Element ranges a launch covers.  A vectorised launch covers [base0, base0 + cnt0*VEC); the
// scalar "edges" launch covers the unaligned head [base0, base0+cnt0) and tail [base1, base1+cnt1).
struct Span {
  int64_t base0, cnt0, base1, cnt1;
};

// How a [d]-long set of rows splits into an aligned vector body plus scalar edges.
struct Split {
  int     vec;     // 4, 2 or 1 elements per thread in the body
  int64_t head;    // scalar elements before the body
  int64_t nvec;    // vectors in the body
  int64_t tail;    // scalar elements after the body
};

// Body vector width: all rows (and out) must share the same misalignment modulo the vector size.
Split make_split(const float* const* rows, int n, const void* out, const void* extra, int64_t d, int want_vec);

// Error plumbing (api.cu)
int  fail(int code, const char* fmt, ...);
int  check_launch(const char* what);

// ---- device helpers ---------------------------------------------------------------------

#ifdef __CUDACC__

template <int VEC> struct VecLoad;
template <> struct VecLoad<4> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&o)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(o[0], o[1], o[2], o[3]));
  }
};
template <> struct VecLoad<2> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[2]) {
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    o[0] = t.x; o[1] = t.y;
  }
  static __device__ __forceinline__ void store(float* p, const float (&o)[2]) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(o[0], o[1]));
  }
};
template <> struct VecLoad<1> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[1]) { o[0] = __ldcs(p); }
  static __device__ __forceinline__ void store(float* p, const float (&o)[1]) { __stcs(p, o[0]); }
};

// First element handled by logical thread i of a launch over `s` with VEC elements per thread.
template <int VEC>
__device__ __forceinline__ int64_t span_element(const Span& s, int64_t i) {
  return (i < s.cnt0) ? s.base0 + i * VEC : s.base1 + (i - s.cnt0) * VEC;
}

// Compare-exchange policies for SortNet<N>::run<Ops>().
// Fast: inputs hold no NaN (plain FMNMX).
struct OpsFast {
  static __device__ __forceinline__ void ce(float& a, float& b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo; b = hi;
  }
};
// NaN-propagating (FMNMX.NAN): a NaN input poisons every output that depends on it, which
// is exactly torch's `median(dim)` semantics (median.py:39 with torch >= 1.7).
struct OpsNaNProp {
  static __device__ __forceinline__ void ce(float& a, float& b) {
    float lo, hi;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(lo) : "f"(a), "f"(b));
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(hi) : "f"(a), "f"(b));
    a = lo; b = hi;
  }
};
// Total order on keys: signed-int image of the float order with every NaN mapped to INT_MAX
// (NaN sorts last, like `Tensor.sort` / `topk(largest=False)`).
struct OpsKey {
  static __device__ __forceinline__ void ce(int& a, int& b) {
    const int lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
  }
};
__device__ __forceinline__ int float_to_key(float x) {
  int k = __float_as_int(x);
  k ^= (k >> 31) & 0x7fffffff;
  return (x != x) ? 0x7fffffff : k;
}
__device__ __forceinline__ float key_to_float(int k) {
  k ^= (k >> 31) & 0x7fffffff;
  return __int_as_float(k);   // INT_MAX -> 0x7fffffff, a quiet NaN
}
// |x| as an integer whose order is the order of |x| with NaN largest.
__device__ __forceinline__ int abs_key(float x) { return __float_as_int(x) & 0x7fffffff; }

__device__ __forceinline__ float quiet_nan() { return __int_as_float(0x7fc00000); }

#endif  // __CUDACC__

}  // namespace bz
