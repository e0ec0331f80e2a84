// k1_select.cuh — K1: coordinate-wise order statistics down the d axis.
//
// One thread owns VEC adjacent coordinates: it loads the n values of each coordinate with
// one coalesced vector load per row (a warp reads 128·VEC contiguous bytes of every row),
// keeps them in registers and runs a literal-index sorting network on them — no shared
// memory, no stack of the n rows (the `torch.stack` of median.py:39 / trmean.py:79 is gone).
//
//   k1_median<N, VEC>            median.py:31-39        FMNMX.NAN network, pruned to one rank
//   k1_sorted<N, VEC, F, MODE>   trmean.py:24-50,69-109 sort, then trmean / phocas / meamed
//
// F / MODE = -1 mean "run-time value": one generic kernel per (N, VEC) serves every f and the
// three epilogues after a full sort.  For the (n, f) pairs of the reference's experiment
// grids the trimmed mean is also instantiated with F and MODE fixed: the compiler's dead-code
// elimination then prunes the network to the comparators ranks f..n-f-1 depend on.
//
// Roofline: HBM (n·4 B read + 4 B written per coordinate); secondary bound: the ALU pipe
// (FMNMX issues at 64 lanes/clk/SM): see DESIGN.md for the per-N operation counts.
#pragma once

#include "common.cuh"
#include "launch.cuh"
#include "networks_gen.cuh"

namespace bz {

constexpr int kK1Threads = 128;

// ---- median ---------------------------------------------------------------------------

template <int N, int VEC>
__global__ void __launch_bounds__(kK1Threads)
k1_median(const __grid_constant__ RowTable rows, const Geom g, float* __restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * kK1Threads + threadIdx.x;
  if (v >= g.nv) return;
  const int64_t e0 = v * VEC - g.shift;
  const bool full = e0 >= 0 && e0 + VEC <= g.d;
  float x[VEC][N];
  load_rows<N, VEC>(rows, e0, g.d, full, x);
  float res[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    SortNet<N>::template run<OpsNaNProp>(x[c]);
    res[c] = x[c][(N - 1) / 2];   // lower median; every other output is dead code
  }
  store_vec<VEC>(out, e0, g.d, full, res);
}

// ---- trimmed mean of a sorted column (trmean.py:33) -----------------------------------------
// `values[f:-f].mean(dim=0)`: ATen sums rows in a cascade of 16-row blocks (sequential when
// R = n - 2f <= 16), then divides by R with one IEEE division.
template <int N>
__device__ __forceinline__ float trmean_sorted(const float (&s)[N], const int f) {
  const int R = N - 2 * f;
  const int hi = N - f;
  float acc0 = 0.f;
  if (N - 2 <= 16 || R <= 16) {
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (k >= f && k < hi) acc0 = __fadd_rn(acc0, s[k]);
    return __fdiv_rn(acc0, (float)R);
  }
  float acc1 = 0.f;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (k >= f && k < hi) {
      acc0 = __fadd_rn(acc0, s[k]);
      if (((k - f) & 15) == 15) { acc1 = __fadd_rn(acc1, acc0); acc0 = 0.f; }
    }
  }
  return __fdiv_rn(__fadd_rn(acc0, acc1), (float)R);
}

// ---- mean of the m values closest to c (trmean.py:35-50) -------------------------------------
// In a sorted column the m closest values form a window [l, l+m).  It always holds the core
// [R, m) (R = N - m removals) when m >= R, and of each pair (s[l], s[l+m]), l < R, the closer
// one.  NaN distances count as largest (`topk(largest=False)`); a NaN centre gives NaN.
// The partner index l+m is run-time uniform, so the sorted column is staged in shared
// memory (`col`, stride `stride` floats between consecutive ranks).
template <int N>
__device__ __forceinline__ float closest_pairs(const float (&s)[N], int m, float c, float* col, int stride) {
  const int R = N - m;            // removals; requires m >= R (phocas/meamed: m = n-f > f)
#pragma unroll
  for (int k = 0; k < N; ++k) col[k * stride] = s[k];
  // (own column only: no barrier needed)
  float acc = 0.f;
  for (int l = 0; l < R; ++l) {
    const float lo = col[l * stride], hi = col[(l + m) * stride];
    const int dlo = abs_key(__fsub_rn(lo, c)), dhi = abs_key(__fsub_rn(hi, c));
    acc = __fadd_rn(acc, (dlo > dhi) ? hi : lo);
  }
  for (int k = R; k < m; ++k) acc = __fadd_rn(acc, col[k * stride]);
  const float r = __fdiv_rn(acc, (float)m);
  return (c != c) ? quiet_nan() : r;
}

template <int N, int VEC, int F, int MODE>
__global__ void __launch_bounds__(kK1Threads)
k1_sorted(const __grid_constant__ RowTable rows, const Geom g, const int mode_rt, const int f_rt,
          float* __restrict__ out) {
  extern __shared__ float smem[];   // closest modes only: [N][VEC][kK1Threads]
  const int mode = (MODE >= 0) ? MODE : mode_rt;
  const int f = (F >= 0) ? F : f_rt;
  const int64_t v = (int64_t)blockIdx.x * kK1Threads + threadIdx.x;
  if (v >= g.nv) return;
  const int64_t e0 = v * VEC - g.shift;
  const bool full = e0 >= 0 && e0 + VEC <= g.d;
  float x[VEC][N];
  load_rows<N, VEC>(rows, e0, g.d, full, x);
  float chk = 0.f;
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < VEC; ++c)
      chk = fmaf(x[c][r], 0.f, chk);   // NaN iff some value is NaN or +-inf (FMA pipe, off the ALU pipe)
  float res[VEC];
  if (chk == chk) {
    // Fast path: all finite, plain FMNMX network
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      SortNet<N>::template run<OpsFast>(x[c]);
      if (MODE == kModeTrmean) res[c] = trmean_sorted<N>(x[c], f);
    }
  } else {
    // Non-finite values present: sort integer keys so that NaN sorts last and +-inf keep their place
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      int k[N];
#pragma unroll
      for (int r = 0; r < N; ++r) k[r] = float_to_key(x[c][r]);
      SortNet<N>::template run<OpsKey>(k);
#pragma unroll
      for (int r = 0; r < N; ++r) x[c][r] = key_to_float(k[r]);
      if (MODE == kModeTrmean) res[c] = trmean_sorted<N>(x[c], f);
    }
  }
  if (MODE != kModeTrmean) {
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      if (mode == kModeTrmean) {
        res[c] = trmean_sorted<N>(x[c], f);
      } else {
        float center;
        if (mode == kModePhocas) {
          center = trmean_sorted<N>(x[c], f);
        } else {
          // lower median; NaN sorts last, so a NaN anywhere shows in the last rank (median.py:39)
          center = (x[c][N - 1] != x[c][N - 1]) ? quiet_nan() : x[c][(N - 1) / 2];
        }
        res[c] = closest_pairs<N>(x[c], N - f, center, smem + c * kK1Threads + threadIdx.x, VEC * kK1Threads);
      }
    }
  }
  store_vec<VEC>(out, e0, g.d, full, res);
}

}  // namespace bz
