// k5_score.cu — K5: scoring and selection on the n x n distance table, one CTA, on device.
// The reference does this part in Python on the host after n(n-1)/2 `.item()` syncs
// (krum.py:49-62, bulyan.py:56-62, brute.py:47-68, aksel.py:49, cge.py:38); here it chains on
// the stream between the distance pass and the reduce pass, and must reproduce Python's
// behaviour exactly: distances are fp32 values widened to double, non-finite -> +inf
// (krum.py:46-47), `sum()` of floats is CPython's (>= 3.12: Neumaier-compensated), sorts are
// stable, brute enumerates `itertools.combinations` in lexicographic order with a strict `<`.
#include <math_constants.h>

#include "dist.cuh"

namespace bz {

constexpr int kK5Threads = 1024;

__device__ __forceinline__ bool finite_d(double x) { return fabs(x) <= 1.7976931348623157e308; }

// Where the partial blocks live: one strided buffer (per-CTA blocks of K2 / K2', or blocks gathered
// by a collective), or one pointer per peer GPU — the blocks are then read IN PLACE from the
// peers' memory over NVLink (ld.global on peer-mapped addresses): the exchange step of the
// d-sharded path is fused into the scoring kernel, no all-gather, no staging copy.
struct StridedParts {
  const double* base;
  size_t stride;
  __device__ __forceinline__ const double* block(int p) const { return base + (size_t)p * stride; }
};
constexpr int kMaxPeers = BZ_MAX_PEERS;
struct PeerParts {
  const double* ptr[kMaxPeers];
  __device__ __forceinline__ const double* block(int p) const { return ptr[p]; }
};

// Block-wide, deterministic sum of `nparts` partial blocks of `len` doubles into smem `out`.
// The parts are split in `slices` interleaved classes (p mod slices); thread (slice, t) walks
// entries t, t + T, ... and adds its class in ascending p; the classes are then added in
// ascending order.  The order depends only on (nparts, slices): bitwise reproducible, and the
// same on every rank of the sharded path.  All loads of a thread are independent across
// entries, so a thread keeps several in flight (the naive per-entry loop over 148 parts cost
// ~20 us of pure latency).
template <class Parts>
__device__ void block_sum_parts(const Parts& parts, int nparts, int len, int slices,
                                double* __restrict__ scratch, double* __restrict__ out) {
  const int T = blockDim.x / slices;            // threads per slice
  const int slice = threadIdx.x / T, t = threadIdx.x - slice * T;
  constexpr int E = 8;                          // entries per thread in flight together
  if (slice < slices) {
    for (int e0 = t; e0 < len; e0 += T * E) {
      double s[E];
#pragma unroll
      for (int q = 0; q < E; ++q) s[q] = 0.;
#pragma unroll 4
      for (int p = slice; p < nparts; p += slices) {
        const double* __restrict__ row = parts.block(p);
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int e = e0 + q * T;
          if (e < len) s[q] += row[e];
        }
      }
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int e = e0 + q * T;
        if (e < len) scratch[slice * len + e] = s[q];
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < len; e += blockDim.x) {
    double s = 0.;
    for (int k = 0; k < slices; ++k) s += scratch[k * len + e];
    out[e] = s;
  }
  __syncthreads();
}

// Aliased rows (the f Byzantine gradients are ONE tensor repeated, attacks/identical.py:86) are
// detected on the host by pointer equality; the distance pass then runs on the u unique rows only
// (with the self pairs i == i) and this map expands its u x u table to the n x n one.
struct RowMap {
  unsigned char to_unique[kMaxN];
  int u;          // number of unique rows (u == n: identity, no expansion)
};

// dist[i][j] = double(fl32(sqrt(sum_k (x_i - x_j)^2))) from the summed u x u table `sq`;
// diagonal = `diag`.  `sq` and `dist` must not overlap when map.u < n.
__device__ void finish_distances_mapped(const double* sq, const RowMap& map, int n, bool map_nonfinite, double diag, double* dist) {
  const int u = map.u;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    if (i < j) {
      // aliases (a == b) read the diagonal: 0 for a finite row, NaN for a row holding NaN / inf,
      // exactly what the reference's x.sub(x).norm() gives
      const int a = map.to_unique[i], b = map.to_unique[j];
      double v = (double)(float)sqrt(sq[min(a, b) * u + max(a, b)]);
      if (map_nonfinite && !finite_d(v)) v = CUDART_INF;
      dist[i * n + j] = v;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    if (i > j) dist[e] = dist[j * n + i];
    else if (i == j) dist[e] = diag;
  }
  __syncthreads();
}

// dist[i][j] = double(fl32(sqrt(sum_k (x_i - x_j)^2))) from the summed table `sq` (in place
// allowed); diagonal = `diag`.
__device__ void finish_distances(const double* sq, int n, bool map_nonfinite, double diag, double* dist) {
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    if (i < j) {
      double v = (double)(float)sqrt(sq[e]);
      if (map_nonfinite && !finite_d(v)) v = CUDART_INF;
      dist[i * n + j] = v;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    if (i > j) dist[e] = dist[j * n + i];
    else if (i == j) dist[e] = diag;
  }
  __syncthreads();
}

// CPython >= 3.12 `sum()` over floats (Objects/bltinmodule.c: Neumaier's compensated sum;
// the compensation is dropped when it is not finite).
__device__ double py_sum(const double* v, int count) {
  if (count <= 0) return 0.;
  double f = v[0], c = 0.;
  for (int k = 1; k < count; ++k) {
    const double x = v[k];
    const double t = f + x;
    if (fabs(f) >= fabs(x)) c += (f - t) + x;
    else                    c += (x - t) + f;
    f = t;
  }
  if (c != 0. && finite_d(c)) f += c;
  return f;
}

// Row-wise ascending sort by ranking: sorted[i][rank] = dist[i][j] (ties by column index).
__device__ void sort_rows(const double* dist, double* sorted, int n) {
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    const double v = dist[e];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const double w = dist[i * n + k];
      rank += (w < v || (w == v && k < j)) ? 1 : 0;
    }
    sorted[i * n + rank] = v;
  }
  __syncthreads();
}

// Stable ascending argsort of key[0..n): order[rank] = index.  NaN keys sort last.
__device__ void stable_order(const double* key, int n, int32_t* __restrict__ order) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = key[i];
    const bool vn = v != v;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double w = key[j];
      const bool wn = w != w;
      const bool less = vn ? (!wn) : (!wn && w < v);
      const bool same = vn ? wn : (w == v);
      rank += (less || (same && j < i)) ? 1 : 0;
    }
    order[rank] = i;
  }
}

// krum.py:52-62 (count = n-f-1, diagonal excluded) and bulyan.py:56-62 (count = m over the
// row INCLUDING its +inf diagonal).  Putting +inf on the diagonal serves both: the extra +inf
// can only be reached after every finite distance, where the sum is +inf either way.
template <class Parts>
__global__ void __launch_bounds__(kK5Threads)
k5_score_select(const __grid_constant__ Parts parts, const __grid_constant__ RowMap map, int nparts, int n, int count,
                int32_t* __restrict__ order, int32_t* __restrict__ status, int f, int m, int bulyan, int slices) {
  extern __shared__ double sm[];
  double* dist = sm;
  double* sorted = sm + n * n;
  double* score = sm + 2 * n * n;
  double* scratch = score + n;
  if (map.u == n) {
    block_sum_parts(parts, nparts, n * n, slices, scratch, dist);
    finish_distances(dist, n, true, CUDART_INF, dist);
  } else {
    block_sum_parts(parts, nparts, map.u * map.u, slices, scratch, sorted);   // u x u table, parked in `sorted`
    finish_distances_mapped(sorted, map, n, true, CUDART_INF, dist);
  }
  sort_rows(dist, sorted, n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) score[i] = py_sum(sorted + i * n, count);
  __syncthreads();
  stable_order(score, n, order);
  if (threadIdx.x == 0 && status != nullptr) {
    int st = BZ_STATUS_OK;
    if (bulyan) {
      // bulyan.py:65-73: from the second iteration on, pruned `(inf, None)` entries precede every
      // +inf score (stable sort), so fewer than m_i finite scores left means gradients[None]
      int finite = 0;
      for (int i = 0; i < n; ++i) finite += finite_d(score[i]) ? 1 : 0;
      const int m_max = n - f - 2, theta = n - 2 * f - 2;
      int mi = m;
      for (int i = 0; i < theta; ++i) {
        mi = min(mi, m_max - i);
        if (i >= 1 && finite - i < mi) st = BZ_STATUS_DEGENERATE;
      }
    }
    *status = st;
  }
}

// aksel.py:39-49 / cge.py:28-38: stable order of n keys.
constexpr int kRowSelThreads = 1024, kRowSelSlices = 16;   // 64 threads per slice: one per row
template <class Parts>
__global__ void __launch_bounds__(kRowSelThreads)
k5_rowdist_select(const __grid_constant__ Parts parts, int nparts, int n, int sqrt_norm, int32_t* __restrict__ order) {
  __shared__ double key[kMaxN];
  __shared__ double scratch[kRowSelSlices * kMaxN];
  block_sum_parts(parts, nparts, n, kRowSelSlices, scratch, key);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double s = key[i];
    double v;
    if (sqrt_norm) {
      v = (double)(float)sqrt(s);
      if (!finite_d(v)) v = CUDART_INF;
    } else {
      v = (double)(float)s;
    }
    key[i] = v;
  }
  __syncthreads();
  stable_order(key, n, order);
}

// ---- brute: exhaustive minimum-diameter subset (brute.py:47-68) -----------------------------
// Thread t scans a contiguous range of lexicographic ranks; the first strict minimum wins,
// so the block-wide winner is the smallest (diameter, thread) pair.
__device__ __forceinline__ unsigned long long sat_add(unsigned long long a, unsigned long long b) {
  const unsigned long long s = a + b;
  return s < a ? ~0ull : s;
}

template <class Parts>
__global__ void __launch_bounds__(kK5Threads)
k5_brute_select(const __grid_constant__ Parts parts, const __grid_constant__ RowMap map, int nparts, int n, int f,
                unsigned long long total, int32_t* __restrict__ sel, int32_t* __restrict__ status, int slices) {
  extern __shared__ double sm[];
  double* dist = sm;                                                  // n*n
  unsigned long long* binom = reinterpret_cast<unsigned long long*>(sm + n * n);   // (n+1)*(n+1)
  double* scratch = sm + n * n + (n + 1) * (n + 1);
  __shared__ double best_diam[kK5Threads / 32];
  __shared__ unsigned long long best_rank[kK5Threads / 32];
  const int k = n - f;
  const int W = n + 1;
  if (map.u == n) {
    block_sum_parts(parts, nparts, n * n, slices, scratch, dist);
    finish_distances(dist, n, false, 0., dist);
  } else {
    double* table = reinterpret_cast<double*>(binom);        // u x u table, parked where the Pascal triangle goes next
    block_sum_parts(parts, nparts, map.u * map.u, slices, scratch, table);
    finish_distances_mapped(table, map, n, false, 0., dist);
  }
  // Pascal triangle, row by row
  for (int a = 0; a <= n; ++a) {
    for (int b = threadIdx.x; b <= n; b += blockDim.x) {
      unsigned long long v;
      if (b == 0) v = 1;
      else if (b > a) v = 0;
      else v = sat_add(binom[(a - 1) * W + b - 1], binom[(a - 1) * W + b]);
      binom[a * W + b] = v;
    }
    __syncthreads();
  }
  // This thread's rank range
  const unsigned long long per = (total + blockDim.x - 1) / blockDim.x;
  const unsigned long long lo = per * threadIdx.x;
  const unsigned long long hi = (lo + per < total) ? lo + per : total;
  double my_diam = CUDART_INF;
  unsigned long long my_rank = ~0ull;
  bool found = false;
  if (lo < hi) {
    int comb[kMaxN];
    {  // unrank `lo`
      unsigned long long r = lo;
      int x = 0;
      for (int pos = 0; pos < k; ++pos) {
        while (true) {
          const unsigned long long c = binom[(n - 1 - x) * W + (k - 1 - pos)];
          if (c > r) break;
          r -= c;
          ++x;
        }
        comb[pos] = x++;
      }
    }
    for (unsigned long long rank = lo; rank < hi; ++rank) {
      double diam = 0.;
      bool ok = true;
      for (int a = 0; a < k - 1 && ok; ++a) {
        const double* row = dist + comb[a] * n;
        for (int b = a + 1; b < k; ++b) {
          const double v = row[comb[b]];
          if (!finite_d(v)) { ok = false; break; }
          if (v > diam) diam = v;
        }
      }
      if (ok && (!found || diam < my_diam)) { found = true; my_diam = diam; my_rank = rank; }
      // next combination in lexicographic order
      int pos = k - 1;
      while (pos >= 0 && comb[pos] == n - k + pos) --pos;
      if (pos < 0) break;
      ++comb[pos];
      for (int q = pos + 1; q < k; ++q) comb[q] = comb[q - 1] + 1;
    }
  }
  // Block-wide minimum of (found ? diam : +inf-with-no-rank, rank)
  double dm = found ? my_diam : CUDART_INF;
  unsigned long long rk = found ? my_rank : ~0ull;
  for (int h = 16; h >= 1; h >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, dm, h);
    const unsigned long long orank = __shfl_xor_sync(0xffffffffu, rk, h);
    if (orank != ~0ull && (rk == ~0ull || od < dm || (od == dm && orank < rk))) { dm = od; rk = orank; }
  }
  if ((threadIdx.x & 31) == 0) { best_diam[threadIdx.x >> 5] = dm; best_rank[threadIdx.x >> 5] = rk; }
  __syncthreads();
  if (threadIdx.x == 0) {
    dm = CUDART_INF; rk = ~0ull;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      const double od = best_diam[w];
      const unsigned long long orank = best_rank[w];
      if (orank != ~0ull && (rk == ~0ull || od < dm || (od == dm && orank < rk))) { dm = od; rk = orank; }
    }
    if (rk == ~0ull) {
      if (status) *status = BZ_STATUS_NO_FINITE_SET;
      for (int pos = 0; pos < k; ++pos) sel[pos] = pos;
    } else {
      if (status) *status = BZ_STATUS_OK;
      unsigned long long r = rk;
      int x = 0;
      for (int pos = 0; pos < k; ++pos) {
        while (true) {
          const unsigned long long c = binom[(n - 1 - x) * W + (k - 1 - pos)];
          if (c > r) break;
          r -= c;
          ++x;
        }
        sel[pos] = x++;
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------

// Interleaved part classes for block_sum_parts: as many as fit ~48 KB of scratch, at most 8.
static int pick_slices(int n, int nparts) {
  int slices = (int)((48 * 1024) / ((size_t)n * n * sizeof(double)));
  if (slices > 8) slices = 8;
  if (slices > nparts) slices = nparts;
  if (slices < 1) slices = 1;
  while (kK5Threads % slices) --slices;
  return slices;
}

template <class K>
static void opt_in_once(K kernel, size_t bytes, unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(mask & bit)) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    mask |= bit;
  }
}

constexpr size_t kScoreSmemMax = (size_t)(2 * kMaxN * kMaxN + kMaxN) * sizeof(double) + 48 * 1024;
constexpr size_t kBruteSmemMax = (size_t)kMaxN * kMaxN * sizeof(double) + (size_t)(kMaxN + 1) * (kMaxN + 1) * sizeof(unsigned long long) + 48 * 1024;

static RowMap identity_map(int n) {
  RowMap map;
  for (int i = 0; i < kMaxN; ++i) map.to_unique[i] = (unsigned char)(i < n ? i : 0);
  map.u = n;
  return map;
}

static RowMap make_map(const int* to_unique, int n, int u) {
  if (to_unique == nullptr || u >= n) return identity_map(n);
  RowMap map;
  for (int i = 0; i < kMaxN; ++i) map.to_unique[i] = (unsigned char)(i < n ? to_unique[i] : 0);
  map.u = u;
  return map;
}

template <class Parts>
static void score_select(const Parts& parts, const RowMap& map, int nparts, int n, int count, int32_t* order, int32_t* status, int f, int m,
                         int bulyan, cudaStream_t st) {
  const int slices = pick_slices(n, nparts);
  const size_t smem = (size_t)(2 * n * n + n + slices * n * n) * sizeof(double);
  static unsigned long long opted = 0;
  opt_in_once(k5_score_select<Parts>, kScoreSmemMax, opted);
  k5_score_select<Parts><<<1, kK5Threads, smem, st>>>(parts, map, nparts, n, count, order, status, f, m, bulyan, slices);
}

template <class Parts>
static int brute_select(const Parts& parts, const RowMap& map, int nparts, int n, int f, int32_t* sel, int32_t* status, cudaStream_t st) {
  // C(n, n-f) on the host, saturating
  const int k = n - f;
  unsigned long long total = 1;
  for (int i = 1; i <= (k < n - k ? k : n - k); ++i) {
    const unsigned long long num = (unsigned long long)(n - i + 1);
    if (total > (~0ull) / num) { total = ~0ull; break; }
    total = total * num / i;
  }
  if (total > (1ull << 31)) return -1;
  const int slices = pick_slices(n, nparts);
  const size_t smem = (size_t)n * n * sizeof(double) + (size_t)(n + 1) * (n + 1) * sizeof(unsigned long long) +
                      (size_t)slices * n * n * sizeof(double);
  static unsigned long long opted = 0;
  opt_in_once(k5_brute_select<Parts>, kBruteSmemMax, opted);
  k5_brute_select<Parts><<<1, kK5Threads, smem, st>>>(parts, map, nparts, n, f, total, sel, status, slices);
  return 0;
}

static PeerParts make_peers(const double* const* ptrs, int npeers) {
  PeerParts p;
  for (int r = 0; r < kMaxPeers; ++r) p.ptr[r] = ptrs[r < npeers ? r : 0];
  return p;
}

// to_unique / u: optional alias map (see RowMap); the blocks are then u x u tables.
void launch_krum_select(const double* parts, int nparts, int n, int f, int32_t* order, cudaStream_t st, const int* to_unique, int u) {
  const RowMap map = make_map(to_unique, n, u);
  score_select(StridedParts{parts, (size_t)map.u * map.u}, map, nparts, n, n - f - 1, order, nullptr, f, 0, 0, st);
}
void launch_bulyan_select(const double* parts, int nparts, int n, int f, int m, int32_t* order, int32_t* status, cudaStream_t st, const int* to_unique, int u) {
  const RowMap map = make_map(to_unique, n, u);
  score_select(StridedParts{parts, (size_t)map.u * map.u}, map, nparts, n, m, order, status, f, m, 1, st);
}
int launch_brute_select(const double* parts, int nparts, int n, int f, int32_t* sel, int32_t* status, cudaStream_t st, const int* to_unique, int u) {
  const RowMap map = make_map(to_unique, n, u);
  return brute_select(StridedParts{parts, (size_t)map.u * map.u}, map, nparts, n, f, sel, status, st);
}
void launch_rowdist_select(const double* parts, int nparts, int n, int sqrt_norm, int32_t* order, cudaStream_t st) {
  k5_rowdist_select<StridedParts><<<1, kRowSelThreads, 0, st>>>(StridedParts{parts, (size_t)n}, nparts, n, sqrt_norm, order);
}

// Peer variants: block p is read from peers[p] (NVLink peer memory)
void launch_krum_select_peers(const double* const* peers, int npeers, int n, int f, int32_t* order, cudaStream_t st) {
  score_select(make_peers(peers, npeers), identity_map(n), npeers, n, n - f - 1, order, nullptr, f, 0, 0, st);
}
void launch_bulyan_select_peers(const double* const* peers, int npeers, int n, int f, int m, int32_t* order, int32_t* status, cudaStream_t st) {
  score_select(make_peers(peers, npeers), identity_map(n), npeers, n, m, order, status, f, m, 1, st);
}
int launch_brute_select_peers(const double* const* peers, int npeers, int n, int f, int32_t* sel, int32_t* status, cudaStream_t st) {
  return brute_select(make_peers(peers, npeers), identity_map(n), npeers, n, f, sel, status, st);
}
void launch_rowdist_select_peers(const double* const* peers, int npeers, int n, int sqrt_norm, int32_t* order, cudaStream_t st) {
  k5_rowdist_select<PeerParts><<<1, kRowSelThreads, 0, st>>>(make_peers(peers, npeers), npeers, n, sqrt_norm, order);
}

}  // namespace bz
