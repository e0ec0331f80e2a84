/* byzoracle.c — plain C restatement of the data-parallel part of the reference's aggregation
 * rules (TEST / BENCH INFRASTRUCTURE: see oracle/byzoracle.py for who may use oracle/).
 *
 * Same arithmetic as oracle/byzoracle.py — fp32 operations in the reference's order — but
 * compiled and threaded (pthreads over coordinate ranges), so that full-size inputs are checked in
 * seconds and `bench.py` can quote a strong CPU baseline next to the reference's own ATen
 * sequence.  Follows (reference root): average.py:29, median.py:39, trmean.py:33,48-50,
 * krum.py:45 / bulyan.py:50 / brute.py:45 (distances), cge.py:36, aksel.py:41, and the
 * ordered-subset means of krum.py:80, brute.py:80, aksel.py:64, cge.py:53-56.
 * Selection logic (scores, stable sorts, subset search) stays in the NumPy oracle: it is
 * O(n^2..) host code, not data-parallel.
 *
 * Built by oracle/c/Makefile into oracle/c/libbyzoracle.so (gcc -O2 -pthread, no fast-math:
 * every fp32 operation must round exactly once).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- tiny fork/join helper (no OpenMP runtime in the image): run body(lo, hi, ctx) on
 * contiguous slices of [0, count) on ORC_THREADS threads (default: online CPUs, capped) ---- */
typedef void (*range_fn)(int64_t lo, int64_t hi, void* ctx);
typedef struct { range_fn fn; int64_t lo, hi; void* ctx; } range_job;
static void* range_trampoline(void* p) { range_job* j = (range_job*)p; j->fn(j->lo, j->hi, j->ctx); return NULL; }
static int orc_threads(void) {
  const char* env = getenv("ORC_THREADS");
  long t = env ? atol(env) : sysconf(_SC_NPROCESSORS_ONLN);
  if (t < 1) t = 1;
  if (t > 256) t = 256;
  return (int)t;
}
static void parallel_for(int64_t count, int64_t grain, range_fn fn, void* ctx) {
  int threads = orc_threads();
  if (count / (grain > 0 ? grain : 1) < threads) threads = (int)(count / (grain > 0 ? grain : 1));
  if (threads <= 1) { fn(0, count, ctx); return; }
  pthread_t tid[256];
  range_job job[256];
  int64_t per = (count + threads - 1) / threads;
  int started = 0;
  for (int t = 0; t < threads; ++t) {
    job[t].fn = fn; job[t].ctx = ctx;
    job[t].lo = t * per; job[t].hi = (t + 1) * per < count ? (t + 1) * per : count;
    if (job[t].lo >= job[t].hi) break;
    if (pthread_create(&tid[t], NULL, range_trampoline, &job[t]) != 0) { range_trampoline(&job[t]); tid[t] = 0; }
    ++started;
  }
  for (int t = 0; t < started; ++t) if (tid[t]) pthread_join(tid[t], NULL);
}

#define ORC_MAX_N 1024

static int is_nan_f(float x) { return x != x; }

/* ascending, NaN last (torch.sort / numpy.sort); insertion sort is fine for n <= 64 */
static void sort_nan_last(float* v, int n) {
  for (int i = 1; i < n; ++i) {
    float x = v[i];
    int j = i - 1;
    while (j >= 0 && (is_nan_f(v[j]) ? !is_nan_f(x) : (!is_nan_f(x) && v[j] > x))) { v[j + 1] = v[j]; --j; }
    v[j + 1] = x;
  }
}

/* ATen mean(dim=0) order over s[lo..hi): cascade with 16-row blocks, one division */
static float aten_mean(const float* s, int lo, int hi) {
  volatile float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  int cnt = 0;
  for (int k = lo; k < hi; ++k) {
    acc0 = acc0 + s[k];
    ++cnt;
    if (cnt % 16 == 0) { acc1 = acc1 + acc0; acc0 = 0.f; if (cnt % 256 == 0) { acc2 = acc2 + acc1; acc1 = 0.f; } }
  }
  volatile float total = (acc0 + acc1);
  total = total + acc2;
  total = total + 0.f;
  return total / (float)(hi - lo);
}

typedef struct { const float* const* rows; int n, f, kind, count, zero_init; int64_t d; float* out; const int32_t* sel; float divisor;
                 const float* center; double* dout; int map_nonfinite; } orc_ctx;

static void average_range(int64_t lo, int64_t hi, void* p) {
  orc_ctx* c = (orc_ctx*)p;
  for (int64_t j = lo; j < hi; ++j) {
    volatile float acc = 0.f + c->rows[0][j];
    for (int r = 1; r < c->n; ++r) acc = acc + c->rows[r][j];
    c->out[j] = acc / (float)c->n;
  }
}
void orc_average(const float* const* rows, int n, int64_t d, float* out) {
  orc_ctx c = {0}; c.rows = rows; c.n = n; c.d = d; c.out = out;
  parallel_for(d, 4096, average_range, &c);
}

static void median_range(int64_t lo, int64_t hi, void* p) {
  orc_ctx* c = (orc_ctx*)p;
  const int n = c->n;
  for (int64_t j = lo; j < hi; ++j) {
    float v[ORC_MAX_N];
    int any_nan = 0;
    for (int r = 0; r < n; ++r) { v[r] = c->rows[r][j]; any_nan |= is_nan_f(v[r]); }
    sort_nan_last(v, n);
    c->out[j] = any_nan ? NAN : v[(n - 1) / 2];
  }
}
void orc_median(const float* const* rows, int n, int64_t d, float* out) {
  orc_ctx c = {0}; c.rows = rows; c.n = n; c.d = d; c.out = out;
  parallel_for(d, 1024, median_range, &c);
}

static void trmean_range(int64_t lo, int64_t hi, void* p) {
  orc_ctx* c = (orc_ctx*)p;
  const int n = c->n;
  for (int64_t j = lo; j < hi; ++j) {
    float v[ORC_MAX_N];
    for (int r = 0; r < n; ++r) v[r] = c->rows[r][j];
    sort_nan_last(v, n);
    c->out[j] = aten_mean(v, c->f, n - c->f);
  }
}
void orc_trmean(const float* const* rows, int n, int f, int64_t d, float* out) {
  orc_ctx c = {0}; c.rows = rows; c.n = n; c.f = f; c.d = d; c.out = out;
  parallel_for(d, 1024, trmean_range, &c);
}

/* mean of the m entries closest to c (NaN keys largest, ties to the lower row), summed in
 * ascending row order — the convention of oracle/byzoracle.py::closest_mean */
static float closest_mean(const float* col, int n, int m, float c) {
  float key[ORC_MAX_N];
  int nan_key[ORC_MAX_N], order[ORC_MAX_N], chosen[ORC_MAX_N];
  for (int r = 0; r < n; ++r) {
    volatile float df = col[r] - c;
    float k = fabsf(df);
    nan_key[r] = is_nan_f(k);
    key[r] = nan_key[r] ? INFINITY : k;
    order[r] = r;
  }
  for (int i = 1; i < n; ++i) {           /* stable insertion sort by (key, nan_key) */
    int x = order[i], j = i - 1;
    while (j >= 0 && (key[order[j]] > key[x] || (key[order[j]] == key[x] && nan_key[order[j]] > nan_key[x]))) { order[j + 1] = order[j]; --j; }
    order[j + 1] = x;
  }
  memset(chosen, 0, sizeof(int) * (size_t)n);
  for (int i = 0; i < m; ++i) chosen[order[i]] = 1;
  volatile float acc = 0.f;
  int first = 1;
  for (int r = 0; r < n; ++r) if (chosen[r]) { if (first) { acc = col[r]; first = 0; } else acc = acc + col[r]; }
  return acc / (float)m;
}

/* center_kind: 0 = trimmed mean (phocas), 1 = median (meamed) */
static void closest_range(int64_t lo, int64_t hi, void* p) {
  orc_ctx* c = (orc_ctx*)p;
  const int n = c->n, f = c->f;
  for (int64_t j = lo; j < hi; ++j) {
    float col[ORC_MAX_N], v[ORC_MAX_N];
    int any_nan = 0;
    for (int r = 0; r < n; ++r) { col[r] = v[r] = c->rows[r][j]; any_nan |= is_nan_f(v[r]); }
    sort_nan_last(v, n);
    float center = c->kind == 0 ? aten_mean(v, f, n - f) : (any_nan ? NAN : v[(n - 1) / 2]);
    c->out[j] = closest_mean(col, n, n - f, center);
  }
}
void orc_closest(const float* const* rows, int n, int f, int center_kind, int64_t d, float* out) {
  orc_ctx c = {0}; c.rows = rows; c.n = n; c.f = f; c.kind = center_kind; c.d = d; c.out = out;
  parallel_for(d, 1024, closest_range, &c);
}

/* D[i*n+j] = double(fl32(sqrt(sum_k fl32(x_i - x_j)^2))) (exact sum in fp64), symmetric, 0 diagonal */
static void pairdist_range(int64_t lo, int64_t hi, void* p) {
  orc_ctx* c = (orc_ctx*)p;
  const int n = c->n;
  for (int64_t q = lo; q < hi; ++q) {
    int i = (int)(q / n), j = (int)(q % n);
    if (i >= j) continue;
    double s = 0.;
    for (int64_t k = 0; k < c->d; ++k) {
      volatile float df = c->rows[i][k] - c->rows[j][k];
      s += (double)df * (double)df;
    }
    double v = (double)(float)sqrt(s);
    if (c->map_nonfinite && !isfinite(v)) v = INFINITY;
    c->dout[i * n + j] = c->dout[j * n + i] = v;
  }
}
void orc_pairdist(const float* const* rows, int n, int64_t d, int map_nonfinite, double* D) {
  for (int i = 0; i < n; ++i) D[i * n + i] = 0.;
  orc_ctx c = {0}; c.rows = rows; c.n = n; c.d = d; c.dout = D; c.map_nonfinite = map_nonfinite;
  parallel_for((int64_t)n * n, 1, pairdist_range, &c);
}

/* center NULL: squared norms (exact squares); else fl32 squares of fl32 differences (aksel.py:41) */
static void rowdist_range(int64_t lo, int64_t hi, void* p) {
  orc_ctx* c = (orc_ctx*)p;
  for (int64_t r = lo; r < hi; ++r) {
    double s = 0.;
    if (c->center == NULL) {
      for (int64_t k = 0; k < c->d; ++k) s += (double)c->rows[r][k] * (double)c->rows[r][k];
    } else {
      for (int64_t k = 0; k < c->d; ++k) {
        volatile float df = c->rows[r][k] - c->center[k];
        volatile float sq = df * df;
        s += (double)sq;
      }
    }
    c->dout[r] = s;
  }
}
void orc_rowdist_sq(const float* const* rows, int n, const float* center, int64_t d, double* out) {
  orc_ctx c = {0}; c.rows = rows; c.n = n; c.center = center; c.d = d; c.dout = out;
  parallel_for(n, 1, rowdist_range, &c);
}

/* (((z + g[sel0]) + g[sel1]) + ...) / divisor */
static void selected_range(int64_t lo, int64_t hi, void* p) {
  orc_ctx* c = (orc_ctx*)p;
  for (int64_t j = lo; j < hi; ++j) {
    volatile float acc = c->zero_init ? 0.f + c->rows[c->sel[0]][j] : c->rows[c->sel[0]][j];
    for (int k = 1; k < c->count; ++k) acc = acc + c->rows[c->sel[k]][j];
    c->out[j] = acc / c->divisor;
  }
}
void orc_average_selected(const float* const* rows, const int32_t* sel, int count, int zero_init, float divisor,
                          int64_t d, float* out) {
  orc_ctx c = {0}; c.rows = rows; c.sel = sel; c.count = count; c.zero_init = zero_init; c.divisor = divisor; c.d = d; c.out = out;
  parallel_for(d, 4096, selected_range, &c);
}

int orc_version(void) { return 1; }
