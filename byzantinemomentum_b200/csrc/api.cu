// api.cu — the C ABI of libbyzagg (include/byzagg.h): argument checks, the launch geometry
// (vector width and alignment shift shared by all rows), and the kernel chains of every rule.  No device memory is allocated here and nothing synchronises the host.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "launch.cuh"
#include "dist.cuh"

namespace bz {

static thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cached[dev & 63];
  if (c == 0) cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
  return c > 0 ? c : 148;
}

int check_launch(const char* what) {
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(BZ_ECUDA, "%s: %s", what, cudaGetErrorString(err));
  return BZ_OK;
}

Geom make_geom(const float* const* rows, int n, const void* out, const void* extra, int64_t d, int want_vec) {
  Geom g{d, d, 0, 1, 1, -1, 0};
  if (want_vec <= 1 || d < 1) return g;
  const uintptr_t bytes = (uintptr_t)want_vec * sizeof(float);
  const uintptr_t mis = (uintptr_t)rows[0] % bytes;
  bool same = ((uintptr_t)out % bytes) == mis;
  if (extra != nullptr) same = same && ((uintptr_t)extra % bytes) == mis;
  for (int r = 1; r < n && same; ++r) same = ((uintptr_t)rows[r] % bytes) == mis;
  if (!same) return g;
  g.vec = want_vec;
  g.shift = (int)(mis / sizeof(float));        // element e sits at (e + shift) modulo VEC
  g.nv = (d + g.shift + want_vec - 1) / want_vec;
  return g;
}

static int check_rows(const float* const* rows, int n, int64_t d, const void* out, const char* who) {
  if (rows == nullptr) return fail(BZ_EINVAL, "%s: rows is NULL", who);
  if (n < 1) return fail(BZ_EINVAL, "%s: n = %d, expected n >= 1", who, n);
  if (n > kMaxN) return fail(BZ_EUNSUPPORTED, "%s: n = %d exceeds BZ_MAX_N = %d", who, n, kMaxN);
  if (d < 0) return fail(BZ_EINVAL, "%s: d = %lld < 0", who, (long long)d);
  if (d > 0) {
    if (out == nullptr) return fail(BZ_EINVAL, "%s: output pointer is NULL", who);
    for (int r = 0; r < n; ++r) {
      if (rows[r] == nullptr) return fail(BZ_EINVAL, "%s: rows[%d] is NULL", who, r);
      if (((uintptr_t)rows[r] % sizeof(float)) != 0) return fail(BZ_EINVAL, "%s: rows[%d] is not 4-byte aligned", who, r);
    }
    if (((uintptr_t)out % sizeof(float)) != 0) return fail(BZ_EINVAL, "%s: output is not 4-byte aligned", who);
  }
  return BZ_OK;
}

static void fill_table(RowTable& t, const float* const* rows, int n) {
  for (int r = 0; r < n; ++r) t.p[r] = rows[r];
  for (int r = n; r < kMaxN; ++r) t.p[r] = rows[0];
}

static bool launch_median(int n, const RowTable& t, const Geom& g, float* out, cudaStream_t st) {
  return launch_median_part0(n, t, g, out, st) || launch_median_part1(n, t, g, out, st) ||
         launch_median_part2(n, t, g, out, st) || launch_median_part3(n, t, g, out, st);
}
static bool launch_sorted(int n, const RowTable& t, const Geom& g, int mode, int f, float* out, cudaStream_t st) {
  return launch_sorted_part4(n, t, g, mode, f, out, st) || launch_sorted_part5(n, t, g, mode, f, out, st) ||
         launch_sorted_part6(n, t, g, mode, f, out, st) || launch_sorted_part7(n, t, g, mode, f, out, st);
}

int run_median(const float* const* rows, int n, int64_t d, float* out, cudaStream_t st) {
  RowTable t;
  fill_table(t, rows, n);
  launch_median(n, t, make_geom(rows, n, out, nullptr, d, body_vec(n)), out, st);
  return check_launch("k1_median");
}

int run_sorted(const float* const* rows, int n, int mode, int f, int64_t d, float* out, cudaStream_t st) {
  RowTable t;
  fill_table(t, rows, n);
  const Geom g = make_geom(rows, n, out, nullptr, d, body_vec(n));
  const bool special = (mode == kModeTrmean) ? launch_trmean_special(n, f, t, g, out, st)
                     : (mode == kModePhocas) ? launch_phocas_special(n, f, t, g, out, st)
                                             : launch_meamed_special(n, f, t, g, out, st);
  if (!special) launch_sorted(n, t, g, mode, f, out, st);
  return check_launch("k1_sorted");
}

int run_average_selected(const float* const* rows, int n, const int32_t* sel, int count, int zero_init,
                         float divisor, const int32_t* status, int64_t d, float* out, cudaStream_t st, int reverse = 0) {
  RowTable t;
  fill_table(t, rows, n);
  Geom g = make_geom(rows, n, out, nullptr, d, 4);
  g.reverse = reverse;
  launch_average(t, g, sel, count, zero_init, divisor, status, out, st);
  return check_launch("k3_average");
}

static int check_f_trim(int n, int f, const char* who) {
  // executable range (the reference's check(), trmean.py:52-64, is the caller's job)
  if (f < 0 || n - 2 * f < 1) return fail(BZ_EINVAL, "%s: f = %d not executable with n = %d (need n - 2f >= 1)", who, f, n);
  return BZ_OK;
}

}  // namespace bz

using namespace bz;

extern "C" {

int bz_version(void) { return 100; }
int bz_max_n(void) { return kMaxN; }
const char* bz_last_error(void) { return g_error; }

int bz_average(const float* const* rows, int n, int64_t d, float* out, void* stream) {
  if (int rc = check_rows(rows, n, d, out, "bz_average")) return rc;
  if (d == 0) return BZ_OK;
  return run_average_selected(rows, n, nullptr, n, 1, (float)n, nullptr, d, out, (cudaStream_t)stream);
}

int bz_median(const float* const* rows, int n, int64_t d, float* out, void* stream) {
  if (int rc = check_rows(rows, n, d, out, "bz_median")) return rc;
  if (d == 0) return BZ_OK;
  return run_median(rows, n, d, out, (cudaStream_t)stream);
}

int bz_trmean(const float* const* rows, int n, int f, int64_t d, float* out, void* stream) {
  if (int rc = check_rows(rows, n, d, out, "bz_trmean")) return rc;
  if (int rc = check_f_trim(n, f, "bz_trmean")) return rc;
  if (d == 0) return BZ_OK;
  return run_sorted(rows, n, kModeTrmean, f, d, out, (cudaStream_t)stream);
}

int bz_phocas(const float* const* rows, int n, int f, int64_t d, float* out, void* stream) {
  if (int rc = check_rows(rows, n, d, out, "bz_phocas")) return rc;
  if (int rc = check_f_trim(n, f, "bz_phocas")) return rc;
  if (d == 0) return BZ_OK;
  return run_sorted(rows, n, kModePhocas, f, d, out, (cudaStream_t)stream);
}

int bz_meamed(const float* const* rows, int n, int f, int64_t d, float* out, void* stream) {
  if (int rc = check_rows(rows, n, d, out, "bz_meamed")) return rc;
  if (int rc = check_f_trim(n, f, "bz_meamed")) return rc;
  if (d == 0) return BZ_OK;
  return run_sorted(rows, n, kModeMeamed, f, d, out, (cudaStream_t)stream);
}

int bz_average_selected(const float* const* rows, int n, const int32_t* sel, int count, int zero_init,
                        double divisor, const int32_t* status, int64_t d, float* out, void* stream) {
  if (int rc = check_rows(rows, n, d, out, "bz_average_selected")) return rc;
  if (count < 1 || count > n) return fail(BZ_EINVAL, "bz_average_selected: count = %d, expected 1..%d", count, n);
  if (d == 0) return BZ_OK;
  return run_average_selected(rows, n, sel, count, zero_init, (float)divisor, status, d, out, (cudaStream_t)stream);
}

}  // extern "C"

#include "api_dist.inc"
