// k7_gradient.cu — K7: the per-worker gradient production of attack.py:776-780 / 791-795 (clip,
// clone) and :799-810 (momentum placement) in ONE pass per gradient, written straight into the rows
// the aggregation rules read.
//
// The reference, per worker: `grad.norm().item()` (a pass + a host sync), `grad.mul_(clip / norm)`
// (read + write), `grad.clone()` (read + write), then for worker-side momentum
// `gmtm.mul_(mu).add_(grad, alpha=1 - dampening)` (two more read-modify-write passes) — 7 vector
// passes and a synchronisation.  Here: [clip only] one read for the squared norm (fp64 partials,
// last-CTA ticket, the scale factor stays on the device), then one kernel that reads the gradient
// (and the momentum vector) once and writes the sampled row and the momentum / honest row.
//
// Arithmetic (fp32, exactly the ATen operator sequence; `fl` = round to fp32):
//   scale  = norm > clip ? fl(clip / norm) : (no multiplication)      attack.py:777-779 (Python double division)
//   g'     = fl(g * scale)                                            grad.mul_(scalar)
//   worker: m <- fma(fl(alpha), g', fl(m * fl(mu)))                   gmtm.mul_(mu).add_(grad, alpha=1-dampening)  :802
//   server: h  = fma(fl(mu), s, fl(g' * fl(alpha)))                   grad.mul(1-dampening).add_(server, alpha=mu) :807
// (`a + alpha * b` is one fused multiply-add in ATen's CPU and CUDA add kernels alike.)
// The norm is the fp32 rounding of the square root of an fp64 sum: ATen's own fp32 reduction order
// differs between builds, so `scale` may differ from the reference's by 1 ulp when clipping occurs.
#include "dist.cuh"
#include "launch.cuh"

namespace bz {

constexpr int kGrThreads = 256;

__global__ void __launch_bounds__(kGrThreads)
k7_sqnorm(const float* __restrict__ g, const int64_t d, const double clip, double* __restrict__ parts,
          unsigned* __restrict__ ticket, float* __restrict__ scale_out) {
  __shared__ double warp_sum[kGrThreads / 32];
  __shared__ bool last;
  double s = 0.;
  for (int64_t i = (int64_t)blockIdx.x * kGrThreads + threadIdx.x; i < d; i += (int64_t)gridDim.x * kGrThreads) {
    const double x = (double)__ldg(g + i);
    s = fma(x, x, s);
  }
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) s += __shfl_xor_sync(0xffffffffu, s, h);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.;
#pragma unroll
    for (int w = 0; w < kGrThreads / 32; ++w) t += warp_sum[w];
    parts[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  double total = 0.;
  for (int p = 0; p < (int)gridDim.x; ++p) total += __ldcg(parts + p);      // index order: deterministic
  const float norm = (float)sqrt(total);                                    // Tensor.norm() is an fp32 value
  // attack.py:778-779: `if grad_norm > clip: grad.mul_(clip / grad_norm)` — the quotient is a Python double
  // handed to mul_ as a scalar, i.e. rounded to fp32; NaN norms compare false: no scaling
  *scale_out = ((double)norm > clip) ? (float)(clip / (double)norm) : 1.f;
  *ticket = 0u;
}

// mode 0: sampled only; 1: worker-side momentum (m in/out); 2: server-side momentum (m = server vector, read only)
template <int MODE>
__global__ void __launch_bounds__(kGrThreads)
k7_produce(const float* __restrict__ g, const int64_t d, const float* __restrict__ scale_ptr, float* __restrict__ sampled,
           float* __restrict__ m, const float mu, const float alpha, float* __restrict__ honest) {
  const int64_t i = (int64_t)blockIdx.x * kGrThreads + threadIdx.x;
  if (i >= d) return;
  pdl_wait();
  float x = __ldcs(g + i);
  if (scale_ptr != nullptr) {
    const float sc = *scale_ptr;
    if (sc != 1.f) x = __fmul_rn(x, sc);           // the reference does not multiply at all when it does not clip
  }
  if (sampled != nullptr) __stcs(sampled + i, x);
  if (MODE == 1) {
    const float t = __fmul_rn(m[i], mu);
    m[i] = __fmaf_rn(alpha, x, t);
  } else if (MODE == 2) {
    const float t = __fmul_rn(x, alpha);
    honest[i] = __fmaf_rn(mu, __ldg(m + i), t);
  }
}

}  // namespace bz

using namespace bz;

extern "C" int bz_gradient_row(const float* grad, int64_t d, double clip, float* sampled, int mode, float* momentum,
                               double mu, double alpha, float* honest, void* ws, size_t ws_bytes, void* stream) {
  if (grad == nullptr || d < 0) return fail(BZ_EINVAL, "bz_gradient_row: grad is NULL or d < 0");
  if (mode < 0 || mode > 2) return fail(BZ_EINVAL, "bz_gradient_row: mode = %d", mode);
  if (mode != 0 && momentum == nullptr) return fail(BZ_EINVAL, "bz_gradient_row: momentum is NULL");
  if (mode == 2 && honest == nullptr) return fail(BZ_EINVAL, "bz_gradient_row: honest is NULL");
  if (mode == 0 && sampled == nullptr) return fail(BZ_EINVAL, "bz_gradient_row: nothing to write");
  if (d == 0) return BZ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float* scale = nullptr;
  if (clip > 0.) {
    Workspace w;
    if (!carve_workspace(ws, ws_bytes, 1, w)) return fail(BZ_EWORKSPACE, "bz_gradient_row: workspace too small or misaligned");
    int64_t grid = (d + (int64_t)kGrThreads * 8 - 1) / ((int64_t)kGrThreads * 8);
    const int64_t cap = (int64_t)sm_count() * 4;
    if (grid > cap) grid = cap;
    if (grid > kMaxParts) grid = kMaxParts;
    if (grid < 1) grid = 1;
    float* scale_dev = reinterpret_cast<float*>(w.status);
    cudaMemsetAsync(w.ticket, 0, sizeof(unsigned), st);
    k7_sqnorm<<<(unsigned)grid, kGrThreads, 0, st>>>(grad, d, clip, w.parts, w.ticket, scale_dev);
    if (int rc = check_launch("k7_sqnorm")) return rc;
    scale = scale_dev;
  }
  const unsigned blocks = (unsigned)((d + kGrThreads - 1) / kGrThreads);
  const float fmu = (float)mu, falpha = (float)alpha;
  cudaError_t err;
  if (mode == 1)      err = launch_after(k7_produce<1>, blocks, kGrThreads, 0, st, grad, d, scale, sampled, momentum, fmu, falpha, honest);
  else if (mode == 2) err = launch_after(k7_produce<2>, blocks, kGrThreads, 0, st, grad, d, scale, sampled, momentum, fmu, falpha, honest);
  else                err = launch_after(k7_produce<0>, blocks, kGrThreads, 0, st, grad, d, scale, sampled, momentum, fmu, falpha, honest);
  (void)err;
  return check_launch("k7_produce");
}
