// fp2_peak.cu — what the FP32 pipe of a B200 SM really sustains on K2's inner block
// (25 packed accumulators; per pair and 4 coordinates: 2 x sub.f32x2 + 2 x fma.rn.f32x2), operands
// in registers, no memory traffic: the FP floor of K2.  Variants: packed (FADD2/FFMA2) vs scalar
// (FADD/FFMA), 4..16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp2_peak fp2_peak.cu && ./fp2_peak
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm volatile("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <bool PACKED>
__global__ void __launch_bounds__(512, 1) block25(const float* __restrict__ in, float* __restrict__ out, int iters) {
  u64 a0[5], a1[5], b0[5], b1[5], acc[25];
  const int t = threadIdx.x + blockIdx.x * blockDim.x;
  for (int i = 0; i < 5; ++i) {
    const ulonglong2 x = reinterpret_cast<const ulonglong2*>(in)[(t * 10 + i) & 1023];
    const ulonglong2 y = reinterpret_cast<const ulonglong2*>(in)[(t * 10 + 5 + i) & 1023];
    a0[i] = x.x; a1[i] = x.y; b0[i] = y.x; b1[i] = y.y;
  }
  for (int p = 0; p < 25; ++p) acc[p] = 0ull;
  for (int it = 0; it < iters; ++it) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const u64 d0 = sub2(a0[i], b0[j]), d1 = sub2(a1[i], b1[j]);
          acc[i * 5 + j] = fma2(d0, d0, acc[i * 5 + j]);
          acc[i * 5 + j] = fma2(d1, d1, acc[i * 5 + j]);
        }
    } else {
      float* A0 = reinterpret_cast<float*>(a0); float* A1 = reinterpret_cast<float*>(a1);
      float* B0 = reinterpret_cast<float*>(b0); float* B1 = reinterpret_cast<float*>(b1);
      float* AC = reinterpret_cast<float*>(acc);
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float d0, d1;
            asm volatile("sub.f32 %0, %1, %2;" : "=f"(d0) : "f"(A0[2 * i + h]), "f"(B0[2 * j + h]));
            asm volatile("sub.f32 %0, %1, %2;" : "=f"(d1) : "f"(A1[2 * i + h]), "f"(B1[2 * j + h]));
            asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(AC[2 * (i * 5 + j) + h]) : "f"(d0));
            asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(AC[2 * (i * 5 + j) + h]) : "f"(d1));
          }
        }
    }
  }
  float s = 0.f;
  for (int p = 0; p < 25; ++p) s += __uint_as_float((unsigned)acc[p]) + __uint_as_float((unsigned)(acc[p] >> 32));
  out[t] = s;
}

int main() {
  float *in, *out;
  cudaMalloc(&in, 1024 * 16); cudaMalloc(&out, 148 * 512 * 4);
  cudaMemset(in, 0, 1024 * 16);
  int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int packed = 1; packed >= 0; --packed)
    for (int warps = 4; warps <= 16; warps += 4) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (packed) block25<true><<<sms, warps * 32>>>(in, out, iters); else block25<false><<<sms, warps * 32>>>(in, out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      // lane-ops: per iteration and warp: 25 pairs x 4 coordinates x 2 ops x 32 lanes
      const double ops = (double)iters * warps * sms * 25 * 4 * 2 * 32;
      printf("%s warps/SM=%2d: %.3f ms  %.2f T lane-ops/s  = %.1f lane-ops/clk/SM at %.0f MHz nominal\n", packed ? "packed" : "scalar", warps, ms,
             ops / ms / 1e9, ops / (ms * 1e-3) / sms / (khz * 1e3), khz / 1e3);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
