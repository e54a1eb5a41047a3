// microbenchmark: scalar FFMA vs packed FFMA2 issue throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float* out, float a, float b, int iters) {
  float2 acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        acc[i].x = fmaf(acc[i].x, A.x, B.x);
        acc[i].y = fmaf(acc[i].y, A.y, B.y);
      } else {
        acc[i] = ffma2(acc[i], A, B);
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
      else k<1><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 8 * 256 * 16.0 * iters;
      printf("mode %d (%s): %.3f ms, %.2f TFMA/s (%.1f TFLOP/s)\n", mode, mode ? "FFMA2" : "FFMA", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
