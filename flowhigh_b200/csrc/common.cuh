// Shared helpers for the flowhigh_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/flowhigh_b200.h"

namespace fh {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FH_ERR_CUDA;
  }
  count_launch();
  return FH_OK;
}

#define FH_REQUIRE(cond, code, ...)        \
  do {                                     \
    if (!(cond)) {                         \
      fh::set_error(__VA_ARGS__);          \
      return (code);                       \
    }                                      \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU, as nn.GELU() / F.gelu default (transformer.py:30,95)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// 16-bit operand helpers: fp16 != 0 -> IEEE half, else bfloat16 (same storage size, same MMA rate)
__device__ __forceinline__ uint32_t pack16(float a, float b, int fp16) {
  if (fp16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack16(uint32_t v, int fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}
__device__ __forceinline__ unsigned short cvt16(float a, int fp16) {
  if (fp16) {
    const __half h = __float2half_rn(a);
    return *reinterpret_cast<const unsigned short*>(&h);
  }
  const __nv_bfloat16 h = __float2bfloat16(a);
  return *reinterpret_cast<const unsigned short*>(&h);
}

// element address of (row t, channel c) in a chunked tensor: ((c/8)*chunk_stride) + t*8 + c%8
__device__ __forceinline__ int64_t chunked_index(int64_t chunk_stride, int64_t t, int c) {
  return (int64_t)(c >> 3) * chunk_stride + t * 8 + (c & 7);
}

}  // namespace fh
