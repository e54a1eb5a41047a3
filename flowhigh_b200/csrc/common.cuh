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
// device address of the caller's status word (fh_set_status_word, thread-local on the host; may be nullptr):
// launchers of kernels that write 16-bit operands pass it on, see Guard16 below
unsigned int* status_word();
// debug word (fh_set_debug_word; pinned HOST memory survives a device trap): a bounded mbarrier wait that times out
// writes (code | block << 8) there before trapping, so a protocol bug names the wait that hung
unsigned int* err_word();

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FH_ERR_CUDA;
  }
  count_launch();
  return FH_OK;
}

// Per-device launch state.  cudaFuncSetAttribute and the SM count belong to a DEVICE, not to the process: a second
// engine on another GPU of the same process must opt its kernels in again (indexed by cudaGetDevice, up to 64 devices).
inline int dev_index() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < 64) ? d : 0;
}
inline int dev_sms() {
  static int sms[64] = {0};
  const int d = dev_index();
  if (!sms[d]) {
    cudaDeviceGetAttribute(&sms[d], cudaDevAttrMultiProcessorCount, d);
    if (sms[d] <= 0) sms[d] = 148;
  }
  return sms[d];
}
// opts `kernel` in to `bytes` of dynamic shared memory on the current device; `table` is a static int[64] of the caller
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, int bytes, int* table) {
  const int d = dev_index();
  if (bytes <= table[d]) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) table[d] = bytes;
  return e;
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

// 16-bit operand helpers: fp16 != 0 -> IEEE half, else bfloat16 (same storage size, same MMA rate).
// The fp16 conversions SATURATE (cvt.rn.satfinite: |x| > 65504 -> +-65504, never inf) -- a saturated operand is wrong
// but finite, and Guard16 reports it, so an overflow can neither poison the rest of the tensor with NaNs nor pass
// silently.  The reference only prints on NaN (models/flow.py:256-267).
__device__ __forceinline__ uint32_t pack16(float a, float b, int fp16) {
  if (fp16) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack16(uint32_t v, int fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}
__device__ __forceinline__ unsigned short cvt16(float a, int fp16) { return (unsigned short)(pack16(a, 0.f, fp16) & 0xffffu); }

// Overflow / NaN guard of the 16-bit operand writers: a running NaN-propagating maximum of |value| over everything a
// thread converts (one HMNMX2 per packed pair), tested once per thread at the end.  fp16: >= 65504 (saturated, inf or
// NaN); bf16: inf or NaN.  A hit ORs bit 0 into the status word the host reads once per generate().
struct Guard16 {
  uint32_t m = 0u;
  __device__ __forceinline__ void see(uint32_t packed, int fp16) {
    if (fp16) {
      const __half2 r = __hmax2_nan(*reinterpret_cast<const __half2*>(&m), __habs2(*reinterpret_cast<const __half2*>(&packed)));
      m = *reinterpret_cast<const uint32_t*>(&r);
    } else {
      const __nv_bfloat162 r = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162*>(&m),
                                           __habs2(*reinterpret_cast<const __nv_bfloat162*>(&packed)));
      m = *reinterpret_cast<const uint32_t*>(&r);
    }
  }
  __device__ __forceinline__ void see1(unsigned short h, int fp16) { see((uint32_t)h, fp16); }
  __device__ __forceinline__ bool bad(int fp16) const {
    const uint32_t thr = fp16 ? 0x7bffu : 0x7f80u;
    return (m & 0xffffu) >= thr || (m >> 16) >= thr;
  }
  __device__ __forceinline__ void commit(unsigned int* status, int fp16) const {
    if (status != nullptr && bad(fp16)) atomicOr(status, 1u);
  }
};

// guarded single-element / packed conversions for the small kernels (the hot kernels batch the test through Guard16)
__device__ __forceinline__ unsigned short cvt16_guard(float a, int fp16, unsigned int* status) {
  const unsigned short h = cvt16(a, fp16);
  if (status != nullptr && (h & 0x7fffu) >= (fp16 ? 0x7bffu : 0x7f80u)) atomicOr(status, 1u);
  return h;
}
__device__ __forceinline__ uint32_t pack16_guard(float a, float b, int fp16, unsigned int* status) {
  const uint32_t h = pack16(a, b, fp16);
  if (status != nullptr) {
    const uint32_t thr = fp16 ? 0x7bffu : 0x7f80u;
    if ((h & 0x7fffu) >= thr || ((h >> 16) & 0x7fffu) >= thr) atomicOr(status, 1u);
  }
  return h;
}

// element address of (row t, channel c) in a chunked tensor: ((c/8)*chunk_stride) + t*8 + c%8
__device__ __forceinline__ int64_t chunked_index(int64_t chunk_stride, int64_t t, int c) {
  return (int64_t)(c >> 3) * chunk_stride + t * 8 + (c & 7);
}

}  // namespace fh
