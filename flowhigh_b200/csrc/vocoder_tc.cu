// Memory-bound vocoder kernels of the tensor-core path, on the chunked [C/8][Lp][8] layout that
// fh_tc_conv_bf16 consumes: fused anti-aliased Snake/SnakeBeta (alias_free_torch/act.py:23-28,
// resample.py:25-33, filter.py:86-94, activations.py:48-59,107-119) and conv_post + tanh
// (bigvgan/models.py:189-192).  The residual stream stays fp32; only MMA operands are bf16.
#include "common.cuh"

namespace {

constexpr int ST = 128;  // time steps per CTA (x 8 channels)

template <bool OUT_BF16>
__global__ void __launch_bounds__(256) snake_aa_chunked_kernel(const float* __restrict__ x, void* __restrict__ y,
                                                               const float* __restrict__ a,
                                                               const float* __restrict__ inv_b,
                                                               const float* __restrict__ filt, long long batch_stride,
                                                               long long chunk_stride, int row0, int nchunk, int L) {
  __shared__ float xs[(ST + 10) * 8];
  __shared__ float ss[(2 * ST + 12) * 8];
  __shared__ float f[12];
  const int ntile = (L + ST - 1) / ST;
  int id = blockIdx.x;
  const int tile = id % ntile;
  id /= ntile;
  const int ch = id % nchunk, b = id / nchunk;
  const int q0 = tile * ST;
  const int e = threadIdx.x & 7, tr = threadIdx.x >> 3;  // channel within chunk, time row 0..31
  const float* xb = x + (long long)b * batch_stride + (long long)ch * chunk_stride;
  if (threadIdx.x < 12) f[threadIdx.x] = filt[threadIdx.x];
  for (int i = tr; i < ST + 10; i += 32) {
    const int t = min(max(q0 - 5 + i, 0), L - 1);  // replicate pad
    xs[i * 8 + e] = __ldg(xb + (long long)(row0 + t) * 8 + e);
  }
  __syncthreads();
  const float al = a[ch * 8 + e], ib = inv_b[ch * 8 + e];
  for (int i = tr; i < 2 * ST + 11; i += 32) {
    int m = 2 * q0 - 5 + i;
    m = min(max(m, 0), 2 * L - 1);
    const int q = m >> 1;
    float u = 0.f;
    if (m & 1) {
#pragma unroll
      for (int d = -2; d <= 3; ++d) {
        const int xi = min(max(q + d, 0), L - 1) - (q0 - 5);
        u = fmaf(xs[xi * 8 + e], f[6 - 2 * d], u);
      }
    } else {
#pragma unroll
      for (int d = -3; d <= 2; ++d) {
        const int xi = min(max(q + d, 0), L - 1) - (q0 - 5);
        u = fmaf(xs[xi * 8 + e], f[5 - 2 * d], u);
      }
    }
    u *= 2.0f;
    const float sn = sinf(u * al);
    ss[i * 8 + e] = u + ib * (sn * sn);
  }
  __syncthreads();
  for (int i = tr; i < ST; i += 32) {
    const int q = q0 + i;
    if (q >= L) break;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc = fmaf(f[k], ss[(2 * i + k) * 8 + e], acc);
    const long long o = (long long)b * batch_stride + (long long)ch * chunk_stride + (long long)(row0 + q) * 8 + e;
    if (OUT_BF16)
      ((__nv_bfloat16*)y)[o] = __float2bfloat16(acc);
    else
      ((float*)y)[o] = acc;
  }
}

__global__ void convpost_tanh_chunked_kernel(const float* __restrict__ x, long long batch_stride,
                                             long long chunk_stride, int row0, const float* __restrict__ w, float bias,
                                             float* __restrict__ y, int C, int L) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const float* xb = x + (long long)b * batch_stride;
  float acc = bias;
  for (int c = 0; c < C; ++c) {
    const float* xc = xb + (long long)(c >> 3) * chunk_stride + (c & 7);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int tt = t + j - 3;
      if (tt >= 0 && tt < L) acc = fmaf(__ldg(w + c * 7 + j), __ldg(xc + (long long)(row0 + tt) * 8), acc);
    }
  }
  y[(long long)b * L + t] = tanhf(acc);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int fh_snake_aa_chunked(const float* x, void* y, const float* a, const float* inv_b, const float* filt,
                                   int64_t batch_stride, int64_t chunk_stride, int row0, int B, int C, int L,
                                   int out_is_bf16, void* stream) {
  FH_REQUIRE(B > 0 && C > 0 && (C % 8) == 0 && L > 0, FH_ERR_BAD_SHAPE, "fh_snake_aa_chunked: C must be a multiple of 8");
  const long long nblk = (long long)((L + ST - 1) / ST) * (C / 8) * B;
  FH_REQUIRE(nblk <= 2147483647LL, FH_ERR_BAD_SHAPE, "fh_snake_aa_chunked: grid too large");
  if (out_is_bf16)
    snake_aa_chunked_kernel<true><<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(x, y, a, inv_b, filt, batch_stride,
                                                                                 chunk_stride, row0, C / 8, L);
  else
    snake_aa_chunked_kernel<false><<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(
        x, y, a, inv_b, filt, batch_stride, chunk_stride, row0, C / 8, L);
  return fh::check_launch("fh_snake_aa_chunked");
}

extern "C" __attribute__((visibility("default"))) int fh_convpost_tanh_chunked(const float* x, int64_t batch_stride, int64_t chunk_stride, int row0,
                                        const float* w, float bias, float* y, int B, int C, int L, void* stream) {
  FH_REQUIRE(B > 0 && C > 0 && L > 0 && B <= 65535, FH_ERR_BAD_SHAPE, "fh_convpost_tanh_chunked: bad shape");
  convpost_tanh_chunked_kernel<<<dim3((L + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(x, batch_stride, chunk_stride,
                                                                                        row0, w, bias, y, C, L);
  return fh::check_launch("fh_convpost_tanh_chunked");
}
