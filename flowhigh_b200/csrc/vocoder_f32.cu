// fp32 CUDA-core kernels of the BigVGAN generator (models/bigvgan/models.py:172-194) on the
// reference's [B, C, L] layout: generic tapped convolution (Conv1d with dilation and the
// polyphase form of ConvTranspose1d), the fused anti-aliased Snake/SnakeBeta activation
// (alias_free_torch/act.py:23-28) and the conv_post + tanh tail.  This is the fp32 parity path.
#include "common.cuh"

namespace {

constexpr int CT_CO = 64, CT_T = 64, CT_CI = 8;

// out[b, co, P*t+p] = alpha*(bias[co] + sum_{ci,m} w[p][co][ci][m] * x[b,ci,t+off[p][m]]) + beta_res*res + acc
__global__ void __launch_bounds__(256) conv1d_taps_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const int* __restrict__ off,
                                                          const float* __restrict__ res, float beta_res, float alpha,
                                                          int accumulate, float* __restrict__ out, int Cin, int Cout,
                                                          int L, int ntaps, int P) {
  extern __shared__ __align__(16) float csm[];
  __shared__ int soff[32];
  const int p = blockIdx.z % P, bi = blockIdx.z / P;
  const int t0 = blockIdx.x * CT_T, co0 = blockIdx.y * CT_CO;
  if (threadIdx.x < ntaps) soff[threadIdx.x] = off[p * ntaps + threadIdx.x];
  __syncthreads();
  int mn = soff[0], mx = soff[0];
  for (int m = 1; m < ntaps; ++m) {
    mn = min(mn, soff[m]);
    mx = max(mx, soff[m]);
  }
  const int span = CT_T + (mx - mn);          // x window length
  float* Xs = csm;                            // [CT_CI][span]
  float* Ws = csm + CT_CI * span;             // [CT_CI][ntaps][CT_CO]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx -> t (4 each), ty -> co (4 each)
  const float* xb = x + (size_t)bi * Cin * L;
  const float* wp = w + (size_t)p * Cout * Cin * ntaps;
  float acc[4][4] = {};
  for (int ci0 = 0; ci0 < Cin; ci0 += CT_CI) {
    for (int i = threadIdx.x; i < CT_CI * span; i += 256) {
      const int c = i / span, tt = i % span;
      const int t = t0 + mn + tt, ci = ci0 + c;
      Xs[i] = (ci < Cin && t >= 0 && t < L) ? __ldg(xb + (size_t)ci * L + t) : 0.f;
    }
    for (int i = threadIdx.x; i < CT_CO * CT_CI * ntaps; i += 256) {
      const int co = i / (CT_CI * ntaps), r = i % (CT_CI * ntaps);  // r = c*ntaps + m, contiguous in w
      const int c = r / ntaps, m = r % ntaps;
      const int cog = co0 + co, ci = ci0 + c;
      Ws[(c * ntaps + m) * CT_CO + co] =
          (cog < Cout && ci < Cin) ? __ldg(wp + ((size_t)cog * Cin + ci) * ntaps + m) : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < CT_CI; ++c) {
      for (int m = 0; m < ntaps; ++m) {
        const float* wr = Ws + (c * ntaps + m) * CT_CO + ty * 4;
        const float* xr = Xs + c * span + tx * 4 + (soff[m] - mn);
        float wv[4], xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = wr[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = xr[j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  const int Lout = L * P;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx * 4 + j;
      if (t >= L) continue;
      const size_t o = ((size_t)bi * Cout + co) * Lout + (size_t)t * P + p;
      float v = acc[i][j];
      if (bias) v += bias[co];
      v *= alpha;
      if (res) v = fmaf(beta_res, res[o], v);
      if (accumulate) v += out[o];
      out[o] = v;
    }
  }
}

// ------------------------------------------------------------------------------ anti-aliased snake
// y[q] = sum_k f[k] * s~[2q+k-5],  s[m] = u[m] + inv_b*sin^2(a*u[m]),  u[m] = 2*sum_i x~[i] f[m+5-2i]
constexpr int SN_T = 256;
__global__ void __launch_bounds__(SN_T) snake_aa_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        const float* __restrict__ a, const float* __restrict__ inv_b,
                                                        const float* __restrict__ filt, int C, int L) {
  __shared__ float xs[SN_T + 10];
  __shared__ float ss[2 * SN_T + 12];
  __shared__ float f[12];
  const int ntile = (L + SN_T - 1) / SN_T;
  const int row = blockIdx.x / ntile;  // b*C + c
  const int c = row % C;
  const int q0 = (blockIdx.x % ntile) * SN_T;
  const float* xr = x + (size_t)row * L;
  if (threadIdx.x < 12) f[threadIdx.x] = filt[threadIdx.x];
  for (int i = threadIdx.x; i < SN_T + 10; i += SN_T) {
    int t = q0 - 5 + i;
    t = min(max(t, 0), L - 1);  // replicate
    xs[i] = __ldg(xr + t);
  }
  __syncthreads();
  const float al = a[c], ib = inv_b[c];
  // s window: m in [2*q0-5, 2*q0 + 2*SN_T + 4]; clamped to [0, 2L-1] (replicate pad of the 2x signal)
  for (int i = threadIdx.x; i < 2 * SN_T + 10; i += SN_T) {  // outputs read ss[2 t + k], k < 12: indices 0 .. 2 SN_T + 9
    int m = 2 * q0 - 5 + i;
    m = min(max(m, 0), 2 * L - 1);
    const int q = m >> 1;
    float u = 0.f;
    if (m & 1) {  // odd: i' = q+d, d in -2..3, tap 6-2d
#pragma unroll
      for (int d = -2; d <= 3; ++d) {
        const int xi = min(max(q + d, 0), L - 1) - (q0 - 5);
        u = fmaf(xs[xi], f[6 - 2 * d], u);
      }
    } else {  // even: d in -3..2, tap 5-2d
#pragma unroll
      for (int d = -3; d <= 2; ++d) {
        const int xi = min(max(q + d, 0), L - 1) - (q0 - 5);
        u = fmaf(xs[xi], f[5 - 2 * d], u);
      }
    }
    u *= 2.0f;
    const float sn = sinf(u * al);
    ss[i] = u + ib * (sn * sn);
  }
  __syncthreads();
  const int q = q0 + threadIdx.x;
  if (q < L) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc = fmaf(f[k], ss[2 * threadIdx.x + k], acc);
    y[(size_t)row * L + q] = acc;
  }
}

// ------------------------------------------------------------------------------ conv_post + tanh
__global__ void convpost_tanh_kernel(const float* __restrict__ x, const float* __restrict__ w, float bias,
                                     float* __restrict__ y, int C, int L) {
  const int bi = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const float* xb = x + (size_t)bi * C * L;
  float acc = bias;
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int tt = t + j - 3;
      if (tt >= 0 && tt < L) acc = fmaf(__ldg(w + c * 7 + j), __ldg(xb + (size_t)c * L + tt), acc);
    }
  }
  y[(size_t)bi * L + t] = tanhf(acc);
}

// ------------------------------------------------------------------------------ layout helpers
__global__ void transpose_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const float* s = src + (size_t)b * R * Cc;
  float* d = dst + (size_t)b * R * Cc;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? s[(size_t)r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) d[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

__global__ void cast_f32_16_kernel(const float* __restrict__ src, unsigned short* __restrict__ dst, long long n,
                                   int fp16, unsigned int* status) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    uint32_t h[2] = {fh::pack16_guard(v.x, v.y, fp16, status), fh::pack16_guard(v.z, v.w, fp16, status)};
    *reinterpret_cast<uint2*>(dst + i) = *reinterpret_cast<uint2*>(h);
  } else {
    for (long long k = i; k < n; ++k) dst[k] = fh::cvt16_guard(src[k], fp16, status);
  }
}

// chunked fp32 [B][nchunk][rows][8] -> chunked fp16 hi + lo [B][2 nchunk][rows][8] (halos included): hi = round(x),
// lo = round(x - hi).  The operand of a convolution with duplicated weights (precision "fp16x2").
__global__ void cast_split_kernel(const float* __restrict__ src, unsigned short* __restrict__ dst, long long chunk_elems,
                                  int nchunk, long long src_batch, long long dst_batch, int fp16, unsigned int* status) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int ch = blockIdx.y, b = blockIdx.z;
  if (i >= chunk_elems) return;
  const float4 v = *reinterpret_cast<const float4*>(src + (long long)b * src_batch + (long long)ch * chunk_elems + i);
  uint32_t h[2] = {fh::pack16_guard(v.x, v.y, fp16, status), fh::pack16_guard(v.z, v.w, fp16, status)};
  const float2 h0 = fh::unpack16(h[0], fp16), h1 = fh::unpack16(h[1], fp16);
  uint32_t l[2] = {fh::pack16(v.x - h0.x, v.y - h0.y, fp16), fh::pack16(v.z - h1.x, v.w - h1.y, fp16)};
  unsigned short* o = dst + (long long)b * dst_batch + (long long)ch * chunk_elems + i;
  *reinterpret_cast<uint2*>(o) = *reinterpret_cast<uint2*>(h);
  *reinterpret_cast<uint2*>(o + (long long)nchunk * chunk_elems) = *reinterpret_cast<uint2*>(l);
}

// out = a + b (+ c) (+ d): the mean over the AMP branches of a stage (bigvgan/models.py:181-187; each branch already
// carries the 1/num_kernels factor), written as fp32 and / or as the 16-bit operand of the next upsampler.  HBM-bound.
__global__ void sum_cast_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                const float* __restrict__ d, float* __restrict__ out32, unsigned short* __restrict__ out16,
                                long long n, int fp16, unsigned int* status) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(a + i);
    if (b) {
      const float4 w = *reinterpret_cast<const float4*>(b + i);
      v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
    }
    if (c) {
      const float4 w = *reinterpret_cast<const float4*>(c + i);
      v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
    }
    if (d) {
      const float4 w = *reinterpret_cast<const float4*>(d + i);
      v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
    }
    if (out32) *reinterpret_cast<float4*>(out32 + i) = v;
    if (out16) {
      uint32_t h[2] = {fh::pack16_guard(v.x, v.y, fp16, status), fh::pack16_guard(v.z, v.w, fp16, status)};
      *reinterpret_cast<uint2*>(out16 + i) = *reinterpret_cast<uint2*>(h);
    }
  } else {
    for (long long k = i; k < n; ++k) {
      float v = a[k];
      if (b) v += b[k];
      if (c) v += c[k];
      if (d) v += d[k];
      if (out32) out32[k] = v;
      if (out16) out16[k] = fh::cvt16_guard(v, fp16, status);
    }
  }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int fh_sum_cast_f32(const float* a, const float* b, const float* c, const float* d,
                                                                      float* out32, void* out16, int64_t n, int fp16,
                                                                      void* stream) {
  if (n <= 0) return FH_OK;
  FH_REQUIRE(a != nullptr && (out32 != nullptr || out16 != nullptr), FH_ERR_BAD_SHAPE, "fh_sum_cast_f32: needs a source and an output");
  FH_REQUIRE(((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0 && ((uintptr_t)c % 16) == 0 && ((uintptr_t)d % 16) == 0 &&
                 ((uintptr_t)out32 % 16) == 0 && ((uintptr_t)out16 % 8) == 0,
             FH_ERR_BAD_ALIGN, "fh_sum_cast_f32: alignment");
  const long long nthreads = (n + 3) / 4;
  sum_cast_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, c, d, out32, (unsigned short*)out16, n,
                                                                                   fp16, fh::status_word());
  return fh::check_launch("fh_sum_cast_f32");
}

extern "C" __attribute__((visibility("default"))) int fh_conv1d_taps_f32(const float* x, const float* w, const float* bias, const int* off, const float* res,
                                  float beta_res, float alpha, int accumulate, float* out, int B, int Cin, int Cout,
                                  int L, int ntaps, int P, void* stream) {
  FH_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && L > 0 && ntaps > 0 && ntaps <= 32 && P > 0, FH_ERR_BAD_SHAPE,
             "fh_conv1d_taps_f32: bad shape (ntaps must be <= 32)");
  FH_REQUIRE((int64_t)B * P <= 65535 && (Cout + CT_CO - 1) / CT_CO <= 65535, FH_ERR_BAD_SHAPE,
             "fh_conv1d_taps_f32: B*P too large");
  // worst-case window: offsets are bounded by the caller's halo; size smem for span <= CT_T + 512
  const int max_span = CT_T + 512;
  const int smem = (CT_CI * max_span + CT_CI * ntaps * CT_CO) * (int)sizeof(float);
  static int smem_set[64] = {0};
  fh::ensure_dyn_smem(conv1d_taps_kernel, smem, smem_set);
  dim3 grid((L + CT_T - 1) / CT_T, (Cout + CT_CO - 1) / CT_CO, B * P);
  conv1d_taps_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, w, bias, off, res, beta_res, alpha, accumulate, out,
                                                                Cin, Cout, L, ntaps, P);
  return fh::check_launch("fh_conv1d_taps_f32");
}

extern "C" __attribute__((visibility("default"))) int fh_snake_aa_f32(const float* x, float* y, const float* a, const float* inv_b, const float* filt, int B,
                               int C, int L, void* stream) {
  const int64_t nblk = (int64_t)((L + SN_T - 1) / SN_T) * B * C;
  FH_REQUIRE(B > 0 && C > 0 && L > 0 && nblk <= 2147483647LL, FH_ERR_BAD_SHAPE, "fh_snake_aa_f32: bad shape");
  dim3 grid((unsigned)nblk);
  snake_aa_kernel<<<grid, SN_T, 0, (cudaStream_t)stream>>>(x, y, a, inv_b, filt, C, L);
  return fh::check_launch("fh_snake_aa_f32");
}

extern "C" __attribute__((visibility("default"))) int fh_convpost_tanh_f32(const float* x, const float* w, float bias, float* y, int B, int C, int L,
                                    void* stream) {
  FH_REQUIRE(B > 0 && C > 0 && L > 0 && B <= 65535, FH_ERR_BAD_SHAPE, "fh_convpost_tanh_f32: bad shape");
  convpost_tanh_kernel<<<dim3((L + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(x, w, bias, y, C, L);
  return fh::check_launch("fh_convpost_tanh_f32");
}

extern "C" __attribute__((visibility("default"))) int fh_transpose_f32(const float* src, float* dst, int B, int R, int Cc, void* stream) {
  FH_REQUIRE(B > 0 && R > 0 && Cc > 0 && B <= 65535 && (R + 31) / 32 <= 65535, FH_ERR_BAD_SHAPE,
             "fh_transpose_f32: bad shape");
  transpose_f32_kernel<<<dim3((Cc + 31) / 32, (R + 31) / 32, B), dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, R, Cc);
  return fh::check_launch("fh_transpose_f32");
}

extern "C" __attribute__((visibility("default"))) int fh_cast_f32_16_split(const float* src, void* dst, int64_t chunk_elems, int nchunk,
                                                                           int64_t src_batch, int64_t dst_batch, int B, int fp16,
                                                                           void* stream) {
  FH_REQUIRE(B > 0 && nchunk > 0 && chunk_elems > 0 && (chunk_elems % 8) == 0 && B <= 65535 && nchunk <= 65535, FH_ERR_BAD_SHAPE,
             "fh_cast_f32_16_split: bad shape");
  FH_REQUIRE(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0 && (src_batch % 8) == 0 && (dst_batch % 8) == 0,
             FH_ERR_BAD_ALIGN, "fh_cast_f32_16_split: alignment");
  const long long nthreads = (chunk_elems + 3) / 4;
  cast_split_kernel<<<dim3((unsigned)((nthreads + 255) / 256), nchunk, B), 256, 0, (cudaStream_t)stream>>>(
      src, (unsigned short*)dst, chunk_elems, nchunk, src_batch, dst_batch, fp16, fh::status_word());
  return fh::check_launch("fh_cast_f32_16_split");
}

extern "C" __attribute__((visibility("default"))) int fh_cast_f32_16(const float* src, void* dst, int64_t n, int fp16, void* stream) {
  if (n <= 0) return FH_OK;
  FH_REQUIRE(((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 8) == 0, FH_ERR_BAD_ALIGN, "fh_cast_f32_16: alignment");
  const long long nthreads = (n + 3) / 4;
  cast_f32_16_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (unsigned short*)dst, n,
                                                                                      fp16, fh::status_word());
  return fh::check_launch("fh_cast_f32_16");
}
