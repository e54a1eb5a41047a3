// Memory-bound vocoder kernels of the tensor-core path, on the chunked [C/8][Lp][8] layout that
// fh_tc_conv consumes: fused anti-aliased Snake/SnakeBeta (alias_free_torch/act.py:23-28,
// resample.py:25-33, filter.py:86-94, activations.py:48-59,107-119) and conv_post + tanh
// (bigvgan/models.py:189-192).  The residual stream stays fp32; only MMA operands are 16-bit.
#include <stdlib.h>
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------ anti-aliased snake
// y[q] = sum_k f[k] s~[2q+k-5];  s[m] = u[m] + inv_b sin^2(a u[m]);  u[m] = 2 sum_i x~[i] f[m+5-2i]
// (x~ / s~ = replicate-clamped; closed form of up2x -> snake -> down2x).
//
// Register-blocked and 2-wide: one thread owns TWO adjacent channels and 17 consecutive outputs;
// every filter tap is one packed FFMA2 (fma.rn.f32x2, sm_100; measured 1.55x the scalar FFMA rate,
// tools/ffma2_bench.cu) on a channel pair, applied from registers.  sin^2(z) is evaluated as
// (1 - cos 2z)/2 so that
//     s' = u - (inv_b/2) cos(2 a u),   y = sum_k f[k] s'~[.] + inv_b/2      (sum_k f[k] = 1)
// costs one packed multiply, two MUFU.COS and one packed FMA per pair.  cos.approx's absolute
// error (~1e-6 for |arg| < 1e2) is far below the 16-bit operand rounding that follows.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

// Persistent CTAs walk (batch, chunk, time-tile) work items; the (17*32 + 10) x 8-channel fp32 input window of the NEXT item is fetched with one
// cp.async.bulk (UBLKCP) into the other shared-memory buffer while the current item is computed;
// 17 outputs per thread makes the un-padded 32-byte-row window bank-conflict free for the
// (4 pairs x 8 groups) 64-bit reads of a warp; replicate clamps are patched in shared memory and
// in registers on the first / last tile of a sequence only (block-uniform branch).
constexpr int PR = 17;
constexpr int PTT = PR * 32;   // 544 time steps per tile
constexpr int PXR = PTT + 10;  // rows per staged window

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pmbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pmbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pmbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void pbulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// BULK_OUT (16-bit outputs): the 17 x 4-byte pieces a thread produces go to a shared-memory image of the output tile
// (544 rows x 16 B, contiguous in HBM) which one cp.async.bulk store writes as whole lines, instead of 17 scattered
// half-sector global stores per thread.
template <int OUT_KIND, bool BULK_OUT>
__global__ void __launch_bounds__(128, 4) snake_aa_chunked_tma_kernel(
    const float* __restrict__ x, void* __restrict__ y, const float* __restrict__ a, const float* __restrict__ inv_b,
    const float* __restrict__ filt, long long batch_stride, long long chunk_stride, int row0, int nchunk, int L,
    int ntile, int total) {
  __shared__ __align__(128) float xs[2][PXR * 8];
  __shared__ __align__(128) unsigned char ys[BULK_OUT ? PTT * 16 : 16];
  __shared__ __align__(8) unsigned long long bars[2];
  const uint32_t bar0 = sm_u32(&bars[0]);
  if (threadIdx.x == 0) {
    pmbar_init(bar0, 1);
    pmbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int rows_per_chunk = (int)(chunk_stride >> 3);
  auto issue = [&](int item, int buf) {
    const int tile = item % ntile;
    const int rest = item / ntile;
    const int ch = rest % nchunk, b = rest / nchunk;
    const int r_first = row0 + tile * PTT - 5;  // >= row0 - 5 >= 0 (left halo)
    int nrows = rows_per_chunk - r_first;       // stay inside this chunk's rows
    nrows = nrows < PXR ? nrows : PXR;
    const float* src = x + (long long)b * batch_stride + (long long)ch * chunk_stride + (long long)r_first * 8;
    pmbar_expect_tx(bar0 + 8 * buf, (uint32_t)nrows * 32u);
    pbulk_g2s(sm_u32(&xs[buf][0]), src, (uint32_t)nrows * 32u, bar0 + 8 * buf);
  };
  float2 fu[12], fd[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float fk = __ldg(filt + k);
    fu[k] = make_float2(2.0f * fk, 2.0f * fk);
    fd[k] = make_float2(fk, fk);
  }
  const int e2 = threadIdx.x & 3, g = threadIdx.x >> 2;
  int item = blockIdx.x;
  if (item < total && threadIdx.x == 0) issue(item, 0);
  int buf = 0;
  uint32_t ph0 = 0, ph1 = 0;
  for (; item < total; item += gridDim.x) {
    const int tile = item % ntile;
    const int rest = item / ntile;
    const int ch = rest % nchunk, b = rest / nchunk;
    const int qt = tile * PTT;
    const bool edge = (tile == 0) || (qt + PTT + 5 > L);
    pmbar_wait(bar0 + 8 * buf, buf ? ph1 : ph0);
    if (buf) ph1 ^= 1; else ph0 ^= 1;
    float* xt = xs[buf];
    if (edge) {  // replicate-pad the window in shared memory: rows t < 0 <- x[0], rows t >= L <- x[L-1]
      for (int i = threadIdx.x; i < PXR * 2; i += 128) {
        const int r = i >> 1, h = i & 1;
        const int t = qt - 5 + r;
        const int tc = min(max(t, 0), L - 1);
        if (tc != t) {
          const int rc = tc - (qt - 5);
          if (rc >= 0 && rc < PXR)
            *reinterpret_cast<float4*>(&xt[r * 8 + h * 4]) = *reinterpret_cast<const float4*>(&xt[rc * 8 + h * 4]);
        }
      }
      __syncthreads();
    }
    const int c0 = ch * 8 + 2 * e2;
    const float2 al2 = make_float2(2.0f * __ldg(a + c0), 2.0f * __ldg(a + c0 + 1));
    const float2 hib = make_float2(0.5f * __ldg(inv_b + c0), 0.5f * __ldg(inv_b + c0 + 1));
    const float2 nhib = make_float2(-hib.x, -hib.y);
    const int q0 = qt + g * PR;
    float2 xv[PR + 10];
    const float* xp = xt + g * (PR * 8) + 2 * e2;
#pragma unroll
    for (int j = 0; j < PR + 10; ++j) xv[j] = *reinterpret_cast<const float2*>(xp + j * 8);
    if (BULK_OUT && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // ys is free again
    __syncthreads();  // every thread has read this buffer's window
    if (threadIdx.x == 0 && item + (int)gridDim.x < total) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(item + gridDim.x, buf ^ 1);  // that buffer was fully read one iteration ago
    }
    if (q0 < L) {
      float2 s[2 * PR + 10];
#pragma unroll
      for (int i = 0; i < 2 * PR + 10; ++i) {
        const int qq = (i - 5) >> 1;
        float2 u = make_float2(0.f, 0.f);
        if ((i & 1) == 0) {
#pragma unroll
          for (int d = -2; d <= 3; ++d) u = ffma2(xv[qq + d + 5], fu[6 - 2 * d], u);
        } else {
#pragma unroll
          for (int d = -3; d <= 2; ++d) u = ffma2(xv[qq + d + 5], fu[5 - 2 * d], u);
        }
        const float2 z = fmul2(u, al2);
        const float2 c = make_float2(__cosf(z.x), __cosf(z.y));
        s[i] = ffma2(c, nhib, u);
      }
      if (edge && (q0 == 0 || q0 + PR + 3 >= L)) {
        const int ic = 2 * (L - q0) + 5;
        float2 prev = s[5];
#pragma unroll
        for (int i = 0; i < 2 * PR + 10; ++i) {
          if (q0 == 0 && i < 5) s[i] = prev;
          if (i < ic) prev = s[i];
          else s[i] = prev;
        }
      }
      const long long obase =
          (long long)b * batch_stride + (long long)ch * chunk_stride + (long long)(row0 + q0) * 8 + 2 * e2;
#pragma unroll
      for (int j = 0; j < PR; ++j) {
        if (!edge || q0 + j < L) {
          float2 acc = hib;
#pragma unroll
          for (int k = 0; k < 12; ++k) acc = ffma2(fd[k], s[2 * j + k], acc);
          if (BULK_OUT)
            *reinterpret_cast<uint32_t*>(ys + (size_t)(g * PR + j) * 16 + 4 * e2) = fh::pack16(acc.x, acc.y, OUT_KIND == 2);
          else if (OUT_KIND)
            *reinterpret_cast<uint32_t*>((unsigned short*)y + obase + j * 8) = fh::pack16(acc.x, acc.y, OUT_KIND == 2);
          else
            *reinterpret_cast<float2*>((float*)y + obase + j * 8) = acc;
        }
      }
    }
    if (BULK_OUT) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0) {
        const int nrows = min(PTT, L - qt);
        const unsigned short* dst =
            (const unsigned short*)y + (long long)b * batch_stride + (long long)ch * chunk_stride + (long long)(row0 + qt) * 8;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(sm_u32(ys)),
                     "r"((uint32_t)nrows * 16u)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    buf ^= 1;
  }
  if (BULK_OUT && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__global__ void convpost_tanh_chunked_kernel(const float* __restrict__ x, long long batch_stride,
                                             long long chunk_stride, int row0, const float* __restrict__ w, float bias,
                                             float* __restrict__ y, int C, int L) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const float* xb = x + (long long)b * batch_stride;
  float acc = bias;
  for (int c0 = 0; c0 < C; c0 += 8) {
    const float* xc = xb + (long long)(c0 >> 3) * chunk_stride;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int tt = t + j - 3;
      if (tt < 0 || tt >= L) continue;
      const float4* p = reinterpret_cast<const float4*>(xc + (long long)(row0 + tt) * 8);
      const float4 v0 = __ldg(p), v1 = __ldg(p + 1);
      const float* wj = w + c0 * 7 + j;
      acc = fmaf(__ldg(wj), v0.x, acc);
      acc = fmaf(__ldg(wj + 7), v0.y, acc);
      acc = fmaf(__ldg(wj + 14), v0.z, acc);
      acc = fmaf(__ldg(wj + 21), v0.w, acc);
      acc = fmaf(__ldg(wj + 28), v1.x, acc);
      acc = fmaf(__ldg(wj + 35), v1.y, acc);
      acc = fmaf(__ldg(wj + 42), v1.z, acc);
      acc = fmaf(__ldg(wj + 49), v1.w, acc);
    }
  }
  y[(long long)b * L + t] = tanhf(acc);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int fh_snake_aa_chunked(
    const float* x, void* y, const float* a, const float* inv_b, const float* filt, int64_t batch_stride,
    int64_t chunk_stride, int row0, int B, int C, int L, int out_kind, void* stream) {
  FH_REQUIRE(B > 0 && C > 0 && (C % 8) == 0 && L > 0, FH_ERR_BAD_SHAPE, "fh_snake_aa_chunked: C must be a multiple of 8");
  FH_REQUIRE(out_kind >= 0 && out_kind <= 2, FH_ERR_BAD_SHAPE, "fh_snake_aa_chunked: out_kind must be 0 (fp32), 1 (bf16), 2 (fp16)");
  FH_REQUIRE(((uintptr_t)x % 16) == 0 && (batch_stride % 8) == 0 && (chunk_stride % 8) == 0, FH_ERR_BAD_ALIGN,
             "fh_snake_aa_chunked: x must be 16-byte aligned and strides multiples of 8");
  const int ntile = (L + PTT - 1) / PTT;
  const long long total = (long long)ntile * (C / 8) * B;
  FH_REQUIRE(total <= 2147483647LL && row0 >= 5, FH_ERR_BAD_SHAPE, "fh_snake_aa_chunked: needs a left halo of >= 5 rows");
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  static int per_sm = 0;  // persistent CTAs per SM (4 fill the SM; 2 leave room for a co-resident conv CTA)
  if (!per_sm) {
    const char* e = getenv("FH_SNAKE_CTAS_PER_SM");
    per_sm = e ? atoi(e) : 4;
    if (per_sm < 1 || per_sm > 4) per_sm = 4;
  }
  const int grid = (int)(total < (long long)sms * per_sm ? total : (long long)sms * per_sm);
  static int bulk = -1;
  if (bulk < 0) {
    const char* e = getenv("FH_SNAKE_BULK");
    bulk = e ? atoi(e) : 1;
  }
#define FH_SNAKE_LAUNCH(KIND, BULK)                                                                                \
  snake_aa_chunked_tma_kernel<KIND, BULK><<<grid, 128, 0, (cudaStream_t)stream>>>(x, y, a, inv_b, filt, batch_stride, \
                                                                                 chunk_stride, row0, C / 8, L, ntile, (int)total)
  if (out_kind == 0) FH_SNAKE_LAUNCH(0, false);
  else if (out_kind == 1 && bulk) FH_SNAKE_LAUNCH(1, true);
  else if (out_kind == 1) FH_SNAKE_LAUNCH(1, false);
  else if (bulk) FH_SNAKE_LAUNCH(2, true);
  else FH_SNAKE_LAUNCH(2, false);
#undef FH_SNAKE_LAUNCH
  return fh::check_launch("fh_snake_aa_chunked");
}

extern "C" __attribute__((visibility("default"))) int fh_convpost_tanh_chunked(
    const float* x, int64_t batch_stride, int64_t chunk_stride, int row0, const float* w, float bias, float* y, int B,
    int C, int L, void* stream) {
  FH_REQUIRE(B > 0 && C > 0 && (C % 8) == 0 && L > 0 && B <= 65535, FH_ERR_BAD_SHAPE,
             "fh_convpost_tanh_chunked: bad shape");
  convpost_tanh_chunked_kernel<<<dim3((L + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(x, batch_stride, chunk_stride,
                                                                                        row0, w, bias, y, C, L);
  return fh::check_launch("fh_convpost_tanh_chunked");
}
