// tcgen05 / TMEM implicit-GEMM tapped convolution for sm_100a.
//
// One kernel serves every GEMM-shaped op of the hot path: the BigVGAN Conv1d / dilated Conv1d
// (bigvgan/models.py:27-43), the ConvTranspose1d upsamplers in polyphase form (models.py:140-146)
// and the backbone Linear layers (k = 1; flow.py:239,261, attend.py:176,189, transformer.py:100-103).
//
// Data layout ("chunked"): an activation [L, C] is stored as [C/8][Lp][8] bf16, so the 8 channels
// of one time step are one 16-byte row and 8 consecutive time steps x 8 channels are one
// contiguous 128-byte tcgen05 core matrix (no-swizzle, K-major canonical layout
// ((8,m),(8,2)):((16B,SBO=128B),(1,LBO))).  Consequences:
//   * a (128 + halo)-row window of one chunk is ONE cp.async.bulk (UBLKCP) of contiguous bytes;
//   * every convolution tap reads the SAME shared-memory window through a descriptor whose start
//     address is shifted by tap_offset*16 B -- the activation tile is fetched once per k taps;
//   * the two 8-channel halves of a K=16 MMA step are addressed by LBO, so no im2col, no swizzle.
// Weights are pre-packed on the host into the exact shared-memory image (packing.py), one bulk
// copy per stage.  Accumulators live in TMEM (2 x bn columns, double buffered across tiles);
// the epilogue (bias, alpha, residual, accumulate, GEGLU, bf16/fp32, chunked or row-major output)
// reads them with tcgen05.ld.  Warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer +
// TMEM allocator, warps 2..5 = epilogue.  Persistent CTAs, static tile striding.
#include <stdlib.h>
#include "common.cuh"
#include "snake_worker.cuh"

namespace {

constexpr int kMaxTapOff = 64;
constexpr int kMaxSpan = 64;                   // max (max_off - min_off) of the taps of one phase
constexpr int kMaxStages = 8;
constexpr int kThreads = 320;                  // 10 warps: producer, MMA, 8 epilogue

struct TcParams {
  const __nv_bfloat16* a;
  const __nv_bfloat16* w;
  const float* bias;
  const void* res;
  void* out;
  long long a_batch, a_chunk;
  long long out_batch, out_chunk, out_row;
  long long res_batch, res_chunk, res_row;
  int a_row0;
  int out_is_16, res_is_16, fp16, accumulate, geglu;
  const float* acc_src;  // accumulate source when it is not the output itself (fp32, output geometry); else nullptr
  int act_gelu;        // exact-erf GELU on (acc + bias) * alpha before the residual (ConvNeXt pwconv1)
  float alpha, beta_res;
  int B, L, Cin, Cout, ntaps, P, bn;
  int m_tiles, n_tiles, total_tiles;
  int ci_pairs;        // ceil(Cin / 16)
  int ci_odd;          // Cin % 16 == 8: the last pair has one real 8-channel chunk; its partner is a zeroed smem window
  int tg, n_groups;    // taps per smem stage, groups per ci-pair
  int kc;              // ci-pairs per smem stage (> 1 only when a stage holds all taps: few-tap convs, Linear)
  int msub;            // 128-row sub-tiles per CTA tile (1, 2, 4 or 8): B operand reuse + epilogue MLP
  int acc_stages;      // TMEM accumulator stages: 2 (epilogue overlaps the next tile) or 1 (msub*bn > 256)
  int wrows;           // rows fetched per chunk window: 128*msub + (max_off - min_off)
  int arows_pad;       // rows reserved per chunk window in an A slot
  int stages, stage_bytes;
  int min_off[16];
  int tap_rel0[16];     // (offset of tap 0) - min_off of the phase
  int tap_step[16];     // offset(tap j+1) - offset(tap j) when the taps of a phase form an arithmetic sequence
  int fast_epi;         // epilogue_fast applies (bias table in shared memory behind the stages)
  int v8;               // fp32 output / residual rows are 32-byte aligned: 256-bit epilogue accesses
  int tap_arith;        // all phases arithmetic: the MMA issuer strides descriptors instead of reading the offset table
  int tap_off[kMaxTapOff];
  unsigned int* err_flag;
  unsigned int* status;  // overflow / NaN status word of the caller (common.cuh Guard16) or nullptr
  // fused anti-aliased snake A-producer (tc_conv_snake_kernel): raw fp32 activations + per-channel parameters
  const float* xf;
  const float* sn_a;
  const float* sn_ib;
  const float* sn_filt;
  int xrows, x_stages, xs_bytes, rows_per_chunk;
};

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug becomes a trap (reported as a launch failure), not a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* err_flag, int code) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, (unsigned)code);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=0 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1,
// K-major A/B, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n, int fp16) {
  const uint32_t fmt = fp16 ? 0u : 1u;  // F16F32Format: F16 = 0, BF16 = 1
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// tcgen05.mma with the two shared-memory descriptors passed as (lo, hi) words: only the 14-bit start-address field of
// the low word changes between taps / sub-tiles / stages, so the issue loop advances plain 32-bit values.
__device__ __forceinline__ void umma_f16_split(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 256-bit global accesses (sm_100: LDG/STG.256): one 8-channel fp32 row of a chunk is one full 32-byte sector per
// lane instead of two half-sector requests
__device__ __forceinline__ void ldg_v8(const float* p, float (&r)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg_v8(float* p, const float (&r)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]),
               "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}

struct TileCoord {
  int b, p, mt, nt;
};
__device__ __forceinline__ TileCoord decode_tile(const TcParams& P, int id) {
  TileCoord c;
  c.nt = id % P.n_tiles;
  id /= P.n_tiles;
  c.mt = id % P.m_tiles;
  id /= P.m_tiles;
  c.p = id % P.P;
  c.b = id / P.P;
  return c;
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// ------------------------------------------------------------------------------ the kernel
// residual rows for one 16-column group (two 8-channel chunks) of one output row
__device__ __forceinline__ void load_res16(const TcParams& P, long long rbase, int n0, bool ok, float (&r)[16]) {
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int n = n0 + hh * 8;
    if (!ok || n >= P.Cout) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[hh * 8 + i] = 0.f;
      continue;
    }
    const long long ridx = rbase + (long long)(n >> 3) * P.res_chunk;
    if (P.res_is_16) {
      const uint4 raw = *reinterpret_cast<const uint4*>((const unsigned short*)P.res + ridx);
      const uint32_t* h = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = fh::unpack16(h[i], P.fp16);
        r[hh * 8 + 2 * i] = f.x;
        r[hh * 8 + 2 * i + 1] = f.y;
      }
    } else if (P.v8) {
      float t[8];
      ldg_v8((const float*)P.res + ridx, t);
#pragma unroll
      for (int i = 0; i < 8; ++i) r[hh * 8 + i] = t[i];
    } else {
      const float4* rp = reinterpret_cast<const float4*>((const float*)P.res + ridx);
      const float4 r0 = rp[0], r1 = rp[1];
      r[hh * 8 + 0] = r0.x, r[hh * 8 + 1] = r0.y, r[hh * 8 + 2] = r0.z, r[hh * 8 + 3] = r0.w;
      r[hh * 8 + 4] = r1.x, r[hh * 8 + 5] = r1.y, r[hh * 8 + 6] = r1.z, r[hh * 8 + 7] = r1.w;
    }
  }
}

// Epilogue role, shared by both kernels.  `gstep` warps share one TMEM lane group (a warp may only touch lanes
// 32*(warp%4)..+31) and split the 16-column groups of a tile round-robin (`half` = index within the share).
// Per warp the groups are software-pipelined: tcgen05.ld and the residual loads of the next group are in flight
// while the current group is scaled, added and stored.
__device__ __forceinline__ void epilogue_role(const TcParams& P, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                              int lane_grp, int half, int gstep, int lane, int tile_rows,
                                              uint32_t acc_cols) {
  const int r = lane_grp * 32 + lane;
  int as = 0, aphase = 0;
  const int groups_per_sub = P.bn >> 4;
  const int n_groups_total = P.msub * groups_per_sub;
  const bool use_res = P.res != nullptr && !P.geglu;
  fh::Guard16 guard;
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const TileCoord tc = decode_tile(P, tile);
    const int n_base = tc.nt * P.bn;
    const int t_base = tc.mt * tile_rows + r;
    const long long res_b = (long long)tc.b * P.res_batch;
    const long long out_b = (long long)tc.b * P.out_batch;
    const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)as * acc_cols;

    auto group_coords = [&](int gi, int& sub, int& c0, int& t) {
      sub = gi / groups_per_sub;
      c0 = (gi - sub * groups_per_sub) << 4;
      t = t_base + sub * 128;
    };
    auto fetch_res = [&](int gi, float (&rr)[16]) {
      if (!use_res || gi >= n_groups_total) return;
      int sub, c0, t;
      group_coords(gi, sub, c0, t);
      load_res16(P, res_b + ((long long)t * P.P + tc.p) * P.res_row, n_base + c0, t < P.L, rr);
    };
    auto issue_ld = [&](int gi, uint32_t (&v)[16]) {
      if (gi >= n_groups_total) return;
      int sub, c0, t;
      group_coords(gi, sub, c0, t);
      if (n_base + c0 < P.Cout) tmem_ld16(taddr + (uint32_t)(sub * P.bn + c0), v);  // warp-uniform
    };
    auto finish = [&](int gi, const uint32_t (&v)[16], const float (&rr)[16]) {
      int sub, c0, t;
      group_coords(gi, sub, c0, t);
      if (n_base + c0 >= P.Cout || t >= P.L) return;
      const long long orow = (long long)t * P.P + tc.p;
      if (P.geglu) {
        // columns (2i, 2i+1) = (x_i, gate_i) -> gelu(gate) * x ; 16 columns -> one 8-channel chunk
        const int n_out = (n_base + c0) >> 1;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = n_base + c0 + 2 * i;
          float xv = __uint_as_float(v[2 * i]), gv = __uint_as_float(v[2 * i + 1]);
          if (P.bias) {
            xv += __ldg(P.bias + n);
            gv += __ldg(P.bias + n + 1);
          }
          o[i] = gelu_f(gv) * xv;
        }
        const long long idx = out_b + (long long)(n_out >> 3) * P.out_chunk + orow * P.out_row;
        if (P.out_is_16) {
          uint32_t h[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            h[i] = fh::pack16(o[2 * i], o[2 * i + 1], P.fp16);
            guard.see(h[i], P.fp16);
          }
          *reinterpret_cast<uint4*>((unsigned short*)P.out + idx) = *reinterpret_cast<uint4*>(h);
        } else {
          float4* dst = reinterpret_cast<float4*>((float*)P.out + idx);
          dst[0] = make_float4(o[0], o[1], o[2], o[3]);
          dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
        return;
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int n0 = n_base + c0 + hh * 8;
        if (n0 >= P.Cout) break;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float acc = __uint_as_float(v[hh * 8 + i]);
          if (P.bias) acc += __ldg(P.bias + n0 + i);
          o[i] = acc * P.alpha;
          if (P.act_gelu) o[i] = gelu_f(o[i]);
          if (use_res) o[i] = fmaf(P.beta_res, rr[hh * 8 + i], o[i]);
        }
        const long long idx = out_b + (long long)(n0 >> 3) * P.out_chunk + orow * P.out_row;
        if (P.out_is_16) {
          uint32_t h[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            h[i] = fh::pack16(o[2 * i], o[2 * i + 1], P.fp16);
            guard.see(h[i], P.fp16);
          }
          *reinterpret_cast<uint4*>((unsigned short*)P.out + idx) = *reinterpret_cast<uint4*>(h);
        } else if (P.v8) {
          float* dst = (float*)P.out + idx;
          if (P.accumulate) {
            float t[8];
            ldg_v8(dst, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += t[i];
          }
          stg_v8(dst, o);
        } else {
          float4* dst = reinterpret_cast<float4*>((float*)P.out + idx);
          if (P.accumulate) {
            const float4 p0 = dst[0], p1 = dst[1];
            o[0] += p0.x, o[1] += p0.y, o[2] += p0.z, o[3] += p0.w;
            o[4] += p1.x, o[5] += p1.y, o[6] += p1.z, o[7] += p1.w;
          }
          dst[0] = make_float4(o[0], o[1], o[2], o[3]);
          dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    };

    uint32_t va[16], vb[16];
    float ra[16], rb[16];
    fetch_res(half, ra);  // residual of the first group is requested BEFORE waiting for the accumulator
    mbar_wait(tfull0 + 8 * as, aphase, P.err_flag, 4);
    tc_fence_after();
    issue_ld(half, va);
    for (int gi = half; gi < n_groups_total; gi += 2 * gstep) {
      tmem_ld_wait();
      issue_ld(gi + gstep, vb);
      fetch_res(gi + gstep, rb);
      finish(gi, va, ra);
      if (gi + gstep >= n_groups_total) break;
      tmem_ld_wait();
      issue_ld(gi + 2 * gstep, va);
      fetch_res(gi + 2 * gstep, ra);
      finish(gi + gstep, vb, rb);
    }
    // all TMEM reads of this warp are complete (wait::ld above) -> release the accumulator
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty0 + 8 * as);
    if (++as == P.acc_stages) {
      as = 0;
      aphase ^= 1;
    }
  }
  if (P.out_is_16) guard.commit(P.status, P.fp16);
}

// Specialised epilogue for the hot vocoder / backbone shapes: fp32 output and residual in 32-byte-aligned rows, no GEGLU.
// The generic role above spends ~210 instructions per 16-column group (per-element __ldg of the bias, runtime flag tests);
// ncu showed the eight epilogue warps -- not HBM -- pacing the HBM-bound C <= 96 stages.  Here the bias (pre-multiplied by
// alpha) comes from a shared-memory table with broadcast 128-bit loads, the flags are template parameters, and a group
// costs ~45 instructions:  o = fma(acc, alpha, alpha * bias) [+ beta * residual] [+ previous output].
// OUT16: 16-bit output rows (16 bytes per chunk row, no residual / accumulate) -- the snake that follows a first AMP
// convolution reads them as MMA operands directly.
template <bool HAS_RES, bool ACCUM, bool OUT16 = false>
__device__ __forceinline__ void epilogue_fast(const TcParams& P, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                              int lane_grp, int half, int gstep, int lane, int tile_rows, uint32_t acc_cols,
                                              const float* s_bias) {
  const int r = lane_grp * 32 + lane;
  int as = 0, aphase = 0;
  const int groups_per_sub = P.bn >> 4;
  const int n_groups_total = P.msub * groups_per_sub;
  const float alpha = P.alpha, beta = P.beta_res;
  const int bmask = P.bias ? ~0 : 0;  // no bias: every group reads the 16 zeros at the head of the table
  const float* resp = (const float*)P.res;
  float* outp = (float*)P.out;
  const float* accp = P.acc_src ? P.acc_src : (const float*)P.out;  // "previous output" rows of the accumulate form
  fh::Guard16 guard;
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const TileCoord tc = decode_tile(P, tile);
    const int n_base = tc.nt * P.bn;
    const int t_base = tc.mt * tile_rows + r;
    const long long res_b = (long long)tc.b * P.res_batch + (long long)tc.p * P.res_row;
    const long long out_b = (long long)tc.b * P.out_batch + (long long)tc.p * P.out_row;
    const long long res_rs = (long long)P.P * P.res_row, out_rs = (long long)P.P * P.out_row;
    const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)as * acc_cols;

    auto coords = [&](int gi, int& c0, int& t) {
      const int sub = gi / groups_per_sub;
      c0 = (gi - sub * groups_per_sub) << 4;
      t = t_base + sub * 128;
      return (uint32_t)(sub * P.bn + c0);
    };
    auto prefetch = [&](int gi, float (&rr)[16], float (&pp)[16]) {
      if (gi >= n_groups_total) return;
      int c0, t;
      coords(gi, c0, t);
      const int n0 = n_base + c0;
      if (n0 >= P.Cout || t >= P.L) return;
      const bool two = n0 + 8 < P.Cout;
      if (HAS_RES) {
        const float* q = resp + res_b + (long long)t * res_rs + (long long)(n0 >> 3) * P.res_chunk;
        ldg_v8(q, *reinterpret_cast<float(*)[8]>(&rr[0]));
        if (two) ldg_v8(q + P.res_chunk, *reinterpret_cast<float(*)[8]>(&rr[8]));
      }
      if (ACCUM) {
        const float* q = accp + out_b + (long long)t * out_rs + (long long)(n0 >> 3) * P.out_chunk;
        ldg_v8(q, *reinterpret_cast<float(*)[8]>(&pp[0]));
        if (two) ldg_v8(q + P.out_chunk, *reinterpret_cast<float(*)[8]>(&pp[8]));
      }
    };
    auto issue_ld = [&](int gi, uint32_t (&v)[16]) {
      if (gi >= n_groups_total) return;
      int c0, t;
      const uint32_t col = coords(gi, c0, t);
      if (n_base + c0 < P.Cout) tmem_ld16(taddr + col, v);  // warp-uniform
    };
    auto finish = [&](int gi, const uint32_t (&v)[16], const float (&rr)[16], const float (&pp)[16]) {
      int c0, t;
      coords(gi, c0, t);
      const int n0 = n_base + c0;
      if (n0 >= P.Cout || t >= P.L) return;
      const long long oidx = out_b + (long long)t * out_rs + (long long)(n0 >> 3) * P.out_chunk;
      float* dst = outp + oidx;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (hh == 1 && n0 + 8 >= P.Cout) break;
        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + ((n0 + hh * 8) & bmask));
        const float4 b1 = *reinterpret_cast<const float4*>(s_bias + ((n0 + hh * 8) & bmask) + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i] = fmaf(__uint_as_float(v[hh * 8 + i]), alpha, bb[i]);
          if (HAS_RES) o[i] = fmaf(beta, rr[hh * 8 + i], o[i]);
          if (ACCUM) o[i] += pp[hh * 8 + i];
        }
        if (OUT16) {
          uint32_t h[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            h[i] = fh::pack16(o[2 * i], o[2 * i + 1], P.fp16);
            guard.see(h[i], P.fp16);
          }
          *reinterpret_cast<uint4*>((unsigned short*)P.out + oidx + (long long)hh * P.out_chunk) = *reinterpret_cast<uint4*>(h);
        } else {
          stg_v8(dst + (long long)hh * P.out_chunk, o);
        }
      }
    };

    uint32_t va[16], vb[16];
    float ra[16], rb[16], pa[16], pb[16];
    prefetch(half, ra, pa);  // requested BEFORE waiting for the accumulator
    mbar_wait(tfull0 + 8 * as, aphase, P.err_flag, 4);
    tc_fence_after();
    issue_ld(half, va);
    for (int gi = half; gi < n_groups_total; gi += 2 * gstep) {
      tmem_ld_wait();
      issue_ld(gi + gstep, vb);
      prefetch(gi + gstep, rb, pb);
      finish(gi, va, ra, pa);
      if (gi + gstep >= n_groups_total) break;
      tmem_ld_wait();
      issue_ld(gi + 2 * gstep, va);
      prefetch(gi + 2 * gstep, ra, pa);
      finish(gi + gstep, vb, rb, pb);
    }
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty0 + 8 * as);
    if (++as == P.acc_stages) {
      as = 0;
      aphase ^= 1;
    }
  }
  if (OUT16) guard.commit(P.status, P.fp16);
}

// MMA issuer role, shared by both kernels.  The ncu source page of the first version showed this warp, not the tensor
// pipe, pacing every shape but the widest (~300 cycles of uniform-datapath work per MMA: an integer modulo per stage,
// a shared-memory load + R2UR per tap, 64-bit descriptor arithmetic per MMA).  Here every per-stage and per-tap
// quantity is a running 32-bit value: stage base, tap stride (taps of a phase are an arithmetic sequence for Conv1d,
// dilated Conv1d and polyphase ConvTranspose1d), weight stride; the sub-tile loop is specialised outside the tap loop.
template <int MSUB>
__device__ __forceinline__ void issue_taps(uint32_t d_tmem, uint32_t bn, uint32_t a_lo, uint32_t a_step, uint32_t b_lo,
                                           uint32_t b_step, uint32_t hi, uint32_t idesc, uint32_t accum, int ntap) {
#pragma unroll 1
  for (int j = 0; j < ntap; ++j) {
#pragma unroll
    for (int sub = 0; sub < MSUB; ++sub)
      umma_f16_split(d_tmem + (uint32_t)sub * bn, a_lo + (uint32_t)sub * 128u, hi, b_lo, hi, idesc, accum);
    accum = 1;
    a_lo += a_step;
    b_lo += b_step;
  }
}

__device__ __forceinline__ void mma_role(const TcParams& P, uint32_t tmem_base, uint32_t full0, uint32_t empty0,
                                         uint32_t tfull0, uint32_t tempty0, uint32_t stage0, uint32_t a_chunk_bytes,
                                         const int* s_off) {
  const bool leader = elect_one();
  int stage = 0, phase = 0, as = 0, aphase = 0;
  const int S = P.stages, msub = P.msub, ntaps = P.ntaps, tg = P.tg, n_groups = P.n_groups, ci_pairs = P.ci_pairs;
  const int kc = P.kc;
  const uint32_t bn = (uint32_t)P.bn;
  const uint32_t idesc = make_idesc(P.bn, P.fp16);
  const uint32_t hi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1 (bit 46)
  const uint32_t a_lbo = (a_chunk_bytes >> 4) << 16;            // LBO fields (bits 16..29 of the low word)
  const uint32_t b_lbo = ((bn * 16u) >> 4) << 16;
  const uint32_t b_step = (bn * 32u) >> 4;                      // one tap of weights
  const uint32_t a_slot_u = (2u * a_chunk_bytes) >> 4;
  const uint32_t stage_u = (uint32_t)P.stage_bytes >> 4, stage0_u = stage0 >> 4;
  const uint32_t acc_cols = (uint32_t)(P.msub * P.bn);
  const bool one_phase = P.P == 1;
  uint32_t rel0 = (uint32_t)P.tap_rel0[0], a_step = (uint32_t)P.tap_step[0];
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    int ph = 0;
    if (!one_phase) {
      ph = decode_tile(P, tile).p;
      rel0 = (uint32_t)P.tap_rel0[ph];
      a_step = (uint32_t)P.tap_step[ph];
    }
    mbar_wait(tempty0 + 8 * as, aphase ^ 1, P.err_flag, 2);
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + (uint32_t)as * acc_cols;
    uint32_t accum = 0;
    for (int cp = 0; cp < ci_pairs; cp += kc) {
      const int nkc = min(kc, ci_pairs - cp);
      int tap0 = 0;
      for (int g = 0; g < n_groups; ++g, tap0 += tg) {
        const int nt_g = min(tg, ntaps - tap0);
        mbar_wait(full0 + 8 * stage, phase, P.err_flag, 3);
        tc_fence_after();
        const uint32_t sa_u = stage0_u + (uint32_t)stage * stage_u;
        if (leader) {
          uint32_t b_lo = b_lbo | (sa_u + (uint32_t)kc * a_slot_u);  // weights sit behind the kc activation slots
          uint32_t a_base_lo = a_lbo | sa_u;
          for (int c = 0; c < nkc; ++c) {
            if (P.tap_arith) {
              const uint32_t a_lo = a_base_lo + rel0 + (uint32_t)tap0 * a_step;
              if (msub == 2) issue_taps<2>(d_tmem, bn, a_lo, a_step, b_lo, b_step, hi, idesc, accum, nt_g);
              else if (msub == 4) issue_taps<4>(d_tmem, bn, a_lo, a_step, b_lo, b_step, hi, idesc, accum, nt_g);
              else if (msub == 8) issue_taps<8>(d_tmem, bn, a_lo, a_step, b_lo, b_step, hi, idesc, accum, nt_g);
              else issue_taps<1>(d_tmem, bn, a_lo, a_step, b_lo, b_step, hi, idesc, accum, nt_g);
            } else {  // irregular tap offsets: table in shared memory
              const int* offs = s_off + ph * ntaps + tap0;
              uint32_t bl = b_lo, acc = accum;
              for (int j = 0; j < nt_g; ++j) {
                const uint32_t a_lo = a_base_lo + (uint32_t)offs[j];
                for (int sub = 0; sub < msub; ++sub)
                  umma_f16_split(d_tmem + (uint32_t)sub * bn, a_lo + (uint32_t)sub * 128u, hi, bl, hi, idesc, acc);
                acc = 1;
                bl += b_step;
              }
            }
            accum = 1;
            a_base_lo += a_slot_u;
            b_lo += (uint32_t)nt_g * b_step;
          }
          umma_commit(empty0 + 8 * stage);  // frees the smem stage when these MMAs retire
        }
        accum = 1;
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if (leader) umma_commit(tfull0 + 8 * as);  // accumulator complete -> epilogue
    __syncwarp();
    if (++as == P.acc_stages) {
      as = 0;
      aphase ^= 1;
    }
  }
}

// Bulk-copy producer role (one elected lane): per stage, the (rows + span)-row window of each 8-channel chunk and the
// packed weights of the stage's taps, all signalled on the stage's full barrier.
__device__ __forceinline__ void producer_role(const TcParams& P, uint32_t full0, uint32_t empty0, uint32_t stage0,
                                              uint32_t a_chunk_bytes, int tile_rows) {
  const int S = P.stages;
  const uint32_t b_tap_bytes = (uint32_t)P.bn * 32u;
  const uint32_t a_slot_bytes = 2u * a_chunk_bytes;
  int stage = 0, phase = 0;
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const TileCoord tc = decode_tile(P, tile);
    const long long row = (long long)P.a_row0 + (long long)tc.mt * tile_rows + P.min_off[tc.p];
    const __nv_bfloat16* a_base = P.a + (long long)tc.b * P.a_batch + row * 8;
    // packed weights: [p][nt][cp][tap][2][bn][8]
    const __nv_bfloat16* w_base =
        P.w + ((long long)(tc.p * P.n_tiles + tc.nt) * P.ci_pairs) * ((long long)P.ntaps * P.bn * 16);
    for (int cp = 0; cp < P.ci_pairs; cp += P.kc) {
      const int nkc = min(P.kc, P.ci_pairs - cp);
      for (int g = 0; g < P.n_groups; ++g) {
        const int tap0 = g * P.tg;
        const int nt_g = min(P.tg, P.ntaps - tap0);
        mbar_wait(empty0 + 8 * stage, phase ^ 1, P.err_flag, 1);
        const uint32_t sa = stage0 + (uint32_t)stage * (uint32_t)P.stage_bytes;
        const uint32_t fb = full0 + 8 * stage;
        const uint32_t a_bytes = (uint32_t)P.wrows * 16u;
        const int nchunk = 2 * nkc - ((P.ci_odd && cp + nkc == P.ci_pairs) ? 1 : 0);
        const uint32_t w_bytes = (uint32_t)(nkc * nt_g) * b_tap_bytes;  // kc > 1 only with n_groups == 1: contiguous
        mbar_expect_tx(fb, (uint32_t)nchunk * a_bytes + w_bytes);
        for (int c = 0; c < nchunk; ++c)
          bulk_g2s(sa + (uint32_t)c * a_chunk_bytes, a_base + (long long)(2 * cp + c) * P.a_chunk, a_bytes, fb);
        bulk_g2s(sa + (uint32_t)P.kc * a_slot_bytes, w_base + ((long long)cp * P.ntaps + tap0) * ((long long)P.bn * 16),
                 w_bytes, fb);
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }
}

__device__ __forceinline__ void epilogue_dispatch(const TcParams& P, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                                  int lg, int hf, int lane, int tile_rows, uint32_t acc_cols,
                                                  const float* s_bias) {
  if (P.fast_epi) {
    if (P.out_is_16 && P.accumulate)  // last AMP branch of a stage: mean of the branches written as the next stage's operand
      epilogue_fast<true, true, true>(P, tmem_base, tfull0, tempty0, lg, hf, 2, lane, tile_rows, acc_cols, s_bias);
    else if (P.out_is_16)
      epilogue_fast<false, false, true>(P, tmem_base, tfull0, tempty0, lg, hf, 2, lane, tile_rows, acc_cols, s_bias);
    else if (P.res != nullptr && P.accumulate)
      epilogue_fast<true, true>(P, tmem_base, tfull0, tempty0, lg, hf, 2, lane, tile_rows, acc_cols, s_bias);
    else if (P.res != nullptr)
      epilogue_fast<true, false>(P, tmem_base, tfull0, tempty0, lg, hf, 2, lane, tile_rows, acc_cols, s_bias);
    else if (P.accumulate)
      epilogue_fast<false, true>(P, tmem_base, tfull0, tempty0, lg, hf, 2, lane, tile_rows, acc_cols, s_bias);
    else
      epilogue_fast<false, false>(P, tmem_base, tfull0, tempty0, lg, hf, 2, lane, tile_rows, acc_cols, s_bias);
  } else {
    epilogue_role(P, tmem_base, tfull0, tempty0, lg, hf, 2, lane, tile_rows, acc_cols);
  }
}

// Per-CTA setup shared by the conv kernels: mbarriers, tap-offset table, alpha * bias table, zeroed partner windows.
// The caller allocates TMEM and issues the __syncthreads that publishes all of it.
__device__ __forceinline__ void conv_cta_setup(const TcParams& P, unsigned char* smem, int nthreads, int epi_warps) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kMaxStages;
  const uint32_t tfull0 = empty0 + 8 * kMaxStages, tempty0 = tfull0 + 16;
  const int S = P.stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, epi_warps);
    }
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + P.P * P.ntaps) {
    const int i = threadIdx.x - 64;
    reinterpret_cast<int*>(smem + 512)[i] = P.tap_off[i] - P.min_off[i / P.ntaps];
  }
  if (P.fast_epi) {  // alpha * bias for every output channel of this launch (16 zeros when there is no bias)
    float* sb = reinterpret_cast<float*>(smem + 1024 + (size_t)S * P.stage_bytes);
    const int n = P.bias ? ((P.n_tiles * P.bn + 15) & ~15) : 16;
    for (int i = threadIdx.x; i < n; i += nthreads) sb[i] = (P.bias && i < P.Cout) ? P.alpha * __ldg(P.bias + i) : 0.f;
  }
  if (P.ci_odd) {
    // Odd chunk count (e.g. 24 channels): the partner window of the last chunk is never fetched.  Its weights are
    // zero, so it only has to hold finite values: zero every stage's second windows once (stale activations of other
    // pairs that land there later are finite as well).
    const uint32_t cb = (uint32_t)P.arows_pad * 16u;
    for (int st = 0; st < S * P.kc; ++st) {
      uint4* z = reinterpret_cast<uint4*>(smem + 1024 + (size_t)(st / P.kc) * P.stage_bytes + (size_t)(2 * (st % P.kc) + 1) * cb);
      for (uint32_t i = threadIdx.x; i < cb / 16u; i += nthreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
}

__global__ void __launch_bounds__(kThreads, 1) tc_conv_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // [0,256): barriers; [256,260): tmem base; [512,768): tap offsets; stages from 1024; bias table behind the stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kMaxStages;
  const uint32_t tfull0 = empty0 + 8 * kMaxStages, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (2 * kMaxStages + 4));
  const uint32_t stage0 = smem_u32(smem + 1024);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  conv_cta_setup(P, smem, kThreads, 8);
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int* s_off = reinterpret_cast<int*>(smem + 512);  // [P][ntaps] tap offsets relative to the phase minimum

  const uint32_t a_chunk_bytes = (uint32_t)P.arows_pad * 16u;  // one chunk window slot
  const int tile_rows = 128 * P.msub;
  const uint32_t acc_cols = (uint32_t)(P.msub * P.bn);   // TMEM columns of one accumulator stage

  if (warp == 0) {
    // ===================================================================== producer
    if (lane == 0) producer_role(P, full0, empty0, stage0, a_chunk_bytes, tile_rows);
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    mma_role(P, tmem_base, full0, empty0, tfull0, tempty0, stage0, a_chunk_bytes, s_off);
  } else {
    // ===================================================================== epilogue (warps 2..9)
    epilogue_dispatch(P, tmem_base, tfull0, tempty0, warp & 3, (warp - 2) >> 2, lane, tile_rows, acc_cols,
                      reinterpret_cast<const float*>(smem + 1024 + (size_t)P.stages * P.stage_bytes));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------ fused snake + conv
// Same implicit GEMM, but the A operand is produced IN the kernel: conv(Activation1d(x)) without the 16-bit
// activation ever touching HBM (the HBM-bound C <= 128 stages drop from 28 to 20 bytes per element-pair).
// Warp roles (20 warps): 0 = bulk-copy producer of raw fp32 activation windows (own ring of x_stages slots),
// 1 = MMA issuer, 2 = bulk-copy producer of weight slots, 3 = idle, 4..7 = epilogue, 8..19 = snake warps.
// One smem stage = one ci-pair (16 channels) x all taps x (256 + span) rows.  The snake warps turn the raw window
// into the 16-bit K-major chunk image of the A slot in TWO phases with the 2x-rate signal staged in shared memory,
// so that nothing is computed twice (the register-blocked standalone kernel recomputes a 10-sample halo per thread):
//   phase 1: position q -> s[2q], s[2q+1] = snake(up-filter(x))      (12 packed FFMA2 + 2 x (mul, 2 cos, fma) per pair)
//   phase 2: row t      -> y[t] = sum_k f[k] s[2t-5+k]  -> 16-bit    (12 packed FFMA2 per channel pair)
// A thread owns a channel PAIR and kSR consecutive positions / rows in both phases (sliding register window);
// kSR odd and the plane paddings below make every 64-bit shared-memory access of a half-warp conflict free.
// Register budget (setmaxnreg): producers 24, epilogue 168, snake 96 per thread.
__device__ __forceinline__ float2 tc_ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 tc_fmul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

constexpr int kFusedThreads = 640;
constexpr int kSnakeWarps = 12;
constexpr int kSnakeThreads = kSnakeWarps * 32;
constexpr int kSnakeGroups = kSnakeThreads / 8;   // 48 row groups x 8 channel pairs
constexpr int kSR = 7;                            // positions / rows per snake thread (odd)
constexpr int kSnakeCap = kSnakeGroups * kSR;     // 336 >= (256 + span) + 6
constexpr int kXrPad = kSnakeCap + 6;             // raw window rows per chunk plane  (== 2 mod 4)
constexpr int kSrPad = 2 * kSnakeCap + 13;        // 2x-rate rows per chunk plane      (== 1 mod 4)
constexpr int kArPad = kSnakeCap + 4;             // A-slot rows per chunk plane       (== 4 mod 8)
static_assert(kXrPad % 4 == 2 && kSrPad % 4 == 1 && kArPad % 8 == 4 && (kSR & 1), "bank-conflict-free paddings");

__global__ void __launch_bounds__(kFusedThreads, 1) tc_conv_snake_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kMaxStages;
  const uint32_t tfull0 = empty0 + 8 * kMaxStages, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (2 * kMaxStages + 4));
  const uint32_t xfull0 = full0 + 192, xempty0 = full0 + 224;
  int* s_off = reinterpret_cast<int*>(smem + 512);
  float* s_filt = reinterpret_cast<float*>(smem + 768);
  float2* s_par = reinterpret_cast<float2*>(smem + 1024);  // [Cin <= 128] (2 alpha, 1/(2 beta)) per channel
  // [2048 ..): x ring | 2x-rate buffer | A/W stages
  const uint32_t xs_bytes = 2u * kXrPad * 32u;
  const uint32_t sbuf_bytes = 2u * kSrPad * 32u;
  unsigned char* xs_ptr = smem + 2048;
  unsigned char* sbuf_ptr = xs_ptr + (size_t)P.x_stages * xs_bytes;
  const uint32_t xs0 = smem_u32(xs_ptr);
  const uint32_t stage0 = smem_u32(sbuf_ptr + sbuf_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = P.stages, XS = P.x_stages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full0 + 8 * i, 1 + kSnakeWarps);  // weight producer (expect_tx) + snake warps
      mbar_init(empty0 + 8 * i, 1);
    }
    for (int i = 0; i < XS; ++i) {
      mbar_init(xfull0 + 8 * i, 1);
      mbar_init(xempty0 + 8 * i, kSnakeWarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull0 + 8 * i, 1);
      mbar_init(tempty0 + 8 * i, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + P.ntaps) s_off[threadIdx.x - 64] = P.tap_off[threadIdx.x - 64] - P.min_off[0];
  if (threadIdx.x >= 128 && threadIdx.x < 140) s_filt[threadIdx.x - 128] = __ldg(P.sn_filt + threadIdx.x - 128);
  if (threadIdx.x >= 256 && threadIdx.x < 256 + 128) {
    const int c = threadIdx.x - 256;
    s_par[c] = c < P.Cin ? make_float2(2.0f * __ldg(P.sn_a + c), 0.5f * __ldg(P.sn_ib + c)) : make_float2(0.f, 0.f);
  }
  if (P.ci_odd) {  // the partner window of the last (single) chunk only has to be finite: zero it once
    const uint32_t cb = (uint32_t)kArPad * 16u;
    for (int st = 0; st < S; ++st) {
      uint4* z = reinterpret_cast<uint4*>(sbuf_ptr + sbuf_bytes + (size_t)st * P.stage_bytes + cb);
      for (uint32_t i = threadIdx.x; i < cb / 16u; i += kFusedThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t b_tap_bytes = (uint32_t)P.bn * 32u;
  const uint32_t a_chunk_bytes = (uint32_t)kArPad * 16u;
  const uint32_t a_slot_bytes = 2u * a_chunk_bytes;
  const uint32_t x_chunk_bytes = (uint32_t)kXrPad * 32u;  // one 8-channel fp32 window
  const int tile_rows = 128 * P.msub;
  const uint32_t acc_cols = (uint32_t)(P.msub * P.bn);
  const int mn = P.min_off[0];

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == 0) {
      // ===================================================================== raw-activation producer
      if (lane == 0) {
        int xs = 0, xph = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
          const TileCoord tc = decode_tile(P, tile);
          const long long row = (long long)P.a_row0 + (long long)tc.mt * tile_rows + mn - 6;  // >= 0 (host check)
          long long nrows = (long long)P.rows_per_chunk - row;  // stay inside this chunk's rows
          if (nrows > P.xrows) nrows = P.xrows;
          const float* x_base = P.xf + (long long)tc.b * P.a_batch + row * 8;
          for (int cp = 0; cp < P.ci_pairs; ++cp) {
            const bool pair = !(P.ci_odd && cp == P.ci_pairs - 1);
            mbar_wait(xempty0 + 8 * xs, xph ^ 1, P.err_flag, 5);
            const uint32_t dst = xs0 + (uint32_t)xs * xs_bytes;
            const uint32_t fb = xfull0 + 8 * xs;
            mbar_expect_tx(fb, (pair ? 2u : 1u) * (uint32_t)nrows * 32u);
            bulk_g2s(dst, x_base + (long long)(2 * cp) * P.a_chunk, (uint32_t)nrows * 32u, fb);
            if (pair) bulk_g2s(dst + x_chunk_bytes, x_base + (long long)(2 * cp + 1) * P.a_chunk, (uint32_t)nrows * 32u, fb);
            if (++xs == XS) {
              xs = 0;
              xph ^= 1;
            }
          }
        }
      }
    } else if (warp == 2) {
      // ===================================================================== weight producer
      if (lane == 0) {
        int stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
          const TileCoord tc = decode_tile(P, tile);
          const __nv_bfloat16* w_base = P.w + ((long long)tc.nt * P.ci_pairs) * ((long long)P.ntaps * P.bn * 16);
          for (int cp = 0; cp < P.ci_pairs; ++cp) {
            mbar_wait(empty0 + 8 * stage, phase ^ 1, P.err_flag, 1);
            const uint32_t sa = stage0 + (uint32_t)stage * (uint32_t)P.stage_bytes;
            const uint32_t fb = full0 + 8 * stage;
            mbar_expect_tx(fb, (uint32_t)P.ntaps * b_tap_bytes);
            bulk_g2s(sa + a_slot_bytes, w_base + (long long)cp * P.ntaps * ((long long)P.bn * 16),
                     (uint32_t)P.ntaps * b_tap_bytes, fb);
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================================================================== MMA issuer (one stage = one ci-pair)
      mma_role(P, tmem_base, full0, empty0, tfull0, tempty0, stage0, a_chunk_bytes, s_off);
    }
  } else if (warp < 8) {
    // ===================================================================== epilogue (4 warps, one per lane group)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    epilogue_role(P, tmem_base, tfull0, tempty0, warp & 3, 0, 1, lane, tile_rows, acc_cols);
  } else {
    // ===================================================================== snake warps
    const int stid = threadIdx.x - 256;
    const int pc = stid & 7, g = stid >> 3;       // channel pair (0..7 over the two chunks), row group (0..47)
    const int csel = pc >> 2, pcl = pc & 3;
    const int np = P.wrows + 6;                   // 2x-rate positions needed per window
    int stage = 0, phase = 0, xs = 0, xph = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(P, tile);
      const int t_slot0 = tc.mt * tile_rows + mn;  // time of A-slot row 0; position 0 is t_slot0 - 3, raw row 0 is t_slot0 - 6
      const bool edge = (t_slot0 - 6 < 0) || (t_slot0 + P.wrows + 6 > P.L);
      for (int cp = 0; cp < P.ci_pairs; ++cp) {
        const bool active = !(P.ci_odd && cp == P.ci_pairs - 1 && csel == 1);
        const int c0 = cp * 16 + csel * 8 + 2 * pcl;
        const float2 par0 = s_par[c0], par1 = s_par[c0 + 1];
        const float2 al2 = make_float2(par0.x, par1.x), hib = make_float2(par0.y, par1.y);
        const float2 nhib = make_float2(-hib.x, -hib.y);
        // ---------------- phase 1: raw window -> 2x-rate snake samples in shared memory
        // Straight-line code, no per-position guards: positions / rows beyond the window compute on stale (finite or
        // not, never consumed) shared memory and land in padding rows that nothing reads.
        mbar_wait(xfull0 + 8 * xs, xph, P.err_flag, 6);
        float2 xv[kSR + 6];
        {
          const float* xw = reinterpret_cast<const float*>(xs_ptr + (size_t)xs * xs_bytes + (size_t)csel * x_chunk_bytes) + 2 * pcl;
          if (!edge) {
#pragma unroll
            for (int j = 0; j < kSR + 6; ++j) xv[j] = *reinterpret_cast<const float2*>(xw + (g * kSR + j) * 8);
          } else {  // replicate pad: rows t < 0 read x[0], rows t >= L read x[L-1]
#pragma unroll
            for (int j = 0; j < kSR + 6; ++j) {
              int t = t_slot0 - 6 + g * kSR + j;
              t = min(max(t, 0), P.L - 1);
              xv[j] = *reinterpret_cast<const float2*>(xw + (t - (t_slot0 - 6)) * 8);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(xempty0 + 8 * xs);  // the raw window of this stage has been consumed
        if (++xs == XS) {
          xs = 0;
          xph ^= 1;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kSnakeThreads) : "memory");  // phase 2 of the previous stage has read the 2x buffer
        if (active) {
          float fu[12];
#pragma unroll
          for (int k = 0; k < 12; ++k) fu[k] = 2.0f * s_filt[k];
          float2 ue[kSR], uo[kSR];
#pragma unroll
          for (int j = 0; j < kSR; ++j) {
            ue[j] = tc_fmul2(xv[j], make_float2(fu[11], fu[11]));       // d = -3
            uo[j] = tc_fmul2(xv[j + 1], make_float2(fu[10], fu[10]));   // d = -2
          }
#pragma unroll
          for (int d = -2; d <= 2; ++d) {
#pragma unroll
            for (int j = 0; j < kSR; ++j) ue[j] = tc_ffma2(xv[j + 3 + d], make_float2(fu[5 - 2 * d], fu[5 - 2 * d]), ue[j]);
          }
#pragma unroll
          for (int d = -1; d <= 3; ++d) {
#pragma unroll
            for (int j = 0; j < kSR; ++j) uo[j] = tc_ffma2(xv[j + 3 + d], make_float2(fu[6 - 2 * d], fu[6 - 2 * d]), uo[j]);
          }
          float* sw = reinterpret_cast<float*>(sbuf_ptr + (size_t)csel * (kSrPad * 32)) + 2 * pcl + (2 * g * kSR) * 8;
#pragma unroll
          for (int j = 0; j < kSR; ++j) {
            const float2 ze = tc_fmul2(ue[j], al2), zo = tc_fmul2(uo[j], al2);
            const float2 ce = make_float2(__cosf(ze.x), __cosf(ze.y)), co = make_float2(__cosf(zo.x), __cosf(zo.y));
            *reinterpret_cast<float2*>(sw + (2 * j) * 8) = tc_ffma2(ce, nhib, ue[j]);
            *reinterpret_cast<float2*>(sw + (2 * j + 1) * 8) = tc_ffma2(co, nhib, uo[j]);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kSnakeThreads) : "memory");  // the 2x-rate buffer is complete
        // ---------------- phase 2: down-filter -> 16-bit rows of the A slot
        mbar_wait(empty0 + 8 * stage, phase ^ 1, P.err_flag, 7);  // the MMAs that last read this A slot retired
        if (active) {
          const float* sw = reinterpret_cast<const float*>(sbuf_ptr + (size_t)csel * (kSrPad * 32)) + 2 * pcl;
          const int i0 = 2 * g * kSR + 1;  // buffer row i holds s[2 (t_slot0 - 3) + i]; row r needs i = 2r + 1 .. 2r + 12
          float2 sv[2 * kSR + 10];
          if (!edge) {
#pragma unroll
            for (int ii = 0; ii < 2 * kSR + 10; ++ii) sv[ii] = *reinterpret_cast<const float2*>(sw + (i0 + ii) * 8);
          } else {  // replicate clamp of the 2x-rate signal: s[m < 0] = s[0], s[m > 2L-1] = s[2L-1]
            const int mb = 2 * (t_slot0 - 3);
#pragma unroll
            for (int ii = 0; ii < 2 * kSR + 10; ++ii) {
              int m = mb + i0 + ii;
              m = min(max(m, 0), 2 * P.L - 1);
              int i = m - mb;
              i = min(max(i, 0), kSrPad - 1);
              sv[ii] = *reinterpret_cast<const float2*>(sw + i * 8);
            }
          }
          float2 acc[kSR];
#pragma unroll
          for (int j = 0; j < kSR; ++j) acc[j] = hib;
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            const float fk = s_filt[k];
#pragma unroll
            for (int j = 0; j < kSR; ++j) acc[j] = tc_ffma2(make_float2(fk, fk), sv[2 * j + k], acc[j]);
          }
          const uint32_t a_dst = stage0 + (uint32_t)stage * (uint32_t)P.stage_bytes + (uint32_t)csel * a_chunk_bytes +
                                 (uint32_t)(g * kSR) * 16u + (uint32_t)pcl * 4u;
          const int t0 = t_slot0 + g * kSR;
          uint32_t hv[kSR];
          if (P.fp16) {  // block-uniform
#pragma unroll
            for (int j = 0; j < kSR; ++j) hv[j] = fh::pack16(acc[j].x, acc[j].y, 1);
          } else {
#pragma unroll
            for (int j = 0; j < kSR; ++j) hv[j] = fh::pack16(acc[j].x, acc[j].y, 0);
          }
#pragma unroll
          for (int j = 0; j < kSR; ++j) {
            const unsigned tt = (unsigned)(t0 + j);
            const uint32_t v = tt < (unsigned)P.L ? hv[j] : 0u;  // conv zero padding outside [0, L)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(a_dst + (uint32_t)j * 16u), "r"(v) : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * stage);
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------ dual kernel: conv || snake
// The vocoder alternates a tensor-pipe / HBM-bound convolution with an FP32-pipe-bound anti-aliased snake; run back to
// back, each leaves the other pipe of the SM idle (and streams could not make them co-resident: the conv CTA holds
// most of the register file and shared memory).  This kernel is both at once, for two independent half-batches: the
// ten conv warps of tc_conv_kernel (producer, MMA issuer, eight epilogue warps) work on half-batch A while two
// 128-thread snake workers (snake_worker.cuh) walk the tiles of half-batch B.  Register budget via setmaxnreg:
// producers 24, epilogue 128, snake 96.  The engine staggers the two half-batches by one operator so that every
// launch pairs a conv with a snake.
constexpr int kDualThreads = 640;  // warps 0..3 producer / MMA / 2 idle, 4..11 epilogue, 12..19 snake (2 workers)
constexpr int kDualWorkers = 2;
constexpr int kDualR = 11;         // snake outputs per thread (96-register budget)

struct DualParams {
  TcParams c;
  fh::SnakeParams s;
  int snake_off;   // byte offset of the snake workers' shared memory
  int s_out_kind;  // 1 bf16, 2 fp16
};

__global__ void __launch_bounds__(kDualThreads, 1) tc_conv_snake_dual_kernel(const __grid_constant__ DualParams D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const TcParams& P = D.c;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * kMaxStages;
  const uint32_t tfull0 = empty0 + 8 * kMaxStages, tempty0 = tfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (2 * kMaxStages + 4));
  const uint32_t stage0 = smem_u32(smem + 1024);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  conv_cta_setup(P, smem, kDualThreads, 8);
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int* s_off = reinterpret_cast<int*>(smem + 512);
  const uint32_t a_chunk_bytes = (uint32_t)P.arows_pad * 16u;
  const int tile_rows = 128 * P.msub;
  const uint32_t acc_cols = (uint32_t)(P.msub * P.bn);

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == 0) {
      if (lane == 0) producer_role(P, full0, empty0, stage0, a_chunk_bytes, tile_rows);
    } else if (warp == 1) {
      mma_role(P, tmem_base, full0, empty0, tfull0, tempty0, stage0, a_chunk_bytes, s_off);
    }
  } else if (warp < 12) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    epilogue_dispatch(P, tmem_base, tfull0, tempty0, warp & 3, (warp - 4) >> 2, lane, tile_rows, acc_cols,
                      reinterpret_cast<const float*>(smem + 1024 + (size_t)P.stages * P.stage_bytes));
  } else {
    const int grp = (warp - 12) >> 2;
    const int tid = threadIdx.x - 384 - grp * 128;
    unsigned char* sm = smem + D.snake_off + (size_t)grp * fh::SnakeGeom<kDualR>::kSmemBytes;
    const int worker = blockIdx.x * kDualWorkers + grp, nworkers = gridDim.x * kDualWorkers;
    fh::snake_worker<3, true, kDualR>(D.s, sm, tid, worker, nworkers, 1 + grp);  // one instantiation: small I-cache footprint
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------ layout helpers
__global__ void to_chunked_bf16_kernel(const float* __restrict__ src, long long src_batch, long long src_c,
                                       long long src_t, __nv_bfloat16* __restrict__ dst, long long dst_batch,
                                       long long dst_chunk, int dst_row0, int C, int L, int fp16, unsigned int* status) {
  // one thread per (t, chunk): gathers 8 channels, writes one 16-byte row
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nch = (C + 7) >> 3;
  const int b = blockIdx.y;
  if (i >= (long long)nch * L) return;
  int t, ch;
  if (src_t == 1) {  // planar source: adjacent threads walk along t
    t = (int)(i % L);
    ch = (int)(i / L);
  } else {  // row-major source: adjacent threads walk along channels
    ch = (int)(i % nch);
    t = (int)(i / nch);
  }
  const float* s = src + (long long)b * src_batch + (long long)t * src_t;
  uint32_t h[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c0 = ch * 8 + 2 * k;
    const float f0 = c0 < C ? s[(long long)c0 * src_c] : 0.f;
    const float f1 = c0 + 1 < C ? s[(long long)(c0 + 1) * src_c] : 0.f;
    h[k] = fh::pack16_guard(f0, f1, fp16, status);
  }
  *reinterpret_cast<uint4*>(dst + (long long)b * dst_batch + (long long)ch * dst_chunk + (long long)(dst_row0 + t) * 8) =
      *reinterpret_cast<uint4*>(h);
}

}  // namespace

// ================================================================================ C ABI
extern "C" __attribute__((visibility("default"))) int64_t fh_tc_packed_weight_bytes(int Cin, int Cout, int ntaps, int P, int bn) {
  if (Cin <= 0 || Cout <= 0 || ntaps <= 0 || P <= 0 || bn <= 0 || (Cin % 8) || (bn % 16)) return -1;
  const int64_t n_tiles = (Cout + bn - 1) / bn;
  return (int64_t)P * n_tiles * ((Cin + 15) / 16) * ntaps * bn * 32;
}

// Validates the arguments and derives the launch plan (tile shape, stages, shared-memory carve-up).  `budget_bytes` is
// the shared memory the conv roles may use (0 = default); the fused-snake variant launches from here (returns 1).
static int tc_plan(const fh_tc_conv_args* a, void* stream, int budget_bytes, TcParams& p, int* smem_out) {
  FH_REQUIRE(a != nullptr, FH_ERR_BAD_SHAPE, "fh_tc_conv: null args");
  FH_REQUIRE(a->B > 0 && a->L > 0 && a->Cin > 0 && a->Cout > 0, FH_ERR_BAD_SHAPE, "fh_tc_conv: bad shape");
  FH_REQUIRE(a->Cin % 8 == 0, FH_ERR_BAD_SHAPE, "fh_tc_conv: Cin=%d must be a multiple of 8", a->Cin);
  FH_REQUIRE(a->Cout % 8 == 0, FH_ERR_BAD_SHAPE, "fh_tc_conv: Cout=%d must be a multiple of 8", a->Cout);
  FH_REQUIRE(a->bn % 16 == 0 && a->bn >= 16 && a->bn <= 256, FH_ERR_UNSUPPORTED_CFG,
             "fh_tc_conv: bn=%d must be a multiple of 16 in [16,256]", a->bn);
  FH_REQUIRE(a->P >= 1 && a->P <= 16 && a->ntaps >= 1 && a->P * a->ntaps <= kMaxTapOff, FH_ERR_UNSUPPORTED_CFG,
             "fh_tc_conv: P=%d ntaps=%d unsupported", a->P, a->ntaps);
  FH_REQUIRE(!(a->geglu && (a->res || a->accumulate || (a->Cout % 16))), FH_ERR_UNSUPPORTED_CFG,
             "fh_tc_conv: geglu epilogue excludes residual/accumulate and needs Cout %% 16 == 0");
  FH_REQUIRE(!(a->accumulate && a->out_is_16) || (a->acc_src && a->res && !a->res_is_16), FH_ERR_UNSUPPORTED_CFG,
             "fh_tc_conv: accumulate into a 16-bit output needs acc_src (fp32) and an fp32 residual");
  FH_REQUIRE(((uintptr_t)a->a % 16) == 0 && ((uintptr_t)a->w % 16) == 0 && ((uintptr_t)a->out % 16) == 0 &&
                 ((uintptr_t)a->res % 16) == 0,
             FH_ERR_BAD_ALIGN, "fh_tc_conv: pointers must be 16-byte aligned");
  FH_REQUIRE(a->a_chunk % 8 == 0 && a->a_batch % 8 == 0, FH_ERR_BAD_ALIGN, "fh_tc_conv: A strides must be x8");
  const int esz_shift = a->out_is_16 ? 3 : 2;  // 16-byte alignment in elements
  FH_REQUIRE((a->out_batch % (1 << esz_shift)) == 0 && (a->out_chunk % (1 << esz_shift)) == 0 &&
                 (a->out_row % (1 << esz_shift)) == 0,
             FH_ERR_BAD_ALIGN, "fh_tc_conv: output strides break 16-byte alignment");

  memset(&p, 0, sizeof(p));
  p.a = (const __nv_bfloat16*)a->a;
  p.w = (const __nv_bfloat16*)a->w;
  p.bias = a->bias;
  p.res = a->res;
  p.out = a->out;
  p.a_batch = a->a_batch, p.a_chunk = a->a_chunk, p.a_row0 = a->a_row0;
  p.out_batch = a->out_batch, p.out_chunk = a->out_chunk, p.out_row = a->out_row;
  p.res_batch = a->res_batch, p.res_chunk = a->res_chunk, p.res_row = a->res_row;
  p.out_is_16 = a->out_is_16, p.res_is_16 = a->res_is_16, p.fp16 = a->fp16;
  p.accumulate = a->accumulate, p.geglu = a->geglu;
  p.acc_src = a->accumulate ? a->acc_src : nullptr;
  p.act_gelu = a->act == 1;
  FH_REQUIRE(a->act == 0 || (a->act == 1 && !a->geglu), FH_ERR_UNSUPPORTED_CFG, "fh_tc_conv: act must be 0 (none) or 1 (GELU, not with geglu)");
  p.alpha = a->alpha, p.beta_res = a->beta_res;
  p.status = fh::status_word();
  p.B = a->B, p.L = a->L, p.Cin = a->Cin, p.Cout = a->Cout, p.ntaps = a->ntaps, p.P = a->P, p.bn = a->bn;
  p.n_tiles = (a->Cout + a->bn - 1) / a->bn;
  const bool fused = a->x_f32 != nullptr;
  if (fused) {
    FH_REQUIRE(a->P == 1 && p.n_tiles == 1 && a->bn <= 128 && a->Cin <= 128 && a->sn_a && a->sn_inv_b && a->sn_filt && !a->geglu,
               FH_ERR_UNSUPPORTED_CFG, "fh_tc_conv: fused snake needs P == 1 and one N tile of <= 128 columns");
    FH_REQUIRE(((uintptr_t)a->x_f32 % 16) == 0, FH_ERR_BAD_ALIGN, "fh_tc_conv: x_f32 must be 16-byte aligned");
  }
  // sub-tiles: reuse each weight slot for up to 4 x 128 rows when the accumulators fit TMEM twice over
  int msub = 256 / a->bn;
  msub = msub >= 8 ? 8 : (msub >= 4 ? 4 : (msub >= 2 ? 2 : 1));  // 8 x 128 rows for bn <= 32 (24-channel last stage)
  // Wide tiles (bn > 128): the weight stream from L2 (bn*32 B per MMA) is the limiter, so two 128-row
  // sub-tiles share every weight slot even though the accumulators (2*bn columns) then fill TMEM and
  // the epilogue no longer overlaps the next tile -- worth it once a tile carries enough MMAs.
  static int wide_msub = -1, wide_min = 96;
  if (wide_msub < 0) {
    const char* e = getenv("FH_TC_WIDE_MSUB");
    wide_msub = e ? atoi(e) : 2;
    const char* m = getenv("FH_TC_WIDE_MIN");
    if (m) wide_min = atoi(m);
  }
  if (a->bn > 128 && wide_msub == 2 && (long long)a->ntaps * ((a->Cin + 15) / 16) >= wide_min) msub = 2;
  while (!fused && msub > 1 &&
         (long long)a->B * a->P * ((a->L + 128 * msub - 1) / (128 * msub)) * p.n_tiles < 2 * 148)
    msub >>= 1;
  if (fused) msub = 2;  // 256-row tiles: the snake warps' two-phase window (tc_conv_snake_kernel) is sized for them
  p.msub = msub;
  p.acc_stages = (2 * msub * a->bn <= 512) ? 2 : 1;
  p.m_tiles = (a->L + 128 * msub - 1) / (128 * msub);
  const long long total = (long long)p.B * p.P * p.m_tiles * p.n_tiles;
  FH_REQUIRE(total < (1ll << 31), FH_ERR_BAD_SHAPE, "fh_tc_conv: too many tiles");
  p.total_tiles = (int)total;
  p.ci_pairs = (a->Cin + 15) / 16;
  p.ci_odd = (a->Cin % 16) != 0;
  int span = 0, arith = 1;
  for (int ph = 0; ph < a->P; ++ph) {
    int mn = a->tap_off[ph * a->ntaps], mx = mn;
    for (int m = 0; m < a->ntaps; ++m) {
      const int o = a->tap_off[ph * a->ntaps + m];
      p.tap_off[ph * a->ntaps + m] = o;
      mn = o < mn ? o : mn;
      mx = o > mx ? o : mx;
    }
    p.min_off[ph] = mn;
    p.tap_rel0[ph] = a->tap_off[ph * a->ntaps] - mn;
    p.tap_step[ph] = a->ntaps > 1 ? a->tap_off[ph * a->ntaps + 1] - a->tap_off[ph * a->ntaps] : 0;
    for (int m = 1; m < a->ntaps; ++m)
      if (a->tap_off[ph * a->ntaps + m] - a->tap_off[ph * a->ntaps + m - 1] != p.tap_step[ph]) arith = 0;
    span = (mx - mn) > span ? (mx - mn) : span;
    FH_REQUIRE(a->a_row0 + mn >= 0, FH_ERR_BAD_SHAPE, "fh_tc_conv: left halo %d too small for tap offset %d",
               a->a_row0, mn);
  }
  FH_REQUIRE(span <= kMaxSpan, FH_ERR_UNSUPPORTED_CFG, "fh_tc_conv: tap span %d exceeds %d rows", span, kMaxSpan);
  p.tap_arith = arith;
  {
    static int use_v8 = -1;
    if (use_v8 < 0) {
      const char* e = getenv("FH_TC_V8");
      use_v8 = e ? atoi(e) : 1;
    }
    const bool out_ok = a->out_is_16 || (((uintptr_t)a->out % 32) == 0 && a->out_batch % 8 == 0 && a->out_chunk % 8 == 0 &&
                                         a->out_row % 8 == 0);
    const bool res_ok = a->res == nullptr || a->res_is_16 ||
                        (((uintptr_t)a->res % 32) == 0 && a->res_batch % 8 == 0 && a->res_chunk % 8 == 0 && a->res_row % 8 == 0);
    const bool acc_ok = !(a->accumulate && a->acc_src) ||
                        (((uintptr_t)a->acc_src % 32) == 0 && a->out_batch % 8 == 0 && a->out_chunk % 8 == 0 && a->out_row % 8 == 0);
    p.v8 = (use_v8 && out_ok && res_ok && acc_ok) ? 1 : 0;
  }
  p.wrows = 128 * msub + span;
  p.arows_pad = 128 * msub + kMaxSpan;
  if (fused) {
    FH_REQUIRE(p.wrows + 6 <= kSnakeCap, FH_ERR_UNSUPPORTED_CFG, "fh_tc_conv: fused snake window too small for this tap span");
    FH_REQUIRE(a->a_row0 + p.min_off[0] - 6 >= 0, FH_ERR_BAD_SHAPE, "fh_tc_conv: fused snake needs 6 more halo rows");
    p.arows_pad = kArPad;
    p.xrows = p.wrows + 12;  // raw rows needed: positions t_slot0 - 3 .. t_slot0 + wrows + 2, +-3 each
    p.xs_bytes = 2 * kXrPad * 32;
    p.rows_per_chunk = (int)(a->a_chunk / 8);
    p.xf = a->x_f32, p.sn_a = a->sn_a, p.sn_ib = a->sn_inv_b, p.sn_filt = a->sn_filt;
  }
  // taps per stage: a weight slot of <= 32 KB (<= 48 KB for narrow tiles, so that all taps of a ci-pair share
  // one activation window fetch), with the taps spread evenly over the groups (11 taps -> 6+5, not 10+1)
  static int slot_small_kb = 0, slot_wide_kb = 0;
  if (!slot_small_kb) {
    const char* e1 = getenv("FH_TC_SLOT_SMALL_KB");
    const char* e2 = getenv("FH_TC_SLOT_WIDE_KB");
    slot_small_kb = e1 ? atoi(e1) : 48;
    slot_wide_kb = e2 ? atoi(e2) : 48;
  }
  int tg = ((a->bn <= 128 ? slot_small_kb : slot_wide_kb) * 1024) / (a->bn * 32);
  if (fused) tg = a->ntaps;  // one stage = one ci-pair with all its taps
  if (tg < 1) tg = 1;
  if (tg > a->ntaps) tg = a->ntaps;
  p.n_groups = (a->ntaps + tg - 1) / tg;
  tg = (a->ntaps + p.n_groups - 1) / p.n_groups;
  p.tg = tg;
  // ci-pairs per stage: a stage should carry >= ~12 MMAs so that its barrier round trip is amortised (Linear layers are
  // one tap per ci-pair, k = 3 convs three), as long as at least 3-4 stages still fit
  static int kc_target = -1;
  if (kc_target < 0) {
    const char* e = getenv("FH_TC_KC_MMAS");
    kc_target = e ? atoi(e) : 12;
  }
  int kc = 1;
  if (!fused && p.n_groups == 1 && kc_target > 0) {
    const int pair_bytes = 2 * p.arows_pad * 16 + a->ntaps * a->bn * 32;
    kc = (kc_target + a->ntaps * msub - 1) / (a->ntaps * msub);
    if (kc > 4) kc = 4;
    if (kc > p.ci_pairs) kc = p.ci_pairs;
    while (kc > 1 && (200 * 1024 - 1024 - 8192) / (kc * pair_bytes) < 4) --kc;
  }
  p.kc = kc;
  p.stage_bytes = kc * (2 * p.arows_pad * 16 + tg * a->bn * 32);
  p.stage_bytes = (p.stage_bytes + 127) & ~127;
  static int budget_kb = 0;
  if (!budget_kb) {
    const char* e = getenv("FH_TC_SMEM_KB");
    budget_kb = e ? atoi(e) : 200;
    if (budget_kb < 64 || budget_kb > 220) budget_kb = 200;
  }
  const int budget = budget_bytes > 0 ? budget_bytes : budget_kb * 1024;
  if (fused) {
    FH_REQUIRE(budget_bytes == 0, FH_ERR_UNSUPPORTED_CFG, "fh_tc_conv: the fused snake prologue cannot run in the dual kernel");
    const int fbudget = 220 * 1024;  // this kernel owns the SM
    const int sbuf = 2 * kSrPad * 32;
    p.x_stages = 3;
    int st = (fbudget - 2048 - sbuf - p.x_stages * p.xs_bytes) / p.stage_bytes;
    if (st < 3) {
      p.x_stages = 2;
      st = (fbudget - 2048 - sbuf - p.x_stages * p.xs_bytes) / p.stage_bytes;
    }
    FH_REQUIRE(st >= 2, FH_ERR_UNSUPPORTED_CFG, "fh_tc_conv: fused snake stages do not fit shared memory");
    p.stages = st > 4 ? 4 : st;
    p.err_flag = nullptr;
    const int fsmem = 2048 + sbuf + p.x_stages * p.xs_bytes + p.stages * p.stage_bytes;
    static int fsmem_set[64] = {0};
    const int fsms = fh::dev_sms();
    {
      cudaError_t e = fh::ensure_dyn_smem(tc_conv_snake_kernel, fsmem, fsmem_set);
      FH_REQUIRE(e == cudaSuccess, FH_ERR_CUDA, "fh_tc_conv: cannot opt in to %d bytes of smem: %s", fsmem,
                 cudaGetErrorString(e));
    }
    const int fgrid = p.total_tiles < fsms ? p.total_tiles : fsms;
    tc_conv_snake_kernel<<<fgrid, kFusedThreads, fsmem, (cudaStream_t)stream>>>(p);
    const int rc = fh::check_launch("fh_tc_conv(fused snake)");
    return rc == FH_OK ? 1 : rc;
  }
  static int fast_on = -1;
  if (fast_on < 0) {
    const char* e = getenv("FH_TC_FAST_EPI");
    fast_on = e ? atoi(e) : 1;
  }
  const int bias_tab = a->bias ? ((p.n_tiles * a->bn + 15) & ~15) * 4 : 64;
  p.fast_epi = (fast_on && p.v8 && !a->geglu && !p.act_gelu && !(a->res && a->res_is_16) && bias_tab <= 8192 &&
                (!a->out_is_16 || (a->res == nullptr && !a->accumulate) || (a->accumulate && a->acc_src && a->res))) ? 1 : 0;
  FH_REQUIRE(p.fast_epi || !(a->accumulate && a->out_is_16), FH_ERR_UNSUPPORTED_CFG,
             "fh_tc_conv: accumulate into a 16-bit output is only implemented by the specialised epilogue");
  const int tail = p.fast_epi ? bias_tab : 0;
  int stages = (budget - 1024 - tail) / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  FH_REQUIRE(stages >= 2, FH_ERR_UNSUPPORTED_CFG, "fh_tc_conv: stage of %d bytes does not fit twice", p.stage_bytes);
  p.stages = stages;
  p.err_flag = nullptr;
  *smem_out = 1024 + stages * p.stage_bytes + tail;
  return FH_OK;
}

static int device_sms() { return fh::dev_sms(); }

extern "C" __attribute__((visibility("default"))) int fh_tc_conv(const fh_tc_conv_args* a, void* stream) {
  TcParams p;
  int smem = 0;
  const int rc = tc_plan(a, stream, 0, p, &smem);
  if (rc != FH_OK) return rc == 1 ? FH_OK : rc;  // 1: the fused-snake variant was launched by the planner
  static int smem_set[64] = {0};
  {
    cudaError_t e = fh::ensure_dyn_smem(tc_conv_kernel, smem, smem_set);
    FH_REQUIRE(e == cudaSuccess, FH_ERR_CUDA, "fh_tc_conv: cannot opt in to %d bytes of smem: %s", smem,
               cudaGetErrorString(e));
  }
  const int num_sms = device_sms();
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  tc_conv_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(p);
  return fh::check_launch("fh_tc_conv");
}

// One launch = the convolution of one half-batch (tensor pipe, HBM) + the anti-aliased snake of the other half-batch
// (FP32 pipe) on the same SMs: see tc_conv_snake_dual_kernel.
extern "C" __attribute__((visibility("default"))) int fh_tc_conv_snake_dual(
    const fh_tc_conv_args* a, const float* sx, void* sy, const float* sa, const float* sinv_b, const float* sfilt,
    int64_t s_batch_stride, int64_t s_chunk_stride, int s_row0, int sB, int sC, int sL, int s_out_kind, void* stream) {
  FH_REQUIRE(a != nullptr && a->x_f32 == nullptr, FH_ERR_BAD_SHAPE, "fh_tc_conv_snake_dual: needs a plain conv");
  FH_REQUIRE(sB > 0 && sC > 0 && (sC % 8) == 0 && sL > 0 && (s_out_kind == 1 || s_out_kind == 2), FH_ERR_BAD_SHAPE,
             "fh_tc_conv_snake_dual: snake needs C %% 8 == 0 and a 16-bit output");
  FH_REQUIRE(((uintptr_t)sx % 16) == 0 && ((uintptr_t)sy % 16) == 0 && (s_batch_stride % 8) == 0 &&
                 (s_chunk_stride % 8) == 0 && s_row0 >= 5,
             FH_ERR_BAD_ALIGN, "fh_tc_conv_snake_dual: snake buffers must be 16-byte aligned with a left halo of >= 5 rows");
  using G = fh::SnakeGeom<kDualR>;
  constexpr int kSnakeSmem = kDualWorkers * G::kSmemBytes;
  DualParams d;
  int csmem = 0;
  const int rc = tc_plan(a, stream, 226 * 1024 - kSnakeSmem, d.c, &csmem);
  if (rc != FH_OK) return rc;
  d.snake_off = (csmem + 127) & ~127;
  const int ntile = (sL + G::kRows - 1) / G::kRows;
  const long long total = (long long)ntile * (sC / 8) * sB;
  FH_REQUIRE(total <= 2147483647LL, FH_ERR_BAD_SHAPE, "fh_tc_conv_snake_dual: too many snake tiles");
  d.s.x = sx, d.s.y = sy, d.s.a = sa, d.s.inv_b = sinv_b, d.s.filt = sfilt;
  d.s.batch_stride = s_batch_stride, d.s.chunk_stride = s_chunk_stride;
  d.s.row0 = s_row0, d.s.nchunk = sC / 8, d.s.L = sL, d.s.ntile = ntile, d.s.total = (int)total;
  d.s_out_kind = s_out_kind;
  d.s.fp16 = s_out_kind == 2;
  d.s.status = fh::status_word();
  const int smem = d.snake_off + kSnakeSmem;
  static int smem_set[64] = {0};
  {
    cudaError_t e = fh::ensure_dyn_smem(tc_conv_snake_dual_kernel, smem, smem_set);
    FH_REQUIRE(e == cudaSuccess, FH_ERR_CUDA, "fh_tc_conv_snake_dual: cannot opt in to %d bytes of smem: %s", smem,
               cudaGetErrorString(e));
  }
  tc_conv_snake_dual_kernel<<<device_sms(), kDualThreads, smem, (cudaStream_t)stream>>>(d);
  return fh::check_launch("fh_tc_conv_snake_dual");
}

extern "C" __attribute__((visibility("default"))) int fh_to_chunked_16(const float* src, int64_t src_batch, int64_t src_c, int64_t src_t, void* dst,
                                  int64_t dst_batch, int64_t dst_chunk, int dst_row0, int B, int C, int L,
                                  int fp16, void* stream) {
  FH_REQUIRE(B > 0 && C > 0 && L > 0 && B <= 65535, FH_ERR_BAD_SHAPE, "fh_to_chunked_16: bad shape");
  const long long n = (long long)((C + 7) / 8) * L;
  to_chunked_bf16_kernel<<<dim3((unsigned)((n + 255) / 256), B), 256, 0, (cudaStream_t)stream>>>(
      src, src_batch, src_c, src_t, (__nv_bfloat16*)dst, dst_batch, dst_chunk, dst_row0, C, L, fp16, fh::status_word());
  return fh::check_launch("fh_to_chunked_16");
}
