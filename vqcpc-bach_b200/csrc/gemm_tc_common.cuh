// Shared pieces of the tcgen05 GEMM kernels (1-CTA / multicast kernel in gemm_tc.cu, CTA-pair kernel in gemm_tc2.cu):
// PTX wrappers, shared-memory / instruction descriptors, parameters and the TMEM -> global epilogue.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace vqtc {


constexpr int BM = 128;
constexpr int UMMA_K = 16;
constexpr int NUM_ACC = 2;
constexpr int THREADS = 192;
// Tile configurations (template parameters BN, BK):
//   128 x 128 x 64, 3 stages  - narrow N; every MMA reads 8 KB of smem per 64 cycles (at the 128 B/clk smem limit)
//   128 x 256 x 32, 4 stages  - default for N >= 256: 12 KB per 128 cycles, and a deeper ring to hide TMA latency
// K-major operand tiles use the swizzle whose span equals the tile row (BK*2 bytes: 128B or 64B);
// MN-major tiles are chunks of [BK k-rows][64 mn] with the 128B swizzle.

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 1023u) == 1023u) {
      uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("vqcpc gemm_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// same, multicast: the box lands at the same smem offset in every CTA of `mask` and completes tx on the mbarrier at
// the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `mask` once the previously issued MMAs have retired
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16 TMEM lanes x 32 columns in the mma-accumulator layout: thread t holds, for each 8-column group g = 0..3,
// r[4g + 0..1] = (lane t/4, columns 8g + 2(t%4) + {0,1}) and r[4g + 2..3] = the same columns of lane t/4 + 8
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// fp32 pair -> packed bf16 (hi, lo) words with x ~= hi + lo: one cvt.rn.bf16x2.f32 per two values (same rounding as the
// element-wise __float2bfloat16_rn, half the conversion instructions and no byte permutes)
__device__ __forceinline__ void split_pair_bf16(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);                     // .x (low half) = x, .y = y
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - __uint_as_float(hi << 16), y - __uint_as_float(hi & 0xffff0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 format: version=1 at bit 46, layout type at bits 61..63; SWIZZLE_128B = 2).
//   K-major  tile [rows][64 bf16] (128 B rows, 8-row swizzle atoms): SBO = 1024 B, LBO unused (=16 B)
//   MN-major tile = chunks of [64 k-rows][64 mn] (8 KB each, one TMA box): LBO = 8192 B between 64-wide mn chunks,
//                   SBO = 1024 B between groups of 8 k-rows
template <int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool kmajor) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  // K-major: rows of BK*2 bytes, 8-row swizzle atoms -> SBO = 8 * BK * 2; LBO unused.
  // MN-major: 64-wide (128 B) mn chunks of BK k-rows each -> LBO = BK * 128 between chunks, SBO = 1024 (8 k-rows).
  uint32_t lbo = kmajor ? 16u : (uint32_t)(BK * 128);
  uint32_t sbo = kmajor ? (uint32_t)(16 * BK) : 1024u;
  uint32_t layout = (kmajor && BK == 32) ? 4u : 2u;       // SWIZZLE_64B : SWIZZLE_128B
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// The descriptor's high word and the LBO half of the low word are compile-time constants; per MMA only the 14-bit start
// address changes, so the issuing thread (which must stay ahead of a 64..128-cycle MMA) does one add per operand.
template <int BK>
__host__ __device__ constexpr uint32_t desc_hi32(bool kmajor) {
  return (uint32_t)((kmajor ? 16 * BK : 1024) >> 4) | (1u << 14) | (((kmajor && BK == 32) ? 4u : 2u) << 29);
}
template <int BK>
__host__ __device__ constexpr uint32_t desc_lo32_const(bool kmajor) {
  return (uint32_t)(((kmajor ? 16 : BK * 128) >> 4)) << 16;
}
__device__ __forceinline__ uint64_t pack_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

template <int BN>
__host__ __device__ constexpr uint32_t make_idesc(bool a_kmajor, bool b_kmajor) {
  return (1u << 4)                         // D format: f32
         | (1u << 7)                       // A format: bf16
         | (1u << 10)                      // B format: bf16
         | ((a_kmajor ? 0u : 1u) << 15)    // A major: 0 = K, 1 = MN
         | ((b_kmajor ? 0u : 1u) << 16)    // B major
         | ((uint32_t)(BN >> 3) << 17)     // N
         | ((uint32_t)(BM >> 4) << 24);    // M
}

struct TcParams {
  int M, N, K;
  float* C;                       // may be NULL when only the split outputs are wanted
  int ldc;
  __nv_bfloat16* out_hi;          // optional fused split of the result (operand of the next GEMM), may be NULL
  __nv_bfloat16* out_lo;
  int ldo;
  EpilogueParams ep;
  int m_tiles, n_tiles, ksplit, kblocks_per_split, kblocks_total;
  int exp;                        // timing experiments (vqcpc_gemm_set_experiment; results invalid when non-zero)
  unsigned long long* dbg;        // optional phase timers (vqcpc_gemm_set_debug): cycles summed over CTAs, see gemm_tc2.cu
};

// Epilogue of one 128 x BN accumulator tile held in TMEM at `taddr` (lane quadrant q of this warp): 32-column chunks,
// tcgen05.ld -> fused bias / scale / residual / ReLU / mask -> fp32 and / or bf16-split stores (or atomics for split-K).
// CH0 / CH1: range of 32-column chunks this warp drains (all of them by default; the 8-epilogue-warp variant of the pair
// kernel gives each of the two warps of a lane quadrant one half of the columns).
// Thread = row.  Round 2: the side inputs are software-pipelined - the tile's bias values are staged once per tile in the
// warp's shared-memory buffer (`wbuf`, before the accumulator is waited for) and the residual / mask rows of chunk c + 1
// are requested before chunk c is processed; round 1 issued them after each tcgen05.ld and exposed one L2 latency per
// chunk, which is what held the K = 512 shapes with a residual at 780 tensor-TFLOP/s against 1200 without one.
template <int BN, int CH0 = 0, int CH1 = BN / 32>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, uint32_t taddr, int m0, int n0, int q, int lane, bool first_split,
                                              uint8_t* wbuf) {
  static_assert(((CH1 - CH0) & 1) == 0, "chunk range must be even (the loop is unrolled by two)");
  const int row = m0 + q * 32 + lane;
  const int fl = p.ep.flags;
  const bool row_ok = row < p.M;
  const bool aligned = (!p.C || (((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.ldc & 3) == 0))) &&
                       (!p.out_hi || (((p.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out_hi) & 15) == 0) &&
                                      ((reinterpret_cast<uintptr_t>(p.out_lo) & 15) == 0))) &&
                       (!(fl & EPI_BIAS) || ((reinterpret_cast<uintptr_t>(p.ep.bias) & 15) == 0)) &&
                       (!(fl & (EPI_RESIDUAL | EPI_RELU_MASK)) ||
                        (((p.ep.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.ep.residual) & 15) == 0)));
  const bool fast = aligned && !(fl & EPI_ATOMIC);
  const bool use_bias = first_split && (fl & EPI_BIAS);
  const bool use_res = first_split && (fl & EPI_RESIDUAL);
  const bool mask_bf16 = (fl & EPI_RELU_MASK) && (fl & EPI_MASK_BF16);
  const bool mask_f32 = (fl & EPI_RELU_MASK) && !(fl & EPI_MASK_BF16);
  float* bias_s = reinterpret_cast<float*>(wbuf);

  if (fast && use_bias) {
    // the bias values of this warp's chunks: (CH1 - CH0) * 32 floats, 8 per lane where needed
    __syncwarp();
    for (int j = lane * 4; j < (CH1 - CH0) * 32; j += 128) {
      const int col = n0 + CH0 * 32 + j;
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col + 4 <= p.N) b = __ldg(reinterpret_cast<const float4*>(p.ep.bias + col));
      *reinterpret_cast<float4*>(bias_s + j) = b;
    }
    __syncwarp();
  }
  // residual / fp32 mask: 8 x float4 of this thread's row; bf16 mask: 4 x uint4 in s[0..3]
  auto load_side = [&](int col0, float4 (&s)[8]) {
    if (!row_ok || !fast || col0 + 32 > p.N) return;
    if (use_res || mask_f32) {
      const float4* rp = reinterpret_cast<const float4*>(p.ep.residual + (size_t)row * p.ep.ldr + col0);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) s[q4] = __ldg(rp + q4);
    } else if (mask_bf16) {
      const float4* mp = reinterpret_cast<const float4*>(reinterpret_cast<const __nv_bfloat16*>(p.ep.residual) + (size_t)row * p.ep.ldr + col0);
#pragma unroll
      for (int q8 = 0; q8 < 4; ++q8) s[q8] = __ldg(mp + q8);
    }
  };
  long long t_ld = 0;
  auto chunk = [&](int ch, float4 (&side)[8], float4 (&side_next)[8]) {
    uint32_t r[32];
    const long long tl0 = p.dbg ? clock64() : 0;
    tmem_ld32(taddr + ch * 32, r);
    const int col0 = n0 + ch * 32;
    if (ch + 1 < CH1) load_side(col0 + 32, side_next);
    tmem_ld_wait();
    if (p.dbg) t_ld += clock64() - tl0;
    if (!(row_ok && col0 < p.N)) return;
    float* dst = p.C ? p.C + (size_t)row * p.ldc + col0 : nullptr;
    if (col0 + 32 <= p.N && fast) {
      // ---- fast path: 32 contiguous columns of one row per thread, all accesses 16-byte vectors
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (use_bias) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 b = *reinterpret_cast<const float4*>(bias_s + (ch - CH0) * 32 + q4 * 4);
          v[q4 * 4] += b.x; v[q4 * 4 + 1] += b.y; v[q4 * 4 + 2] += b.z; v[q4 * 4 + 3] += b.w;
        }
      }
      if (fl & EPI_SCALE_COLS) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (col0 + j < p.ep.scale_ncols) ? v[j] * p.ep.alpha : v[j];
      }
      if (use_res) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 b = side[q4];
          v[q4 * 4] += b.x; v[q4 * 4 + 1] += b.y; v[q4 * 4 + 2] += b.z; v[q4 * 4 + 3] += b.w;
        }
      }
      if (fl & EPI_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (mask_bf16) {
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          const float4 mv = side[q8];
          const uint32_t w[4] = {__float_as_uint(mv.x), __float_as_uint(mv.y), __float_as_uint(mv.z), __float_as_uint(mv.w)};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            // bf16 > 0  <=>  sign bit clear and magnitude non-zero
            uint32_t lo16 = w[h] & 0xFFFFu, hi16 = w[h] >> 16;
            if (!((lo16 & 0x8000u) == 0 && (lo16 & 0x7FFFu) != 0)) v[q8 * 8 + h * 2] = 0.f;
            if (!((hi16 & 0x8000u) == 0 && (hi16 & 0x7FFFu) != 0)) v[q8 * 8 + h * 2 + 1] = 0.f;
          }
        }
      } else if (mask_f32) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 b = side[q4];
          if (!(b.x > 0.f)) v[q4 * 4] = 0.f;
          if (!(b.y > 0.f)) v[q4 * 4 + 1] = 0.f;
          if (!(b.z > 0.f)) v[q4 * 4 + 2] = 0.f;
          if (!(b.w > 0.f)) v[q4 * 4 + 3] = 0.f;
        }
      }
      if (dst) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          float4 o = make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
          if (fl & EPI_ACCUMULATE) {
            float4 c = ld4(dst + q4 * 4);
            o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
            v[q4 * 4] = o.x; v[q4 * 4 + 1] = o.y; v[q4 * 4 + 2] = o.z; v[q4 * 4 + 3] = o.w;
          }
          st4(dst + q4 * 4, o);
        }
      }
      if (p.out_hi) {
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split_pair_bf16(v[2 * j], v[2 * j + 1], hw[j], lw[j]);
        uint4* hp = reinterpret_cast<uint4*>(p.out_hi + (size_t)row * p.ldo + col0);
        uint4* lp = reinterpret_cast<uint4*>(p.out_lo + (size_t)row * p.ldo + col0);
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          hp[q8] = make_uint4(hw[q8 * 4], hw[q8 * 4 + 1], hw[q8 * 4 + 2], hw[q8 * 4 + 3]);
          lp[q8] = make_uint4(lw[q8 * 4], lw[q8 * 4 + 1], lw[q8 * 4 + 2], lw[q8 * 4 + 3]);
        }
      }
    } else {
      // ---- edge / split-K path: element-wise
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < p.N) {
          float v = epilogue_apply(__uint_as_float(r[j]), row, col0 + j, p.ep, first_split);
          if (dst) {
            if (fl & EPI_ATOMIC) atomicAdd(dst + j, v);
            else if (fl & EPI_ACCUMULATE) { v += dst[j]; dst[j] = v; }
            else dst[j] = v;
          }
          if (p.out_hi) {
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            p.out_hi[(size_t)row * p.ldo + col0 + j] = h;
            p.out_lo[(size_t)row * p.ldo + col0 + j] = __float2bfloat16_rn(v - __bfloat162float(h));
          }
        }
      }
    }
  };

  float4 sa[8], sb[8];
  load_side(n0 + CH0 * 32, sa);
#pragma unroll 1
  for (int ch = CH0; ch < CH1; ch += 2) {
    chunk(ch, sa, sb);
    chunk(ch + 1, sb, sa);
  }
  if (p.dbg && lane == 0 && q == 0) atomicAdd(p.dbg + 3, (unsigned long long)t_ld);   // quadrant 0: tcgen05.ld issue -> data
}

// Quad epilogue (round 2, the default): same contract and bit-identical results as `epilogue_tile`, with 4x fewer cache
// lines per load / store instruction.  Why: with thread = row every 16-byte access of a warp touches 32 different lines, and
// the L1/LSU pipe spends one cycle per line - 256 cycles per 32 x 32 chunk and matrix, 8 k (fp32 or split output) to 16 k
// (plus an fp32 residual) cycles per 128 x 256 tile against 12.3 k cycles of MMA at K = 512.  Measured (gemm_check, stores
// removed): the K = 512 shapes run at 1160 - 1660 tensor-TFLOP/s without their stores and 815 - 1100 with them.  Staging
// through shared memory (`epilogue_tile_staged`) is no way out: the MMA operand reads and the TMA writes already take
// ~106 of the 128 B/clk of shared-memory bandwidth.  Here the accumulator is read in the 16x256b (mma fragment) layout - a
// quad of lanes holds 8 consecutive columns of a row - and ONE shuffle exchange between neighbouring lanes gives every
// lane 4 consecutive columns: a quad then covers 64 contiguous bytes of a row and a warp instruction 8 rows (8 lines).
//   lane = 4 rq + j,  b = j & 1, c = j >> 1;  after the exchange the lane holds, for h, s, k in {0, 1},
//   row 16 h + 8 s + rq,  columns 8 (2 k + b) + 4 c + {0 .. 3}   of the chunk.
// Split-K atomics, accumulate and unaligned operands take an element-wise path in the same fragment layout.
// SCFG >= 0: the epilogue is known when the kernel is compiled (SCFG_* bits: the EPI_* flags, which outputs exist; the
// host launches such a kernel only for aligned operands, N % 32 == 0 and no split-K) - the code below then folds into a
// straight line.  With run-time flags every (h, s, k) block carries ~10 branches, and one warp per scheduler pays the full
// branch latency for each: measured 2.7 k cycles per chunk with the stores removed.
constexpr int SCFG_C = 0x100, SCFG_SPLIT = 0x200;
template <int BN, int CH0 = 0, int CH1 = BN / 32, int SCFG = -1>
__device__ __forceinline__ void epilogue_tile_quad(const TcParams& p, uint32_t taddr, int m0, int n0, int q, int lane,
                                                   bool first_split_rt, uint8_t* wbuf) {
  static_assert(((CH1 - CH0) & 1) == 0, "chunk range must be even (the loop is unrolled by two)");
  constexpr bool ST = SCFG >= 0;
  const int fl = ST ? (SCFG & 0xff) : p.ep.flags;
  const bool first_split = ST ? true : first_split_rt;
  const bool has_c = ST ? ((SCFG & SCFG_C) != 0) : (p.C != nullptr);
  const bool has_split = ST ? ((SCFG & SCFG_SPLIT) != 0) : (p.out_hi != nullptr);
  const bool aligned = ST ? true : (!p.C || (((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.ldc & 3) == 0))) &&
                       (!p.out_hi || (((p.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out_hi) & 15) == 0) &&
                                      ((reinterpret_cast<uintptr_t>(p.out_lo) & 15) == 0))) &&
                       (!(fl & EPI_BIAS) || ((reinterpret_cast<uintptr_t>(p.ep.bias) & 15) == 0)) &&
                       (!(fl & (EPI_RESIDUAL | EPI_RELU_MASK)) ||
                        (((p.ep.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.ep.residual) & 15) == 0)));
  const bool fast = aligned && !(fl & (EPI_ATOMIC | EPI_ACCUMULATE));
  const bool use_bias = fast && first_split && (fl & EPI_BIAS);
  const bool use_res = first_split && (fl & EPI_RESIDUAL);
  const bool mask_bf16 = (fl & EPI_RELU_MASK) && (fl & EPI_MASK_BF16);
  const bool mask_f32 = (fl & EPI_RELU_MASK) && !(fl & EPI_MASK_BF16);
  const int rq = lane >> 2, j = lane & 3, b = j & 1, c = j >> 1;
  const int row0 = m0 + q * 32 + rq;                 // + 16 h + 8 s
  const int cofs = 8 * b + 4 * c;                    // + 16 k
  float* bias_s = reinterpret_cast<float*>(wbuf);

  if (use_bias) {
    __syncwarp();
    for (int i = lane * 4; i < (CH1 - CH0) * 32; i += 128) {
      const int col = n0 + CH0 * 32 + i;
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col + 4 <= p.N) bv = __ldg(reinterpret_cast<const float4*>(p.ep.bias + col));
      *reinterpret_cast<float4*>(bias_s + i) = bv;
    }
    __syncwarp();
  }
  // side inputs of one chunk in the post-exchange layout: index (2 h + s) * 2 + k; bf16 masks use .x / .y
  auto load_side = [&](int col0, float4 (&sd)[8]) {
    if (!ST && (col0 + 32 > p.N || !fast)) return;
    if (use_res || mask_f32) {
#pragma unroll
      for (int hs = 0; hs < 4; ++hs) {
        const int row = row0 + 8 * hs;
        if (row < p.M) {
          const float* rp = p.ep.residual + (size_t)row * p.ep.ldr + col0 + cofs;
          sd[hs * 2] = __ldg(reinterpret_cast<const float4*>(rp));
          sd[hs * 2 + 1] = __ldg(reinterpret_cast<const float4*>(rp + 16));
        }
      }
    } else if (mask_bf16) {
#pragma unroll
      for (int hs = 0; hs < 4; ++hs) {
        const int row = row0 + 8 * hs;
        if (row < p.M) {
          const __nv_bfloat16* mp = reinterpret_cast<const __nv_bfloat16*>(p.ep.residual) + (size_t)row * p.ep.ldr + col0 + cofs;
          const uint2 m0v = __ldg(reinterpret_cast<const uint2*>(mp)), m1v = __ldg(reinterpret_cast<const uint2*>(mp + 16));
          sd[hs * 2].x = __uint_as_float(m0v.x); sd[hs * 2].y = __uint_as_float(m0v.y);
          sd[hs * 2 + 1].x = __uint_as_float(m1v.x); sd[hs * 2 + 1].y = __uint_as_float(m1v.y);
        }
      }
    }
  };
  long long t_ld = 0;
  auto chunk = [&](int ch, float4 (&side)[8], float4 (&side_next)[8], int dist) {
    uint32_t a[2][16];
    const long long tl0 = p.dbg ? clock64() : 0;
    tmem_ld_16x256b_x4(taddr + ch * 32, a[0]);
    tmem_ld_16x256b_x4(taddr + (16u << 16) + ch * 32, a[1]);
    const int col0 = n0 + ch * 32;
    if (ch + dist < CH1) load_side(col0 + 32 * dist, side_next);
    tmem_ld_wait();
    if (p.dbg) t_ld += clock64() - tl0;
    if (col0 >= p.N) return;
    if (!ST && (col0 + 32 > p.N || !fast)) {
      // ---- chunk partial in N, split-K atomics, accumulate, unaligned operands: element-wise from the fragment layout
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int row = row0 + 16 * h + 8 * ((i >> 1) & 1), col = col0 + 8 * (i >> 2) + 2 * j + (i & 1);
          if (row < p.M && col < p.N) {
            float v = epilogue_apply(__uint_as_float(a[h][i]), row, col, p.ep, first_split);
            if (p.C) {
              float* dst = p.C + (size_t)row * p.ldc + col;
              if (fl & EPI_ATOMIC) atomicAdd(dst, v);
              else if (fl & EPI_ACCUMULATE) { v += *dst; *dst = v; }
              else *dst = v;
            }
            if (p.out_hi) {
              const __nv_bfloat16 hv = __float2bfloat16_rn(v);
              p.out_hi[(size_t)row * p.ldo + col] = hv;
              p.out_lo[(size_t)row * p.ldo + col] = __float2bfloat16_rn(v - __bfloat162float(hv));
            }
          }
        }
      return;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          // even / odd 8-column group of this pair of groups: the lane keeps group 2 k + b and trades the other one
          const uint32_t e0 = a[h][8 * k + 2 * s], e1 = a[h][8 * k + 2 * s + 1];
          const uint32_t o0 = a[h][8 * k + 4 + 2 * s], o1 = a[h][8 * k + 4 + 2 * s + 1];
          const uint32_t g0 = __shfl_xor_sync(0xffffffffu, b ? e0 : o0, 1);
          const uint32_t g1 = __shfl_xor_sync(0xffffffffu, b ? e1 : o1, 1);
          float v0 = __uint_as_float(b ? g0 : e0), v1 = __uint_as_float(b ? g1 : e1);
          float v2 = __uint_as_float(b ? o0 : g0), v3 = __uint_as_float(b ? o1 : g1);
          const int row = row0 + 16 * h + 8 * s;
          const bool row_ok = row < p.M;                 // (stores are predicated: no branch in the block)
          const int col = col0 + cofs + 16 * k;
          const float4 sdv = side[(2 * h + s) * 2 + k];
          if (use_bias) {
            const float4 bv = *reinterpret_cast<const float4*>(bias_s + (ch - CH0) * 32 + cofs + 16 * k);
            v0 += bv.x; v1 += bv.y; v2 += bv.z; v3 += bv.w;
          }
          if (fl & EPI_SCALE_COLS) {
            v0 = (col < p.ep.scale_ncols) ? v0 * p.ep.alpha : v0;
            v1 = (col + 1 < p.ep.scale_ncols) ? v1 * p.ep.alpha : v1;
            v2 = (col + 2 < p.ep.scale_ncols) ? v2 * p.ep.alpha : v2;
            v3 = (col + 3 < p.ep.scale_ncols) ? v3 * p.ep.alpha : v3;
          }
          if (use_res) { v0 += sdv.x; v1 += sdv.y; v2 += sdv.z; v3 += sdv.w; }
          if (fl & EPI_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
          if (mask_f32) {
            if (!(sdv.x > 0.f)) v0 = 0.f;
            if (!(sdv.y > 0.f)) v1 = 0.f;
            if (!(sdv.z > 0.f)) v2 = 0.f;
            if (!(sdv.w > 0.f)) v3 = 0.f;
          } else if (mask_bf16) {
            // bf16 > 0  <=>  sign bit clear and magnitude non-zero
            const uint32_t w0 = __float_as_uint(sdv.x), w1 = __float_as_uint(sdv.y);
            if (!((w0 & 0x8000u) == 0 && (w0 & 0x7FFFu) != 0)) v0 = 0.f;
            if (!((w0 & 0x80000000u) == 0 && (w0 & 0x7FFF0000u) != 0)) v1 = 0.f;
            if (!((w1 & 0x8000u) == 0 && (w1 & 0x7FFFu) != 0)) v2 = 0.f;
            if (!((w1 & 0x80000000u) == 0 && (w1 & 0x7FFF0000u) != 0)) v3 = 0.f;
          }
          if (has_c && row_ok) st4(p.C + (size_t)row * p.ldc + col, make_float4(v0, v1, v2, v3));
          if (has_split) {
            uint2 hw, lw;
            split_pair_bf16(v0, v1, hw.x, lw.x);
            split_pair_bf16(v2, v3, hw.y, lw.y);
            if (row_ok) {
              *reinterpret_cast<uint2*>(p.out_hi + (size_t)row * p.ldo + col) = hw;
              *reinterpret_cast<uint2*>(p.out_lo + (size_t)row * p.ldo + col) = lw;
            }
          }
        }
  };

  constexpr bool DEEP = ST && ((SCFG & (EPI_RESIDUAL | EPI_RELU_MASK)) != 0) && (CH1 - CH0 >= 3);
  if constexpr (DEEP) {
    // side inputs requested TWO chunks ahead (three register sets, chunk loop fully unrolled): one chunk of lead does not
    // cover the load latency once the rest of the chunk is a straight line of ~230 instructions
    float4 sd[3][8];
    load_side(n0 + CH0 * 32, sd[0]);
    load_side(n0 + CH0 * 32 + 32, sd[1]);
#pragma unroll
    for (int ch = CH0; ch < CH1; ++ch) chunk(ch, sd[(ch - CH0) % 3], sd[(ch - CH0 + 2) % 3], 2);
  } else {
    float4 sa[8], sb[8];
    load_side(n0 + CH0 * 32, sa);
#pragma unroll 1
    for (int ch = CH0; ch < CH1; ch += 2) {
      chunk(ch, sa, sb, 1);
      chunk(ch + 1, sb, sa, 1);
    }
  }
  if (p.dbg && lane == 0 && q == 0) atomicAdd(p.dbg + 3, (unsigned long long)t_ld);   // quadrant 0: tcgen05.ld issue -> data
}

// Epilogue variant, chosen when the library is built (-DVQ_EPI_KIND=k; one variant per kernel keeps its register allocation
// clean - all three inlined behind a run-time switch cost 255 registers and spills): 0 = quad (default; the only one with
// compile-time epilogue configurations), 1 = staged through shared memory, 2 = thread-per-row.  Results are bit-identical; tools/gemm_check.py --libs compares builds.
#ifndef VQ_EPI_KIND
#define VQ_EPI_KIND 0
#endif

template <int BN, int CH0 = 0, int CH1 = BN / 32, int SCFG = -1>
__device__ __forceinline__ void epilogue_dispatch(const TcParams& p, uint32_t taddr, int m0, int n0, int q, int lane,
                                                  bool first_split, uint8_t* wbuf);

// Bytes of shared memory each epilogue warp owns as its transpose buffer (32 rows x 128 bytes).
constexpr int EPI_STAGE_BYTES = 4096;

// Staged epilogue (the default): same contract and bit-identical results as `epilogue_tile`, but every global access of a
// warp covers whole 128-byte lines.  A 32-column chunk goes TMEM -> registers (thread = row) -> the warp's swizzled
// shared-memory buffer -> registers again with 8 lanes per row (lane = 4 consecutive columns, 4 rows per instruction).
// The thread-per-row form issues 16-byte accesses to 32 different lines per instruction and serialises one global-load
// latency (bias / residual / mask) behind every tcgen05.ld: ncu on the K = 512 shapes showed 4 epilogue warps needing about
// twice the tile's MMA time (tensor pipe 48-51 % active).  Here the next chunk's tcgen05.ld and its residual / mask / bias
// loads are in flight while the current chunk is written out.  Chunks that are partial in N and the atomic / accumulate
// modes take the element-wise path of `epilogue_tile`.
template <int BN, int CH0 = 0, int CH1 = BN / 32>
__device__ __forceinline__ void epilogue_tile_staged(const TcParams& p, uint32_t taddr, int m0, int n0, int q, int lane,
                                                     bool first_split, uint8_t* stage) {
  const int fl = p.ep.flags;
  const bool aligned = (!p.C || (((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.ldc & 3) == 0))) &&
                       (!p.out_hi || (((p.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out_hi) & 15) == 0) &&
                                      ((reinterpret_cast<uintptr_t>(p.out_lo) & 15) == 0))) &&
                       (!(fl & EPI_BIAS) || ((reinterpret_cast<uintptr_t>(p.ep.bias) & 15) == 0)) &&
                       (!(fl & (EPI_RESIDUAL | EPI_RELU_MASK)) ||
                        (((p.ep.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.ep.residual) & 15) == 0)));
  const bool fast = aligned && !(fl & (EPI_ATOMIC | EPI_ACCUMULATE)) && ((n0 & 3) == 0);
  const bool use_bias = first_split && (fl & EPI_BIAS);
  const bool use_res = first_split && (fl & EPI_RESIDUAL);
  const bool mask_bf16 = (fl & EPI_RELU_MASK) && (fl & EPI_MASK_BF16);
  const bool mask_f32 = (fl & EPI_RELU_MASK) && !(fl & EPI_MASK_BF16);
  const int sr = lane >> 3, cg = lane & 7;          // transposed phase: sub-row and 16-byte column group of this lane
  const int row_t = m0 + q * 32 + sr;               // + 4 i
  const int row_own = m0 + q * 32 + lane;           // thread-per-row phase

  // side inputs of one chunk for the transposed phase: aux[i] = residual / fp32 mask (float4) or bf16 mask (x, y words)
  float4 aux[8], aux_n[8];
  float4 bia = make_float4(0.f, 0.f, 0.f, 0.f), bia_n = bia;
  auto load_side = [&](int col0, float4 (&a)[8], float4& b) {
    const int col = col0 + cg * 4;
    if (use_bias) b = __ldg(reinterpret_cast<const float4*>(p.ep.bias + col));
    if (use_res || mask_f32) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (row_t + 4 * i < p.M) a[i] = __ldg(reinterpret_cast<const float4*>(p.ep.residual + (size_t)(row_t + 4 * i) * p.ep.ldr + col));
    } else if (mask_bf16) {
      const __nv_bfloat16* mp = reinterpret_cast<const __nv_bfloat16*>(p.ep.residual);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (row_t + 4 * i < p.M) {
          const uint2 mv = __ldg(reinterpret_cast<const uint2*>(mp + (size_t)(row_t + 4 * i) * p.ep.ldr + col));
          a[i].x = __uint_as_float(mv.x);
          a[i].y = __uint_as_float(mv.y);
        }
    }
  };

  uint32_t r[32];
  tmem_ld32(taddr + CH0 * 32, r);
  if (fast && n0 + CH0 * 32 + 32 <= p.N) load_side(n0 + CH0 * 32, aux, bia);
#pragma unroll 1
  for (int ch = CH0; ch < CH1; ++ch) {
    const int col0 = n0 + ch * 32;
    tmem_ld_wait();
    if (col0 >= p.N) break;
    const bool full = fast && (col0 + 32 <= p.N);
    if (full) {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(stage + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
    } else if (row_own < p.M) {
      // ---- partial chunk / split-K / accumulate: element-wise
      float* dst = p.C ? p.C + (size_t)row_own * p.ldc + col0 : nullptr;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < p.N) {
          float v = epilogue_apply(__uint_as_float(r[j]), row_own, col0 + j, p.ep, first_split);
          if (dst) {
            if (fl & EPI_ATOMIC) atomicAdd(dst + j, v);
            else if (fl & EPI_ACCUMULATE) { v += dst[j]; dst[j] = v; }
            else dst[j] = v;
          }
          if (p.out_hi) {
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            p.out_hi[(size_t)row_own * p.ldo + col0 + j] = h;
            p.out_lo[(size_t)row_own * p.ldo + col0 + j] = __float2bfloat16_rn(v - __bfloat162float(h));
          }
        }
      }
    }
    // the accumulator registers are dead: fetch the next chunk while this one is written out
    const bool more = ch + 1 < CH1 && col0 + 32 < p.N;
    if (more) tmem_ld32(taddr + (ch + 1) * 32, r);
    if (!full) continue;
    __syncwarp();
    const bool next_full = more && (col0 + 64 <= p.N);
    if (next_full) load_side(col0 + 32, aux_n, bia_n);
    const int col = col0 + cg * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = 4 * i + sr;
      const float4 t = *reinterpret_cast<const float4*>(stage + rr * 128 + ((cg ^ (rr & 7)) << 4));
      const int row = row_t + 4 * i;
      if (row < p.M) {
        float v0 = t.x, v1 = t.y, v2 = t.z, v3 = t.w;
        if (use_bias) { v0 += bia.x; v1 += bia.y; v2 += bia.z; v3 += bia.w; }
        if (fl & EPI_SCALE_COLS) {
          v0 = (col < p.ep.scale_ncols) ? v0 * p.ep.alpha : v0;
          v1 = (col + 1 < p.ep.scale_ncols) ? v1 * p.ep.alpha : v1;
          v2 = (col + 2 < p.ep.scale_ncols) ? v2 * p.ep.alpha : v2;
          v3 = (col + 3 < p.ep.scale_ncols) ? v3 * p.ep.alpha : v3;
        }
        if (use_res) { v0 += aux[i].x; v1 += aux[i].y; v2 += aux[i].z; v3 += aux[i].w; }
        if (fl & EPI_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
        if (mask_f32) {
          if (!(aux[i].x > 0.f)) v0 = 0.f;
          if (!(aux[i].y > 0.f)) v1 = 0.f;
          if (!(aux[i].z > 0.f)) v2 = 0.f;
          if (!(aux[i].w > 0.f)) v3 = 0.f;
        } else if (mask_bf16) {
          // bf16 > 0  <=>  sign bit clear and magnitude non-zero
          const uint32_t w0 = __float_as_uint(aux[i].x), w1 = __float_as_uint(aux[i].y);
          if (!((w0 & 0x8000u) == 0 && (w0 & 0x7FFFu) != 0)) v0 = 0.f;
          if (!((w0 & 0x80000000u) == 0 && (w0 & 0x7FFF0000u) != 0)) v1 = 0.f;
          if (!((w1 & 0x8000u) == 0 && (w1 & 0x7FFFu) != 0)) v2 = 0.f;
          if (!((w1 & 0x80000000u) == 0 && (w1 & 0x7FFF0000u) != 0)) v3 = 0.f;
        }
        if (p.C) st4(p.C + (size_t)row * p.ldc + col, make_float4(v0, v1, v2, v3));
        if (p.out_hi) {
          uint2 hw, lw;
          split_pair_bf16(v0, v1, hw.x, lw.x);
          split_pair_bf16(v2, v3, hw.y, lw.y);
          *reinterpret_cast<uint2*>(p.out_hi + (size_t)row * p.ldo + col) = hw;
          *reinterpret_cast<uint2*>(p.out_lo + (size_t)row * p.ldo + col) = lw;
        }
      }
    }
    __syncwarp();                                   // the buffer is rewritten by the next chunk
    if (next_full) {
#pragma unroll
      for (int i = 0; i < 8; ++i) aux[i] = aux_n[i];
      bia = bia_n;
    }
  }
}

template <int BN, int CH0, int CH1, int SCFG>
__device__ __forceinline__ void epilogue_dispatch(const TcParams& p, uint32_t taddr, int m0, int n0, int q, int lane,
                                                  bool first_split, uint8_t* wbuf) {
#if VQ_EPI_KIND == 0
  epilogue_tile_quad<BN, CH0, CH1, SCFG>(p, taddr, m0, n0, q, lane, first_split, wbuf);
#elif VQ_EPI_KIND == 1
  epilogue_tile_staged<BN, CH0, CH1>(p, taddr, m0, n0, q, lane, first_split, wbuf);
#else
  epilogue_tile<BN, CH0, CH1>(p, taddr, m0, n0, q, lane, first_split, wbuf);
#endif
}

// host helpers shared by both kernels (defined in gemm_tc.cu)
int tc_get_encode();
int tc_make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows);
// CTA-pair kernel (gemm_tc2.cu)
extern int g_no_static_epilogue;      // test hook: 1 = always the run-time-flag epilogue
int tc_launch_pair(int a_kmajor, int b_kmajor, int passes, const CUtensorMap* maps, const TcParams& p, cudaStream_t st,
                   bool eight_epilogue_warps = false);

}  // namespace vqtc
