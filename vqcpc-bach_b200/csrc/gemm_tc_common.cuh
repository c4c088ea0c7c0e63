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
};

// Epilogue of one 128 x BN accumulator tile held in TMEM at `taddr` (lane quadrant q of this warp): 32-column chunks,
// tcgen05.ld -> fused bias / scale / residual / ReLU / mask -> fp32 and / or bf16-split stores (or atomics for split-K).
// CH0 / CH1: range of 32-column chunks this warp drains (all of them by default; the 8-epilogue-warp variant of the pair
// kernel gives each of the two warps of a lane quadrant one half of the columns).
template <int BN, int CH0 = 0, int CH1 = BN / 32>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, uint32_t taddr, int m0, int n0, int q, int lane, bool first_split) {
  const int row = m0 + q * 32 + lane;
#pragma unroll 1
for (int ch = CH0; ch < CH1; ++ch) {
  uint32_t r[32];
  tmem_ld32(taddr + ch * 32, r);
  tmem_ld_wait();
  const int col0 = n0 + ch * 32;
  if (row < p.M && col0 < p.N) {
    float* dst = p.C ? p.C + (size_t)row * p.ldc + col0 : nullptr;
    const int fl = p.ep.flags;
    const bool aligned = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                         (!p.out_hi || (((p.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out_hi) & 15) == 0) &&
                                        ((reinterpret_cast<uintptr_t>(p.out_lo) & 15) == 0))) &&
                         (!(fl & EPI_BIAS) || ((reinterpret_cast<uintptr_t>(p.ep.bias) & 15) == 0)) &&
                         (!(fl & (EPI_RESIDUAL | EPI_RELU_MASK)) ||
                          (((p.ep.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.ep.residual) & 15) == 0)));
    if (col0 + 32 <= p.N && aligned && !(fl & EPI_ATOMIC)) {
      // ---- fast path: 32 contiguous columns of one row per thread, all accesses 16-byte vectors
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (first_split && (fl & EPI_BIAS)) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          float4 b = __ldg(reinterpret_cast<const float4*>(p.ep.bias + col0) + q4);
          v[q4 * 4] += b.x; v[q4 * 4 + 1] += b.y; v[q4 * 4 + 2] += b.z; v[q4 * 4 + 3] += b.w;
        }
      }
      if (fl & EPI_SCALE_COLS) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (col0 + j < p.ep.scale_ncols) ? v[j] * p.ep.alpha : v[j];
      }
      if (first_split && (fl & EPI_RESIDUAL)) {
        const float4* rp = reinterpret_cast<const float4*>(p.ep.residual + (size_t)row * p.ep.ldr + col0);
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          float4 b = __ldg(rp + q4);
          v[q4 * 4] += b.x; v[q4 * 4 + 1] += b.y; v[q4 * 4 + 2] += b.z; v[q4 * 4 + 3] += b.w;
        }
      }
      if (fl & EPI_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (fl & EPI_RELU_MASK) {
        if (fl & EPI_MASK_BF16) {
          const uint4* mp = reinterpret_cast<const uint4*>(
              reinterpret_cast<const __nv_bfloat16*>(p.ep.residual) + (size_t)row * p.ep.ldr + col0);
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            uint4 mv = __ldg(mp + q8);
            const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              // bf16 > 0  <=>  sign bit clear and magnitude non-zero
              uint32_t lo16 = w[h] & 0xFFFFu, hi16 = w[h] >> 16;
              if (!((lo16 & 0x8000u) == 0 && (lo16 & 0x7FFFu) != 0)) v[q8 * 8 + h * 2] = 0.f;
              if (!((hi16 & 0x8000u) == 0 && (hi16 & 0x7FFFu) != 0)) v[q8 * 8 + h * 2 + 1] = 0.f;
            }
          }
        } else {
          const float4* mp = reinterpret_cast<const float4*>(p.ep.residual + (size_t)row * p.ep.ldr + col0);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            float4 b = __ldg(mp + q4);
            if (!(b.x > 0.f)) v[q4 * 4] = 0.f;
            if (!(b.y > 0.f)) v[q4 * 4 + 1] = 0.f;
            if (!(b.z > 0.f)) v[q4 * 4 + 2] = 0.f;
            if (!(b.w > 0.f)) v[q4 * 4 + 3] = 0.f;
          }
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
        for (int j = 0; j < 16; ++j) {
          __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
          __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
          __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
          hw[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lw[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
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
  }
}
}

// host helpers shared by both kernels (defined in gemm_tc.cu)
int tc_get_encode();
int tc_make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows);
// CTA-pair kernel (gemm_tc2.cu)
int tc_launch_pair(int a_kmajor, int b_kmajor, int passes, const CUtensorMap* maps, const TcParams& p, cudaStream_t st,
                   bool eight_epilogue_warps = false);

}  // namespace vqtc
