// Tensor-core GEMM for sm_100a: TMA-fed tcgen05.mma with TMEM accumulators, persistent and warp-specialised.
//
//   C[M,N] = epi( A . B ) with A, B given as PRE-SPLIT bf16 pairs (hi, lo), x ~= hi + lo (see vqcpc_split_bf16):
//     passes = 3 ("bf16x3", parity mode): D += A_lo.B_hi + A_hi.B_lo + A_hi.B_hi   (fp32 accumulation in TMEM;
//                 drops only the lo.lo term: relative error ~2^-16 per product, measured 7e-6 end to end)
//     passes = 1 ("bf16",  fast mode)   : D += A_hi.B_hi
//   Operand layouts (same convention as vqcpc_gemm_f32):
//     a_kmajor=1: A(m,k)=A[m*lda+k]   a_kmajor=0: A(m,k)=A[k*lda+m]   (tcgen05 "MN-major", used by weight gradients)
//     b_kmajor=1: B(k,n)=B[n*ldb+k]   b_kmajor=0: B(k,n)=B[k*ldb+n]   (data gradients read the weight untransposed)
//
// CTA = 192 threads: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner, warps 2..5 = epilogue
// (TMEM -> registers -> fused bias/scale/residual/ReLU/mask -> global).  Tile 128 x 128 x 64, kStages-deep smem ring
// (full/empty mbarriers), two TMEM accumulator stages (tmem_full/tmem_empty) so the epilogue of tile i overlaps the
// main loop of tile i+1.  Grid = min(#work items, #SMs); work item = (m tile, n tile, k split).
// Split-K partials are reduced with fp32 atomics (red.global.add) into C.
#include <cuda.h>
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace {

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

// CL = cluster size along M (1, 2 or 4): the CL CTAs of a cluster work on CL consecutive m-tiles of the SAME n-tile and
// k-range in lock step; each loads its own A tile and 1/CL of the B tile, which it TMA-multicasts to all of them.
// L2 -> SM operand traffic per tile drops from A + B to A + B/CL (B is the larger tile at BN = 256): the 1-CTA kernel
// was bound by L2 bandwidth (~17 TB/s demanded at full tensor rate), not by the tensor pipe.
template <bool A_KMAJOR, bool B_KMAJOR, int PASSES, int BN, int BK, int STAGES, int CL>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, TcParams p) {
  constexpr int TILE_A_BYTES = BM * BK * 2;
  constexpr int TILE_B_BYTES = BN * BK * 2;
  constexpr int TMEM_COLS = NUM_ACC * BN;                  // 256 or 512 (power of two)
  constexpr int CHUNK_BYTES = BK * 128;                    // one MN-major TMA box: [BK k-rows][64 mn]
  constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (TILE_A_BYTES + TILE_B_BYTES);
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + NUM_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NUM_ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work items are per CLUSTER: (k split, group of CL m-tiles, n tile)
  const int m_groups = (p.m_tiles + CL - 1) / CL;
  const int total_items = m_groups * p.n_tiles * p.ksplit;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), CL);       // every CTA of the cluster must have drained the slot
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 4);     // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                   // peers' barriers are initialised before anything targets them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters) {
        const int ks = item / (m_groups * p.n_tiles);
        const int rem = item % (m_groups * p.n_tiles);
        const int m0 = ((rem / p.n_tiles) * CL + crank) * BM, n0 = (rem % p.n_tiles) * BN;
        const int kb0 = ks * p.kblocks_per_split;
        const int kb1 = min(p.kblocks_total, kb0 + p.kblocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, STAGE_BYTES);
          uint8_t* st = smem + stage * STAGE_BYTES;
          const uint32_t a_hi = smem_u32(st), b_hi = a_hi + TILE_A_BYTES;
          const uint32_t a_lo = b_hi + TILE_B_BYTES, b_lo = a_lo + TILE_A_BYTES;
          const int k0 = kb * BK;
          if (A_KMAJOR) {
            tma_load_2d(a_hi, &tmA_hi, fb, k0, m0);
            if (PASSES == 3) tma_load_2d(a_lo, &tmA_lo, fb, k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) {
              tma_load_2d(a_hi + c * CHUNK_BYTES, &tmA_hi, fb, m0 + c * 64, k0);
              if (PASSES == 3) tma_load_2d(a_lo + c * CHUNK_BYTES, &tmA_lo, fb, m0 + c * 64, k0);
            }
          }
          if (CL == 1) {
            if (B_KMAJOR) {
              tma_load_2d(b_hi, &tmB_hi, fb, k0, n0);
              if (PASSES == 3) tma_load_2d(b_lo, &tmB_lo, fb, k0, n0);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c) {
                tma_load_2d(b_hi + c * CHUNK_BYTES, &tmB_hi, fb, n0 + c * 64, k0);
                if (PASSES == 3) tma_load_2d(b_lo + c * CHUNK_BYTES, &tmB_lo, fb, n0 + c * 64, k0);
              }
            }
          } else {
            // this CTA fetches slice `crank` of the B tile (BN / CL rows of n) and multicasts it to the whole cluster
            constexpr int SL = BN / CL;
            if (B_KMAJOR) {
              const uint32_t off = (uint32_t)(crank * SL * BK * 2);
              tma_load_2d_mc(b_hi + off, &tmB_hi, fb, k0, n0 + crank * SL, CMASK);
              if (PASSES == 3) tma_load_2d_mc(b_lo + off, &tmB_lo, fb, k0, n0 + crank * SL, CMASK);
            } else {
#pragma unroll
              for (int c = 0; c < SL / 64; ++c) {
                const int cc = crank * (SL / 64) + c;
                tma_load_2d_mc(b_hi + cc * CHUNK_BYTES, &tmB_hi, fb, n0 + cc * 64, k0, CMASK);
                if (PASSES == 3) tma_load_2d_mc(b_lo + cc * CHUNK_BYTES, &tmB_lo, fb, n0 + cc * 64, k0, CMASK);
              }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>(A_KMAJOR, B_KMAJOR);
      constexpr uint32_t A_STEP = A_KMAJOR ? (UMMA_K * 2) : (UMMA_K * 128);   // bytes per K=16 step inside a stage
      constexpr uint32_t B_STEP = B_KMAJOR ? (UMMA_K * 2) : (UMMA_K * 128);
      constexpr uint32_t A_HI_C = desc_hi32<BK>(A_KMAJOR), B_HI_C = desc_hi32<BK>(B_KMAJOR);
      constexpr uint32_t A_LO_C = desc_lo32_const<BK>(A_KMAJOR), B_LO_C = desc_lo32_const<BK>(B_KMAJOR);
      // (shared-window addresses carry the cluster CTA id in their high bits: keep the 18 address bits only)
      const uint32_t smem_base_u32 = smem_u32(smem) & 0x3FFFFu;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters) {
        const int ks = item / (m_groups * p.n_tiles);
        const int kb0 = ks * p.kblocks_per_split;
        const int kb1 = min(p.kblocks_total, kb0 + p.kblocks_per_split);
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tcgen05_fence_after();
          const uint32_t a_hi = smem_base_u32 + (uint32_t)(stage * STAGE_BYTES), b_hi = a_hi + TILE_A_BYTES;
          const uint32_t a_lo = b_hi + TILE_B_BYTES, b_lo = a_lo + TILE_A_BYTES;
          // low descriptor words at k = 0 (start address in 16-byte units; smem addresses are < 256 KB)
          const uint32_t la_hi = A_LO_C | (a_hi >> 4), lb_hi = B_LO_C | (b_hi >> 4);
          const uint32_t la_lo = A_LO_C | (a_lo >> 4), lb_lo = B_LO_C | (b_lo >> 4);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da_hi = pack_desc(la_hi + k * (A_STEP >> 4), A_HI_C);
            const uint64_t db_hi = pack_desc(lb_hi + k * (B_STEP >> 4), B_HI_C);
            const uint32_t first = (kb > kb0 || k > 0) ? 1u : 0u;
            if (PASSES == 3) {
              const uint64_t da_lo = pack_desc(la_lo + k * (A_STEP >> 4), A_HI_C);
              const uint64_t db_lo = pack_desc(lb_lo + k * (B_STEP >> 4), B_HI_C);
              umma_bf16(tmem_d, da_lo, db_hi, idesc, first);
              umma_bf16(tmem_d, da_hi, db_lo, idesc, 1u);
              umma_bf16(tmem_d, da_hi, db_hi, idesc, 1u);
            } else {
              umma_bf16(tmem_d, da_hi, db_hi, idesc, first);
            }
          }
          if (CL == 1) umma_commit(smem_u32(&empty_bar[stage]));   // frees the smem slot when these MMAs retire
          else umma_commit_mc(smem_u32(&empty_bar[stage]), CMASK);  // ... in every CTA of the cluster (they all write it)
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&tfull_bar[acc]));           // accumulator complete -> epilogue
        if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // =============================== epilogue (warps 2..5) ===============================
    const int q = warp & 3;                               // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = cluster_id; item < total_items; item += num_clusters) {
      const int ks = item / (m_groups * p.n_tiles);
      const int rem = item % (m_groups * p.n_tiles);
      const int m0 = ((rem / p.n_tiles) * CL + crank) * BM, n0 = (rem % p.n_tiles) * BN;
      const bool first_split = (ks == 0);
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tcgen05_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
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
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                   // nobody exits while a peer may still multicast into / signal it
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int get_encode() {
  if (g_encode) return VQ_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    vq_set_error("gemm_bf16x3: cuTensorMapEncodeTiled driver entry point unavailable (%s)", cudaGetErrorString(e));
    return VQ_ERR_DRIVER;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return VQ_OK;
}

// 2-D bf16 tensor map over a row-major [rows][cols] matrix (leading dimension ld elements), 128B swizzle.
// box = 64 columns (128 bytes) x box_rows rows.
int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
             uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;   // span = box row bytes
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vq_set_error("gemm_bf16x3: cuTensorMapEncodeTiled failed (CUresult %d; rows=%llu cols=%llu ld=%llu)", (int)r,
                 (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    return VQ_ERR_DRIVER;
  }
  return VQ_OK;
}

template <bool AK, bool BKM, int PASSES, int BN, int BK, int STAGES, int CL>
int launch_cfg(const CUtensorMap* maps, const TcParams& p, cudaStream_t st) {
  constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (BM * BK * 2 + BN * BK * 2);
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(SMEM <= 227 * 1024, "tile configuration exceeds shared memory");
  auto kern = gemm_tc_kernel<AK, BKM, PASSES, BN, BK, STAGES, CL>;
  static bool done = false;
  if (!done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    done = true;
  }
  const int m_groups = (p.m_tiles + CL - 1) / CL;
  const int items = m_groups * p.n_tiles * p.ksplit;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent grid = the number of clusters that can be co-resident (GPC boundaries can strand SMs for CL = 4)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cfg.gridDim = dim3((unsigned)(vq_sm_count() / CL * CL));
    int n = 0;
    if (CL > 1 && cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) max_clusters = n;
    else max_clusters = vq_sm_count() / CL;
    (void)cudaGetLastError();
  }
  const int clusters = std::max(1, std::min(items, max_clusters));
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], p);
  if (e != cudaSuccess) {
    vq_set_error("gemm_tc_kernel: launch failed: %s", cudaGetErrorString(e));
    return VQ_ERR_CUDA;
  }
  return VQ_OK;
}

// tile configuration ids: 0 = 128x128x64 (3 / 6 stages, no cluster), 1 = 128x256x32 (4 / 8 stages, no cluster),
//                         2 = 128x256x32 with the B tile multicast over a cluster of 2 CTAs along M, 3 = cluster of 4
int g_force_cfg = -1;
inline int pick_cfg(int M, int N) {
  if (g_force_cfg >= 0) return (N < 256 && g_force_cfg > 0) ? 0 : g_force_cfg;
  if (N < 256) return 0;
  return M >= 4 * BM ? 3 : (M >= 2 * BM ? 2 : 1);
}

template <bool AK, bool BKM>
int launch_tc(int cfg, int passes, const CUtensorMap* maps, const TcParams& p, cudaStream_t st) {
  if (cfg == 3) {
    if (passes == 3) return launch_cfg<AK, BKM, 3, 256, 32, 4, 4>(maps, p, st);
    return launch_cfg<AK, BKM, 1, 256, 32, 8, 4>(maps, p, st);
  }
  if (cfg == 2) {
    if (passes == 3) return launch_cfg<AK, BKM, 3, 256, 32, 4, 2>(maps, p, st);
    return launch_cfg<AK, BKM, 1, 256, 32, 8, 2>(maps, p, st);
  }
  if (cfg == 1) {
    if (passes == 3) return launch_cfg<AK, BKM, 3, 256, 32, 4, 1>(maps, p, st);
    return launch_cfg<AK, BKM, 1, 256, 32, 8, 1>(maps, p, st);
  }
  if (passes == 3) return launch_cfg<AK, BKM, 3, 128, 64, 3, 1>(maps, p, st);
  return launch_cfg<AK, BKM, 1, 128, 64, 6, 1>(maps, p, st);
}

}  // namespace

// Extended form: additionally (or instead of C, which may then be NULL) writes the result pre-split into bf16
// (out_hi, out_lo; leading dimension ldo) so that the next GEMM consumes it without a separate split pass.
// `residual` is void*: fp32 for VQ_EPI_RESIDUAL / VQ_EPI_RELU_MASK, bf16 when VQ_EPI_MASK_BF16 is also set.
VQ_API int vqcpc_gemm_bf16x3_ex(int a_kmajor, int b_kmajor, int M, int N, int K, const void* A_hi, const void* A_lo, int lda,
                                const void* B_hi, const void* B_lo, int ldb, float* C, int ldc, void* out_hi, void* out_lo,
                                int ldo, const float* bias, const void* residual_v, int ldr, int flags, float alpha,
                                int scale_ncols, int passes, int ksplit, void* stream) {
  const float* residual = reinterpret_cast<const float*>(residual_v);
  VQ_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm_bf16x3: bad shape M=%d N=%d K=%d", M, N, K);
  VQ_CHECK_ARG(passes == 1 || passes == 3, "gemm_bf16x3: passes must be 1 or 3");
  VQ_CHECK_ARG(A_hi && B_hi && (C || out_hi) && (passes == 1 || (A_lo && B_lo)), "gemm_bf16x3: null operand");
  VQ_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr), "gemm_bf16x3: out_hi/out_lo must be given together");
  VQ_CHECK_ARG(!(flags & EPI_MASK_BF16) || (flags & EPI_RELU_MASK), "gemm_bf16x3: MASK_BF16 needs RELU_MASK");
  VQ_CHECK_ARG(!((flags & EPI_ACCUMULATE) && (flags & EPI_RELU_MASK)), "gemm_bf16x3: ACCUMULATE cannot be combined with RELU_MASK");
  VQ_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm_bf16x3: leading dimensions must be multiples of 8 bf16 (lda=%d ldb=%d)", lda, ldb);
  VQ_CHECK_ARG(((uintptr_t)A_hi & 15) == 0 && ((uintptr_t)B_hi & 15) == 0 && ((uintptr_t)A_lo & 15) == 0 && ((uintptr_t)B_lo & 15) == 0,
               "gemm_bf16x3: operands must be 16-byte aligned");
  VQ_CHECK_ARG(!(flags & (EPI_RESIDUAL | EPI_RELU_MASK)) || residual, "gemm_bf16x3: residual/mask flag without pointer");
  VQ_CHECK_ARG(!(flags & EPI_BIAS) || bias, "gemm_bf16x3: bias flag without pointer");
  if (M == 0) return VQ_OK;
  int rc = get_encode();
  if (rc != VQ_OK) return rc;

  TcParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(out_hi); p.out_lo = reinterpret_cast<__nv_bfloat16*>(out_lo); p.ldo = ldo;
  p.ep.bias = bias; p.ep.residual = residual; p.ep.ldr = ldr; p.ep.flags = flags; p.ep.alpha = alpha; p.ep.scale_ncols = scale_ncols;
  const int cfg = pick_cfg(M, N);
  const int BN = cfg >= 1 ? 256 : 128, BK = cfg >= 1 ? 32 : 64;
  const int CLh = cfg == 3 ? 4 : (cfg == 2 ? 2 : 1);          // cluster size: B K-major boxes cover BN / CL rows
  p.m_tiles = (int)vq_cdiv(M, BM);
  p.n_tiles = (int)vq_cdiv(N, BN);
  p.kblocks_total = (int)vq_cdiv(K, BK);
  if (ksplit < 1) ksplit = 1;
  p.kblocks_per_split = (int)vq_cdiv(p.kblocks_total, ksplit);
  p.ksplit = (int)vq_cdiv(p.kblocks_total, p.kblocks_per_split);
  cudaStream_t st = vq_stream(stream);
  if (p.ksplit > 1) {
    VQ_CHECK_ARG(!(flags & (EPI_RELU | EPI_RELU_MASK)) && !out_hi && C, "gemm_bf16x3: split-K cannot fuse a non-linear epilogue or split outputs");
    p.ep.flags |= EPI_ATOMIC;
    if (!(flags & EPI_ACCUMULATE))
      VQ_CUDA_CALL(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
  }

  CUtensorMap maps[4];
  const void* alo = passes == 3 ? A_lo : A_hi;
  const void* blo = passes == 3 ? B_lo : B_hi;
  if (a_kmajor) {   // [M rows][K cols]: box = BK cols x 128 rows
    if ((rc = make_map(&maps[0], A_hi, M, K, lda, BK, BM)) != VQ_OK) return rc;
    if ((rc = make_map(&maps[1], alo, M, K, lda, BK, BM)) != VQ_OK) return rc;
  } else {          // [K rows][M cols]: box = 64 cols x BK rows
    if ((rc = make_map(&maps[0], A_hi, K, M, lda, 64, BK)) != VQ_OK) return rc;
    if ((rc = make_map(&maps[1], alo, K, M, lda, 64, BK)) != VQ_OK) return rc;
  }
  if (b_kmajor) {   // [N rows][K cols]
    if ((rc = make_map(&maps[2], B_hi, N, K, ldb, BK, BN / CLh)) != VQ_OK) return rc;
    if ((rc = make_map(&maps[3], blo, N, K, ldb, BK, BN / CLh)) != VQ_OK) return rc;
  } else {          // [K rows][N cols]
    if ((rc = make_map(&maps[2], B_hi, K, N, ldb, 64, BK)) != VQ_OK) return rc;
    if ((rc = make_map(&maps[3], blo, K, N, ldb, 64, BK)) != VQ_OK) return rc;
  }

  if (a_kmajor && b_kmajor) return launch_tc<true, true>(cfg, passes, maps, p, st);
  if (a_kmajor && !b_kmajor) return launch_tc<true, false>(cfg, passes, maps, p, st);
  if (!a_kmajor && b_kmajor) return launch_tc<false, true>(cfg, passes, maps, p, st);
  return launch_tc<false, false>(cfg, passes, maps, p, st);
}

// Testing / tuning hook: force a tile configuration (0 = 128x128x64, 1 = 128x256x32, 2 / 3 = 128x256x32 with the B tile
// multicast over a cluster of 2 / 4 CTAs, -1 = automatic).
VQ_API int vqcpc_gemm_set_tile_config(int cfg) {
  VQ_CHECK_ARG(cfg >= -1 && cfg <= 3, "gemm_set_tile_config: cfg must be -1 .. 3");
  g_force_cfg = cfg;
  return VQ_OK;
}

VQ_API int vqcpc_gemm_bf16x3(int a_kmajor, int b_kmajor, int M, int N, int K, const void* A_hi, const void* A_lo, int lda,
                             const void* B_hi, const void* B_lo, int ldb, float* C, int ldc, const float* bias,
                             const float* residual, int ldr, int flags, float alpha, int scale_ncols, int passes, int ksplit,
                             void* stream) {
  VQ_CHECK_ARG(C != nullptr, "gemm_bf16x3: null output");
  return vqcpc_gemm_bf16x3_ex(a_kmajor, b_kmajor, M, N, K, A_hi, A_lo, lda, B_hi, B_lo, ldb, C, ldc, nullptr, nullptr, 0, bias,
                              residual, ldr, flags, alpha, scale_ncols, passes, ksplit, stream);
}
