// CTA-pair variant of the tensor-core GEMM: tcgen05.mma.cta_group::2 on 256 x 256 x 64 tiles.
//
// Two CTAs of a cluster (same TPC) form one MMA: CTA r owns rows [r*128, r*128+128) of the 256-row tile (its half of
// the accumulator lives in its own TMEM) and columns [r*128, r*128+128) of the B tile; the tensor cores of the pair
// read both halves.  Per CTA and k-step this stages A (128 x 32) and HALF of B (128 x 32) - hi and lo - i.e. 32 KB for
// twice the MMA work of the 1-CTA kernel's 48 KB: operand bytes per MMA cycle drop from 62 to 21 B/clk/SM, which is what
// the short-K (K = 512) encoder GEMMs were starved by (ncu: 48 % tensor-pipe active, MMA warp waiting on TMA).
//
// Protocol (per cluster; rank 0 = leader):
//   full[s]   (leader, count 2) : leader producer arrive.expect_tx(bytes of BOTH CTAs) + peer producer remote arrive;
//                                 both producers' TMA loads (cta_group::2) complete_tx on the LEADER's barrier
//   empty[s]  (each CTA, count 1): tcgen05.commit.cta_group::2 multicast from the leader's MMA thread
//   tfull[a]  (each CTA, count 1): tcgen05.commit.cta_group::2 multicast when a tile's MMAs retire
//   tempty[a] (leader, count 8) : the 4 epilogue warps of both CTAs arrive (peer: remote arrive) after draining TMEM
#include "gemm_tc_common.cuh"

using namespace vqtc;

namespace {

constexpr int BN2 = 256;        // tile columns (128 per CTA of the B operand)
constexpr int BK2 = 64;        // 128-byte rows: SWIZZLE_128B for K-major tiles (measured faster than 64B/BK=32 here)
constexpr int HALF_N = BN2 / 2;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // .relaxed: a release arrive on a remote barrier costs MEMBAR.ALL.GPU + ERRBAR each time (ncu round 2: 14 % of the pair
  // kernel's stall samples - one per k-block on the peer producer, one per tile and epilogue warp), and neither hand-off
  // publishes generic-proxy data: TMA bytes are tracked by the barrier's tx count, TMEM reads are complete at
  // tcgen05.wait::ld and ordered by tcgen05.fence::before_thread_sync
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on the mbarrier of the LEADER CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_pair(bool a_kmajor, bool b_kmajor) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_kmajor ? 0u : 1u) << 15) | ((b_kmajor ? 0u : 1u) << 16) |
         ((uint32_t)(BN2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);      // M = 256 across the pair, N = 256
}

// EPW = number of epilogue warps per CTA:
//   4: 192 threads, warps 2..5, one per TMEM lane quadrant (= scheduler partition), all 256 columns each.
//   8: 384 threads = warpgroup 0 (warp 0 producer, warp 1 MMA issuer, two idle warps; setmaxnreg 56) + warpgroups 1, 2
//      (warps 4..11, setmaxnreg 224): warps w and w + 4 share a lane quadrant - and therefore a scheduler partition - and drain
//      128 columns each.  ncu on the K = 512 shapes shows the 4-warp epilogue issuing one instruction every ~7 cycles (one warp
//      per scheduler, dependent chains, ~350 instructions per 32 x 32 chunk) and needing 1.5 - 2x the tile's MMA time; a second
//      warp per scheduler hides those latencies.  Round 1 tried 8 warps in a 320-thread block: ten warps put three on two of the
//      partitions, which caps every thread at 168 registers - hence the warpgroup layout with setmaxnreg.
// SCFG: static epilogue configuration (see epilogue_tile_quad), -1 = run-time flags.
template <bool A_KMAJOR, bool B_KMAJOR, int PASSES, int STAGES, int EPW = 4, int SCFG = -1>
__global__ void __launch_bounds__(EPW == 8 ? 384 : 192, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, TcParams p) {
  constexpr int TILE_A_BYTES = BM * BK2 * 2;            // 8 KB  (own 128 rows)
  constexpr int TILE_B_BYTES = HALF_N * BK2 * 2;        // 8 KB  (own 128 of the 256 tile columns)
  constexpr int TMEM_COLS = NUM_ACC * BN2;              // 512
  constexpr int CHUNK_BYTES = BK2 * 128;                // one MN-major TMA box: [BK k-rows][64 mn]
  constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (TILE_A_BYTES + TILE_B_BYTES);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;     // EPW transpose buffers, one per epilogue warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + EPW * EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + NUM_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NUM_ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int EPI_WARP0 = EPW == 8 ? 4 : 2;           // first epilogue warp
  const int crank = (int)cluster_ctarank();             // 0 = leader
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int m_groups = (p.m_tiles + 1) / 2;             // pairs of 128-row tiles
  const int total_items = m_groups * p.n_tiles * p.ksplit;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 2);             // one arrive per producer of the pair (leader's carries the tx count)
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 2 * EPW);     // EPW epilogue warps x 2 CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    if constexpr (EPW == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters) {
        const int ks = item / (m_groups * p.n_tiles);
        const int rem = item % (m_groups * p.n_tiles);
        const int m0 = ((rem / p.n_tiles) * 2 + crank) * BM;                 // own 128 rows of the 256-row tile
        const int nb0 = (rem % p.n_tiles) * BN2 + crank * HALF_N;            // own 128 columns of the B tile
        const int kb0 = ks * p.kblocks_per_split;
        const int kb1 = min(p.kblocks_total, kb0 + p.kblocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          // completion goes to the LEADER's full barrier: clear the CTA-rank bit of the shared::cluster address
          const uint32_t fb_local = smem_u32(&full_bar[stage]);
          const uint32_t fb = mapa_u32(fb_local, 0);
          if (crank == 0) mbar_expect_tx(fb_local, 2 * STAGE_BYTES);         // bytes of both CTAs land on this barrier
          else mbar_arrive_cluster(fb);
          const uint32_t a_hi = smem_u32(smem + stage * STAGE_BYTES), b_hi = a_hi + TILE_A_BYTES;
          const uint32_t a_lo = b_hi + TILE_B_BYTES, b_lo = a_lo + TILE_A_BYTES;
          const int k0 = kb * BK2;
          if (A_KMAJOR) {
            tma_load_2d_pair(a_hi, &tmA_hi, fb, k0, m0);
            if (PASSES == 3) tma_load_2d_pair(a_lo, &tmA_lo, fb, k0, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) {
              tma_load_2d_pair(a_hi + c * CHUNK_BYTES, &tmA_hi, fb, m0 + c * 64, k0);
              if (PASSES == 3) tma_load_2d_pair(a_lo + c * CHUNK_BYTES, &tmA_lo, fb, m0 + c * 64, k0);
            }
          }
          if (B_KMAJOR) {
            tma_load_2d_pair(b_hi, &tmB_hi, fb, k0, nb0);
            if (PASSES == 3) tma_load_2d_pair(b_lo, &tmB_lo, fb, k0, nb0);
          } else {
#pragma unroll
            for (int c = 0; c < HALF_N / 64; ++c) {
              tma_load_2d_pair(b_hi + c * CHUNK_BYTES, &tmB_hi, fb, nb0 + c * 64, k0);
              if (PASSES == 3) tma_load_2d_pair(b_lo + c * CHUNK_BYTES, &tmB_lo, fb, nb0 + c * 64, k0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if constexpr (EPW == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (lane == 0 && crank == 0) {
      constexpr uint32_t idesc = make_idesc_pair(A_KMAJOR, B_KMAJOR);
      constexpr uint32_t A_STEP = A_KMAJOR ? (UMMA_K * 2) : (UMMA_K * 128);
      constexpr uint32_t B_STEP = B_KMAJOR ? (UMMA_K * 2) : (UMMA_K * 128);
      constexpr uint32_t A_HI_C = desc_hi32<BK2>(A_KMAJOR), B_HI_C = desc_hi32<BK2>(B_KMAJOR);
      constexpr uint32_t A_LO_C = desc_lo32_const<BK2>(A_KMAJOR), B_LO_C = desc_lo32_const<BK2>(B_KMAJOR);
      const uint32_t smem_base_u32 = smem_u32(smem) & 0x3FFFFu;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      long long t_tempty = 0, t_full = 0, t_all = clock64();
      for (int item = cluster_id; item < total_items; item += num_clusters) {
        const int ks = item / (m_groups * p.n_tiles);
        const int kb0 = ks * p.kblocks_per_split;
        const int kb1 = min(p.kblocks_total, kb0 + p.kblocks_per_split);
        long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
        if (p.dbg) t_tempty += clock64() - t0;
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN2);
        for (int kb = kb0; kb < kb1; ++kb) {
          t0 = p.dbg ? clock64() : 0;
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          if (p.dbg) t_full += clock64() - t0;
          tcgen05_fence_after();
          const uint32_t a_hi = smem_base_u32 + (uint32_t)(stage * STAGE_BYTES), b_hi = a_hi + TILE_A_BYTES;
          const uint32_t a_lo = b_hi + TILE_B_BYTES, b_lo = a_lo + TILE_A_BYTES;
          const uint32_t la_hi = A_LO_C | (a_hi >> 4), lb_hi = B_LO_C | (b_hi >> 4);
          const uint32_t la_lo = A_LO_C | (a_lo >> 4), lb_lo = B_LO_C | (b_lo >> 4);
#pragma unroll
          for (int k = 0; k < BK2 / UMMA_K; ++k) {
            const uint64_t da_hi = pack_desc(la_hi + k * (A_STEP >> 4), A_HI_C);
            const uint64_t db_hi = pack_desc(lb_hi + k * (B_STEP >> 4), B_HI_C);
            const uint32_t first = (kb > kb0 || k > 0) ? 1u : 0u;
            if (PASSES == 3) {
              const uint64_t da_lo = pack_desc(la_lo + k * (A_STEP >> 4), A_HI_C);
              const uint64_t db_lo = pack_desc(lb_lo + k * (B_STEP >> 4), B_HI_C);
              umma_bf16_pair(tmem_d, da_lo, db_hi, idesc, first);
              umma_bf16_pair(tmem_d, da_hi, db_lo, idesc, 1u);
              umma_bf16_pair(tmem_d, da_hi, db_hi, idesc, 1u);
            } else {
              umma_bf16_pair(tmem_d, da_hi, db_hi, idesc, first);
            }
          }
          umma_commit_pair(smem_u32(&empty_bar[stage]), 3);      // frees the slot in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(smem_u32(&tfull_bar[acc]), 3);          // both halves of the accumulator are complete
        if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1u; }
      }
      if (p.dbg) {   // [0] MMA-thread cycles in the tile loop, [1] of them waiting for a free accumulator, [2] for operands
        atomicAdd(p.dbg + 0, (unsigned long long)(clock64() - t_all));
        atomicAdd(p.dbg + 1, (unsigned long long)t_tempty);
        atomicAdd(p.dbg + 2, (unsigned long long)t_full);
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =============================== epilogue (both CTAs, own 128 rows) ===============================
    if constexpr (EPW == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long t_wait = 0, t_epi = 0;
    for (int item = cluster_id; item < total_items; item += num_clusters) {
      const int ks = item / (m_groups * p.n_tiles);
      const int rem = item % (m_groups * p.n_tiles);
      const int m0 = ((rem / p.n_tiles) * 2 + crank) * BM, n0 = (rem % p.n_tiles) * BN2;
      const bool first_split = (ks == 0);
      const long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      const long long t1 = p.dbg ? clock64() : 0;
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN2);
      uint8_t* stg = epi_stage + (warp - EPI_WARP0) * EPI_STAGE_BYTES;
      if constexpr (EPW == 8) {
        if (warp < 8) epilogue_dispatch<BN2, 0, BN2 / 64>(p, taddr, m0, n0, q, lane, first_split, stg);
        else epilogue_dispatch<BN2, BN2 / 64, BN2 / 32>(p, taddr, m0, n0, q, lane, first_split, stg);
      } else {
        epilogue_dispatch<BN2, 0, BN2 / 32, SCFG>(p, taddr, m0, n0, q, lane, first_split, stg);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));     // leader's barrier
      if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1u; }
      if (p.dbg) { t_wait += t1 - t0; t_epi += clock64() - t1; }
    }
    if (p.dbg && lane == 0) {   // [4 + q] epilogue-warp cycles waiting for an accumulator, [8 + q] draining it (per lane quadrant)
      atomicAdd(p.dbg + 4 + q, (unsigned long long)t_wait);
      atomicAdd(p.dbg + 8 + q, (unsigned long long)t_epi);
    }
  }

  else {
    if constexpr (EPW == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");     // idle warps 2, 3 of warpgroup 0
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

template <bool AK, bool BKM, int PASSES, int STAGES, int EPW = 4, int SCFG = -1>
int launch_pair(const CUtensorMap* maps, const TcParams& p, cudaStream_t st) {
  constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (BM * BK2 * 2 + HALF_N * BK2 * 2);
  constexpr int SMEM = STAGES * STAGE_BYTES + EPW * EPI_STAGE_BYTES + 1024 + 256;
  static_assert(SMEM <= 227 * 1024, "tile configuration exceeds shared memory");
  auto kern = gemm_tc2_kernel<AK, BKM, PASSES, STAGES, EPW, SCFG>;
  static bool done = false;
  if (!done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    done = true;
  }
  const int m_groups = (p.m_tiles + 1) / 2;
  const int items = m_groups * p.n_tiles * p.ksplit;
  const int clusters = std::max(1, std::min(items, vq_sm_count() / 2));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  cfg.blockDim = dim3(EPW == 8 ? 384 : 192);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], p);
  if (e != cudaSuccess) {
    vq_set_error("gemm_tc2_kernel: launch failed: %s", cudaGetErrorString(e));
    return VQ_ERR_CUDA;
  }
  return VQ_OK;
}

}  // namespace

namespace vqtc { int g_no_static_epilogue = 0; }

// Tile configuration 4 of vqcpc_gemm_bf16x3_ex: CTA pairs, 256 x 256 x 32 tiles.  `p` carries m_tiles in units of 128
// rows and n_tiles in units of 256 columns; maps[2..3] (B) must use boxes of 128 rows (K-major) / 64 x 32 (MN-major).
int vqtc::tc_launch_pair(int a_kmajor, int b_kmajor, int passes, const CUtensorMap* maps, const TcParams& p, cudaStream_t st,
                         bool eight_epilogue_warps) {
  if (eight_epilogue_warps && passes == 3) {
    if (a_kmajor && b_kmajor) return launch_pair<true, true, 3, 3, 8>(maps, p, st);
    if (a_kmajor && !b_kmajor) return launch_pair<true, false, 3, 3, 8>(maps, p, st);
    if (!a_kmajor && b_kmajor) return launch_pair<false, true, 3, 3, 8>(maps, p, st);
    return launch_pair<false, false, 3, 3, 8>(maps, p, st);
  }
  // static epilogue configurations of the encoder's forward / data-gradient GEMMs (everything else: run-time flags)
  if (passes == 3 && a_kmajor && p.ksplit == 1 && !g_no_static_epilogue) {
    const int fl = p.ep.flags;
    const bool aligned = (!p.C || (((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.ldc & 3) == 0))) &&
                         (!p.out_hi || (((p.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out_hi) & 15) == 0) &&
                                        ((reinterpret_cast<uintptr_t>(p.out_lo) & 15) == 0))) &&
                         (!(fl & EPI_BIAS) || ((reinterpret_cast<uintptr_t>(p.ep.bias) & 15) == 0)) &&
                         (!(fl & (EPI_RESIDUAL | EPI_RELU_MASK)) ||
                          (((p.ep.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.ep.residual) & 15) == 0)));
    if (aligned && (p.N % 32) == 0) {
      const int scfg = fl | (p.C ? SCFG_C : 0) | (p.out_hi ? SCFG_SPLIT : 0);
      if (b_kmajor) {
        if (scfg == (EPI_BIAS | EPI_SCALE_COLS | SCFG_C)) return launch_pair<true, true, 3, 3, 4, EPI_BIAS | EPI_SCALE_COLS | SCFG_C>(maps, p, st);
        if (scfg == (EPI_BIAS | EPI_RESIDUAL | SCFG_C)) return launch_pair<true, true, 3, 3, 4, EPI_BIAS | EPI_RESIDUAL | SCFG_C>(maps, p, st);
        if (scfg == (EPI_BIAS | EPI_RELU | SCFG_SPLIT)) return launch_pair<true, true, 3, 3, 4, EPI_BIAS | EPI_RELU | SCFG_SPLIT>(maps, p, st);
        if (scfg == (EPI_BIAS | SCFG_C)) return launch_pair<true, true, 3, 3, 4, EPI_BIAS | SCFG_C>(maps, p, st);
        if (scfg == (EPI_BIAS | EPI_RELU | SCFG_C)) return launch_pair<true, true, 3, 3, 4, EPI_BIAS | EPI_RELU | SCFG_C>(maps, p, st);      // dropout: FFN1
      } else {
        if (scfg == (EPI_RELU_MASK | EPI_MASK_BF16 | SCFG_SPLIT)) return launch_pair<true, false, 3, 3, 4, EPI_RELU_MASK | EPI_MASK_BF16 | SCFG_SPLIT>(maps, p, st);
        if (scfg == SCFG_C) return launch_pair<true, false, 3, 3, 4, SCFG_C>(maps, p, st);
        if (scfg == (EPI_RESIDUAL | SCFG_C)) return launch_pair<true, false, 3, 3, 4, EPI_RESIDUAL | SCFG_C>(maps, p, st);
        if (scfg == (EPI_SCALE_COLS | EPI_RELU_MASK | EPI_MASK_BF16 | SCFG_SPLIT))                                                         // dropout: FFN2 dgrad
          return launch_pair<true, false, 3, 3, 4, EPI_SCALE_COLS | EPI_RELU_MASK | EPI_MASK_BF16 | SCFG_SPLIT>(maps, p, st);
      }
    }
  }
  if (a_kmajor && b_kmajor) return passes == 3 ? launch_pair<true, true, 3, 3>(maps, p, st) : launch_pair<true, true, 1, 6>(maps, p, st);
  if (a_kmajor && !b_kmajor) return passes == 3 ? launch_pair<true, false, 3, 3>(maps, p, st) : launch_pair<true, false, 1, 6>(maps, p, st);
  if (!a_kmajor && b_kmajor) return passes == 3 ? launch_pair<false, true, 3, 3>(maps, p, st) : launch_pair<false, true, 1, 6>(maps, p, st);
  return passes == 3 ? launch_pair<false, false, 3, 3>(maps, p, st) : launch_pair<false, false, 1, 6>(maps, p, st);
}
