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

#include "gemm_tc_common.cuh"

using namespace vqtc;

namespace {

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
  uint8_t* epi_stage = smem + STAGES * STAGE_BYTES;     // 4 transpose buffers, one per epilogue warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + 4 * EPI_STAGE_BYTES);
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
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      uint8_t* stg = epi_stage + (warp - 2) * EPI_STAGE_BYTES;
      epilogue_dispatch<BN>(p, taddr, m0, n0, q, lane, first_split, stg);
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

}  // namespace

int vqtc::tc_get_encode() {
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
int vqtc::tc_make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
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

namespace {

template <bool AK, bool BKM, int PASSES, int BN, int BK, int STAGES, int CL>
int launch_cfg(const CUtensorMap* maps, const TcParams& p, cudaStream_t st) {
  constexpr int STAGE_BYTES = (PASSES == 3 ? 2 : 1) * (BM * BK * 2 + BN * BK * 2);
  constexpr int SMEM = STAGES * STAGE_BYTES + 4 * EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
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
//                         2 = 128x256x32 with the B tile multicast over a cluster of 2 CTAs along M, 3 = cluster of 4,
//                         4 = CTA pairs (tcgen05.mma.cta_group::2), 256x256x32 tiles, 6 / 12 stages (gemm_tc2.cu)
int g_force_cfg = -1;
unsigned long long* g_dbg = nullptr;
int g_exp = 0;
int g_wgrad_cfg = 2;
inline int pick_cfg(int M, int N, int K, bool a_kmajor, bool b_kmajor) {
  if (g_force_cfg >= 0) return (N < 256 && g_force_cfg > 0) ? 0 : g_force_cfg;
  if (N < 256) return 0;
  if (M < 2 * BM) return 1;
  // measured on the encoder shapes (profiles/r02_gemm_check_epi*.txt): with the remote mbarrier arrives relaxed and the
  // epilogue's side inputs pipelined, the CTA-pair kernel (4 epilogue warps) wins on every forward and data-gradient shape
  // (B K-major or MN-major); round 1 preferred the multicast kernel for MN-major B and the eight-epilogue-warp pair variant
  // for out-proj.  Weight gradients (A and B MN-major, split-K): see g_wgrad_cfg.
  if (!a_kmajor) return g_wgrad_cfg;
  return 4;
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
  int rc = tc_get_encode();
  if (rc != VQ_OK) return rc;

  TcParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(out_hi); p.out_lo = reinterpret_cast<__nv_bfloat16*>(out_lo); p.ldo = ldo;
  p.ep.bias = bias; p.ep.residual = residual; p.ep.ldr = ldr; p.ep.flags = flags; p.ep.alpha = alpha; p.ep.scale_ncols = scale_ncols;
  p.dbg = g_dbg;
  p.exp = g_exp;
  int cfg = pick_cfg(M, N, K, a_kmajor != 0, b_kmajor != 0);
  const bool epw8 = cfg == 5;                        // configuration 5 = configuration 4 with eight epilogue warps per CTA
  if (epw8) cfg = 4;
  const int BN = cfg >= 1 ? 256 : 128, BK = (cfg >= 1 && cfg <= 3) ? 32 : 64;
  const int CLh = cfg == 3 ? 4 : (cfg == 2 || cfg == 4 ? 2 : 1);   // B K-major boxes cover BN / CL rows (pair: own half)
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
    if ((rc = tc_make_map(&maps[0], A_hi, M, K, lda, BK, BM)) != VQ_OK) return rc;
    if ((rc = tc_make_map(&maps[1], alo, M, K, lda, BK, BM)) != VQ_OK) return rc;
  } else {          // [K rows][M cols]: box = 64 cols x BK rows
    if ((rc = tc_make_map(&maps[0], A_hi, K, M, lda, 64, BK)) != VQ_OK) return rc;
    if ((rc = tc_make_map(&maps[1], alo, K, M, lda, 64, BK)) != VQ_OK) return rc;
  }
  if (b_kmajor) {   // [N rows][K cols]
    if ((rc = tc_make_map(&maps[2], B_hi, N, K, ldb, BK, BN / CLh)) != VQ_OK) return rc;
    if ((rc = tc_make_map(&maps[3], blo, N, K, ldb, BK, BN / CLh)) != VQ_OK) return rc;
  } else {          // [K rows][N cols]
    if ((rc = tc_make_map(&maps[2], B_hi, K, N, ldb, 64, BK)) != VQ_OK) return rc;
    if ((rc = tc_make_map(&maps[3], blo, K, N, ldb, 64, BK)) != VQ_OK) return rc;
  }

  if (cfg == 4) return tc_launch_pair(a_kmajor, b_kmajor, passes, maps, p, st, epw8);
  if (a_kmajor && b_kmajor) return launch_tc<true, true>(cfg, passes, maps, p, st);
  if (a_kmajor && !b_kmajor) return launch_tc<true, false>(cfg, passes, maps, p, st);
  if (!a_kmajor && b_kmajor) return launch_tc<false, true>(cfg, passes, maps, p, st);
  return launch_tc<false, false>(cfg, passes, maps, p, st);
}

// Testing / tuning hook: force a tile configuration (0 = 128x128x64, 1 = 128x256x32, 2 / 3 = 128x256x32 with the B tile
// multicast over a cluster of 2 / 4 CTAs, 4 = CTA pairs with 256x256x32 tiles, 5 = 4 with eight epilogue warps per CTA
// (experimental), -1 = automatic).
VQ_API int vqcpc_gemm_set_tile_config(int cfg) {
  VQ_CHECK_ARG(cfg >= -1 && cfg <= 5, "gemm_set_tile_config: cfg must be -1 .. 5");
  g_force_cfg = cfg;
  return VQ_OK;
}

// Profiling hook: device buffer of 12 uint64 the CTA-pair kernel adds its phase timers to (NULL = off): [0] cycles the MMA
// threads spent in their tile loops, [1] of them waiting for a free accumulator stage (epilogue-bound), [2] waiting for
// operands (TMA-bound); [4 + q] cycles the epilogue warps of lane quadrant q waited for an accumulator, [8 + q] spent draining.
VQ_API int vqcpc_gemm_set_debug(void* buf) {
  g_dbg = reinterpret_cast<unsigned long long*>(buf);
  return VQ_OK;
}

// Timing experiments inside the epilogue (0 = off; anything else produces INVALID results - profiling tools only).
VQ_API int vqcpc_gemm_set_experiment(int e) {
  g_exp = e & 0xff;
  vqtc::g_no_static_epilogue = (e >> 8) & 1;      // bit 8: keep results valid but use the run-time-flag epilogue everywhere
  return VQ_OK;
}

// Which epilogue variant this library was built with (VQ_EPI_KIND): 0 = quad, 1 = staged, 2 = thread-per-row.
VQ_API int vqcpc_gemm_epilogue_kind(void) { return VQ_EPI_KIND; }

VQ_API int vqcpc_gemm_bf16x3(int a_kmajor, int b_kmajor, int M, int N, int K, const void* A_hi, const void* A_lo, int lda,
                             const void* B_hi, const void* B_lo, int ldb, float* C, int ldc, const float* bias,
                             const float* residual, int ldr, int flags, float alpha, int scale_ncols, int passes, int ksplit,
                             void* stream) {
  VQ_CHECK_ARG(C != nullptr, "gemm_bf16x3: null output");
  return vqcpc_gemm_bf16x3_ex(a_kmajor, b_kmajor, M, N, K, A_hi, A_lo, lda, B_hi, B_lo, ldb, C, ldc, nullptr, nullptr, 0, bias,
                              residual, ldr, flags, alpha, scale_ncols, passes, ksplit, stream);
}
