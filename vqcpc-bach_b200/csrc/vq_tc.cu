// Nearest-code search on the tensor cores with exact verification ("approximate, then verify") - the forward of the
// fused vector-quantisation bottleneck (ProductVectorQuantizer.forward, quantizer/vector_quantizer.py:104-157, with the
// dropout-free MlpUpscaler folded into the table T = upscaler(E), upscalers/mlp_upscaler.py:21-35).
//
// Why: the reference's argmin is over fl32 distances in direct form; evaluating that for every (vector, code) pair costs
// ~11 fp32 instructions per pair (vq_fused_fwd_kernel in quantizer.cu: 87 us ALU floor at P = 557 056, Kc = 512, against
// a 23 us HBM roofline for the whole quantizer + InfoNCE path).  Here the tensor pipe ranks the codes and the CUDA cores
// only (a) reduce the scores and (b) re-evaluate the reference arithmetic on the few codes that can still win:
//
//   score S[v, k] = z_v . e_k - |e_k|^2 / 2   (argmax S == argmin distance), as ONE tcgen05.mma (kind::f16, K = 16 per
//     D <= 4) per 128 vectors x 256 codes: operands are bf16 hi/lo splits laid out along K as
//        A row (vector):  [ z1 | z1 | z2 | 1  1  1  | 0 ]        z = z1 + z2 + O(2^-18 |z|)
//        B row (code)  :  [ e1 | e2 | e1 | n1 n2 n3 | 0 ]        n = -|e|^2/2 = n1 + n2 + n3
//     products of bf16 are exact in the fp32 accumulator, so |S^ - S| <= ~2^-16 (sum_d |z_d| max_k|e_kd| + max_k |n_k|).
//   epilogue (thread == vector, TMEM lane == vector): tcgen05.ld the 512 scores of a tile, 3-input max (FMNMX3) down to
//     64 group maxima (8 codes each) and the row maximum; every group whose maximum is within `margin` of the row
//     maximum is re-evaluated EXACTLY (dist_fixed: the reference's fp32 direct form, torch summation order, a tournament
//     that keeps the lower index on ties).  margin = 2 * (score error bound) + (rounding bound of the reference
//     distance), so the reference argmin is always among the candidates -> indices are bit-identical to the direct
//     kernel (tests/test_gpu_vq_tc.py), at ~1.5 instead of 11 instructions per (vector, code) pair; on average 1.03
//     groups per vector are re-evaluated.  Rows whose bound is not finite (inf / NaN / overflow) scan every code.
//
// CTA = 384 threads = 3 warpgroups, persistent over 128-vector tiles, one CTA per SM (all 512 TMEM columns = two
// 256-column accumulator stages = one tile of 512 codes; larger codebooks cycle through the stages in chunk pairs).
//   warpgroup 0 (setmaxnreg 56): warp 0 streams the z tiles into a 4-slot shared-memory ring with cp.async.bulk (one
//     1-D TMA copy per tile, mbarrier completion); all four warps build the SWIZZLE_32B K-major A tile from it (generic
//     stores + fence.proxy.async); one lane issues the MMAs and commits to mbarriers.
//   warpgroups 1, 2 (setmaxnreg 224): the two epilogue sets alternate tiles, so the score reduction of tile i+1 (TMEM
//     reads) overlaps the verification and output of tile i (shared memory + HBM writes: idx, loss, histogram,
//     y = T[idx] from the shared-memory copy of the table as coalesced 128-byte rows).
// Shared memory: operands, the exact fp32 codebook (padded against bank conflicts), the (Kc, zd) table (one bulk copy,
// overlapped with the set-up), the z ring, the usage histogram.
#include "gemm_tc_common.cuh"
#include "vq_dist.cuh"
#include "vq_tc.h"

using namespace vqtc;

namespace {

// Two thread layouts (template parameter SETS; tile i of a CTA belongs to epilogue set i % SETS):
//   SETS = 2: 384 threads = producer warpgroup (4 warps, setmaxnreg 56) + 2 epilogue warpgroups (setmaxnreg 224)
//   SETS = 3: 416 threads = ONE producer warp + 3 epilogue sets of 4 warps, no setmaxnreg (ptxas budget 65536 / 416 -> 152
//             registers for everybody; a 512-thread variant with a full producer warpgroup is capped at 128 and spills)
constexpr int VT_MAXSETS = 3;
__host__ __device__ constexpr int vt_threads(int sets) { return sets == 2 ? 384 : 32 + 128 * sets; }
constexpr int NABUF = 3;                  // A-operand buffers (see the reuse argument in the producer loop)
constexpr int CH = 256;               // codes per MMA / accumulator stage (two stages = all 512 TMEM columns)
// codes per group maximum: 8 with two epilogue sets (64 group maxima in registers), 16 with three (32: the 128-register
// budget of that layout; each candidate group then costs 16 exact distances instead of 8)
constexpr int ZS = 4;                 // depth of the z-tile ring in shared memory
constexpr int VT_PROF = 8;            // clock64 stamps per CTA in the stats buffer (after the 2 counters)

struct VtParams {
  const float* x;
  const float* E;
  const float* T;
  int64_t n;
  int Kc, nch, zd;
  float beta;
  int squared;
  int64_t* idx;
  float* loss;
  float* zsg;
  float* y;
  int64_t y_row0, y_rows;   // y holds T[idx] for the rows [y_row0, y_row0 + y_rows) only, stored from row 0 of y
  int* counts;
  float* dbg;            // optional (n, dbg_ld) dump of the tensor-core scores (tests only)
  int dbg_ld;
  int t_smem;            // 1: the (Kc, zd) table is staged in shared memory by one bulk copy
  int x_bulk;            // 1: z tiles can be fetched with cp.async.bulk (16-byte aligned base)
  int* stats;            // optional [2 + VT_PROF * gridDim] (callers pass 4098 ints): exact group re-evaluations,
                         // full-scan fallbacks, clock stamps
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
template <int G>
__device__ __forceinline__ float maxg(const uint32_t* r);
__device__ __forceinline__ float max8(const uint32_t* r);
template <>
__device__ __forceinline__ float maxg<8>(const uint32_t* r) { return max8(r); }
template <>
__device__ __forceinline__ float maxg<16>(const uint32_t* r) {
  float m = fmaxf(max8(r), max8(r + 8));
  return m;
}
__device__ __forceinline__ float max8(const uint32_t* r) {
  float m = fmax3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
  m = fmax3(m, __uint_as_float(r[3]), __uint_as_float(r[4]));
  m = fmax3(m, __uint_as_float(r[5]), __uint_as_float(r[6]));
  return fmaxf(m, __uint_as_float(r[7]));
}
// mask |= (a >= b) ? bit : 0   as two instructions (FSET + LOP3)
__device__ __forceinline__ uint32_t mask_ge(uint32_t mask, float a, float b, uint32_t bit) {
  uint32_t t;
  asm("set.ge.u32.f32 %0, %1, %2;" : "=r"(t) : "f"(a), "f"(b));
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(mask) : "r"(t), "r"(bit), "r"(mask));
  return mask;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
}

// byte offset of 16-byte chunk c (0/1) of K-row r inside one SWIZZLE_32B [rows][16 bf16] operand slab (bit 4 ^= bit 7)
__device__ __forceinline__ uint32_t krow_off(int r, int c) { return (uint32_t)(r * 32 + ((c ^ ((r >> 2) & 1)) << 4)); }

// K layout of one operand row: KTOT = 3*DS + 3 values, zero padded to KB*16 (see the header comment).
template <int DS, bool IS_A>
__device__ __forceinline__ void write_krow(uint8_t* base, uint32_t slab_bytes, int r, const float* v, float nrm) {
  constexpr int KTOT = 3 * DS + 3, KB = (KTOT + 15) / 16;
  float col[KB * 16];
#pragma unroll
  for (int j = 0; j < KB * 16; ++j) col[j] = 0.f;
#pragma unroll
  for (int d = 0; d < DS; ++d) {
    const float h = __bfloat162float(__float2bfloat16_rn(v[d]));
    const float l = __bfloat162float(__float2bfloat16_rn(v[d] - h));
    col[d] = h;
    col[DS + d] = IS_A ? h : l;          // A: z1 (meets e2)     B: e2
    col[2 * DS + d] = IS_A ? l : h;      // A: z2 (meets e1)     B: e1
  }
  if (IS_A) {
    col[3 * DS] = col[3 * DS + 1] = col[3 * DS + 2] = 1.f;
  } else {
    const float n1 = __bfloat162float(__float2bfloat16_rn(nrm));
    const float n2 = __bfloat162float(__float2bfloat16_rn(nrm - n1));
    col[3 * DS] = n1;
    col[3 * DS + 1] = n2;
    col[3 * DS + 2] = (nrm - n1) - n2;
  }
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
    uint8_t* slab = base + (size_t)kb * slab_bytes;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float* p = col + kb * 16 + c * 8;
      uint4 w = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
      *reinterpret_cast<uint4*>(slab + krow_off(r, c)) = w;
    }
  }
}

// fp32 codebook in shared memory: 16 bytes of padding after every 8 codes, so that lanes verifying the j-th code of
// DIFFERENT groups (addresses 128 B apart without the padding: the same four banks) spread over all bank groups
__device__ __forceinline__ int cb_off(int k, int cs) { return k * cs + ((k >> 3) << 2); }

// mbarrier indices (sized for the largest layout)
enum { BAR_FULL = 0 /*[set][stage]*/, BAR_EMPTY = 2 * VT_MAXSETS /*[set][stage]*/, BAR_ZFULL = 4 * VT_MAXSETS /*[ZS]*/,
       BAR_ZEMPTY = 4 * VT_MAXSETS + ZS /*[ZS]*/, BAR_T = 4 * VT_MAXSETS + 2 * ZS, BAR_COUNT = 4 * VT_MAXSETS + 2 * ZS + 1 };

// DBG: instantiation with the score dump and the clock / counter instrumentation (tests, tools/vq_bench.py --prof)
template <int DS, bool DBG, int SETS>
__global__ void __launch_bounds__(vt_threads(SETS), 1)
vq_tc_fwd_kernel(VtParams p) {
  constexpr int VT_SETS = SETS, VT_THREADS = vt_threads(SETS);
  constexpr int GROUP = SETS == 2 ? 8 : 16, NG = CH / GROUP;
  constexpr uint32_t GMASK = NG == 32 ? 0xFFFFFFFFu : ((1u << NG) - 1u);
  constexpr int NPW = SETS == 2 ? 4 : 1;                       // producer warps
  constexpr int KTOT = 3 * DS + 3, KB = (KTOT + 15) / 16;
  constexpr int CS = (DS == 3) ? 4 : DS;                       // padded fp32 code stride in smem
  constexpr uint32_t A_SLAB = 128 * 32, A_BUF = KB * A_SLAB;
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array itself keeps every access an LDS / STS / ATOMS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int kpad = p.nch * CH;
  const uint32_t B_SLAB = (uint32_t)kpad * 32;
  uint8_t* a_smem = smem;                                       // [NABUF][KB][128 rows][32 B]
  uint8_t* b_smem = a_smem + NABUF * A_BUF;                         // [KB][kpad rows][32 B]
  float* z_sm = reinterpret_cast<float*>(b_smem + (size_t)KB * B_SLAB);     // [ZS][128][DS]
  float* cb = z_sm + ZS * 128 * DS;                                         // [kpad][CS]
  float* t_sm = cb + (size_t)kpad * CS + (kpad >> 1);                       // [Kc][zd] (if t_smem)
  int* hist = reinterpret_cast<int*>(t_sm + (p.t_smem ? (((size_t)p.Kc * p.zd + 3) & ~(size_t)3) : 0));   // [Kc]
  float* red = reinterpret_cast<float*>(hist + ((p.Kc + 3) & ~3));          // [8]: max |e_d| per dim, then max |n|
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 8);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  const uint32_t bar0 = smem_u32(bars);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (p.n + 127) / 128;
  const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);     // tiles blockIdx.x + i * gridDim.x
  int* prof = (DBG && p.stats) ? p.stats + 2 + VT_PROF * blockIdx.x : nullptr;
  const long long t_start = clock64();

  // ------------------------------------------------------------------ set-up (all threads)
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * VT_SETS; ++i) {
      mbar_init(BAR(BAR_FULL + i), 1);
      mbar_init(BAR(BAR_EMPTY + i), 4);
    }
    for (int i = 0; i < ZS; ++i) {
      mbar_init(BAR(BAR_ZFULL + i), 1);
      mbar_init(BAR(BAR_ZEMPTY + i), NPW + 4);    // the producer warps + the four warps of the set that owns the tile
    }
    mbar_init(BAR(BAR_T), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.t_smem) {
      const uint32_t bytes = (uint32_t)((size_t)p.Kc * p.zd * 4);
      mbar_expect_tx(BAR(BAR_T), bytes);
      bulk_g2s(smem_u32(t_sm), p.T, bytes, BAR(BAR_T));     // overlaps the whole set-up and the first tile
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.Kc; i += VT_THREADS) hist[i] = 0;
  if (threadIdx.x < 8) red[threadIdx.x] = 0.f;
  __syncthreads();
  {
    float mx[DS + 1];
#pragma unroll
    for (int d = 0; d <= DS; ++d) mx[d] = 0.f;
    for (int k = threadIdx.x; k < kpad; k += VT_THREADS) {
      float e[DS];
      float nrm;
      if (k < p.Kc) {
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < DS; ++d) {
          e[d] = __ldg(p.E + (size_t)k * DS + d);
          s = fmaf(e[d], e[d], s);
          mx[d] = fmaxf(mx[d], fabsf(e[d]));
        }
        nrm = -0.5f * s;
        mx[DS] = fmaxf(mx[DS], fabsf(nrm));
      } else {
#pragma unroll
        for (int d = 0; d < DS; ++d) e[d] = 0.f;
        nrm = -3.0e38f;                                            // padding codes never rank
      }
      write_krow<DS, false>(b_smem, B_SLAB, k, e, nrm);
#pragma unroll
      for (int d = 0; d < CS; ++d) cb[cb_off(k, CS) + d] = d < DS ? (k < p.Kc ? e[d] : INFINITY) : 0.f;
    }
#pragma unroll
    for (int d = 0; d <= DS; ++d) {
      float m = warp_max(mx[d]);                                   // NaN codes: fmaxf drops them, inf survives
      if (lane == 0) atomicMax(reinterpret_cast<int*>(&red[d]), __float_as_int(m));   // non-negative floats order as ints
    }
  }
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (prof && threadIdx.x == 0) prof[0] = (int)(clock64() - t_start);

  if (warp < NPW) {
    // ============================== producer warps: z-tile fetch, operand producer, MMA issuer ==============================
    if (SETS == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const uint32_t idesc = make_idesc<CH>(true, true);
    constexpr uint32_t hi_c = (256u >> 4) | (1u << 14) | (6u << 29);      // SBO 256 B, descriptor version 1, SWIZZLE_32B
    constexpr uint32_t lo_c = 1u << 16;                                   // LBO (unused for swizzled K-major)
    const uint32_t a_base = smem_u32(a_smem) & 0x3FFFFu, b_base = smem_u32(b_smem) & 0x3FFFFu;
    // warp 0 fetches z tile `i` of this CTA into ring slot i % ZS (after all eight consumer warps released the slot)
    auto fetch = [&](int i) {
      if (i >= my_tiles) return;
      const int slot = i % ZS;
      if (i >= ZS) mbar_wait(BAR(BAR_ZEMPTY + slot), ((i / ZS) & 1) ^ 1);
      const int64_t row0 = (blockIdx.x + (int64_t)i * gridDim.x) * 128;
      const int rows = (int)min((int64_t)128, p.n - row0);
      const uint32_t bytes = (uint32_t)rows * DS * 4;
      float* dst = z_sm + slot * 128 * DS;
      if (p.x_bulk && (bytes & 15u) == 0) {
        if (lane == 0) {
          mbar_expect_tx(BAR(BAR_ZFULL + slot), bytes);
          bulk_g2s(smem_u32(dst), p.x + row0 * DS, bytes, BAR(BAR_ZFULL + slot));
        }
      } else {                                              // ragged tail / unaligned base: plain loads
        for (int j = lane; j < rows * DS; j += 32) dst[j] = p.x[row0 * DS + j];
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(BAR_ZFULL + slot));
      }
    };
    if (warp == 0)
      for (int i = 0; i < ZS - 1; ++i) fetch(i);
    uint32_t eph = 0;                              // bit (set*2 + stage): parity of BAR_EMPTY[set][stage]
    int last0 = -1, last1 = -1;                    // which epilogue set read accumulator stage 0 / 1 last
    for (int it = 0; it < my_tiles; ++it) {
      if (warp == 0) fetch(it + ZS - 1);
      const int slot = it % ZS, set = it % VT_SETS;
      mbar_wait(BAR(BAR_ZFULL + slot), (it / ZS) & 1);
      // A(it) goes to buffer it % 3, last read by the MMAs of tile it-3: warp 0 saw them retire (accumulator waits) while
      // issuing tile it-2, i.e. before it arrived at the producer barrier of tile it-1, which this thread has passed
      uint8_t* abuf = a_smem + (it % NABUF) * A_BUF;
#pragma unroll 1
      for (int rr = 0; rr < 4 / NPW; ++rr) {       // the operand rows this thread builds (one with four producer warps, four with one)
        const int rl = (warp + rr * NPW) * 32 + lane;
        const int64_t row = (blockIdx.x + (int64_t)it * gridDim.x) * 128 + rl;
        float zr[DS];
#pragma unroll
        for (int d = 0; d < DS; ++d) zr[d] = (row < p.n) ? z_sm[(slot * 128 + rl) * DS + d] : 0.f;
        write_krow<DS, true>(abuf, A_SLAB, rl, zr, 0.f);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(BAR_ZEMPTY + slot));
      if (NPW == 4) asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 0 && lane == 0) {
        const uint32_t a_addr = a_base + (uint32_t)(it % NABUF) * A_BUF;
        for (int c = 0; c < p.nch; ++c) {
          const int st = c & 1;
          const int last = st ? last1 : last0;
          if (last >= 0) {
            const int b = last * 2 + st;
            mbar_wait(BAR(BAR_EMPTY + b), (eph >> b) & 1u);
            eph ^= 1u << b;
          }
          if (st) last1 = set; else last0 = set;
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(st * CH);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t da = pack_desc(lo_c | ((a_addr + kb * A_SLAB) >> 4), hi_c);
            const uint64_t db = pack_desc(lo_c | ((b_base + kb * B_SLAB + (uint32_t)c * CH * 32) >> 4), hi_c);
            umma_bf16(tmem_d, da, db, idesc, kb > 0 ? 1u : 0u);
          }
          umma_commit(BAR(BAR_FULL + set * 2 + st));
        }
        if (prof && it == 0) prof[1] = (int)(clock64() - t_start);
      }
    }
  } else {
    // ============================== score reduction + exact verification ==============================
    // set s (warpgroup 1 + s) owns the tiles s, s + VT_SETS, ... of this CTA; thread == vector.
    // Register budget: 128 x 56 (producers) + 256 x 224 (two sets) <= 64 K
    if (SETS == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int set = (warp - NPW) >> 2;
    const int q = warp & 3;                        // TMEM lane quadrant (hardware: warp id % 4); each set covers all four
    const bool stamp = DBG && prof && warp == NPW && lane == 0;
    const float nmax = red[DS];
    float emax[DS];
#pragma unroll
    for (int d = 0; d < DS; ++d) emax[d] = red[d];
    uint32_t fph = 0;                              // bit stage: parity of BAR_FULL[set][stage]
    long long tk = 0;                              // profiling (warp 4 only): clocks spent waiting for z, waiting for
    int acc_z = 0, acc_w = 0, acc_p1 = 0, acc_v = 0, acc_o = 0, acc_m = 0;   // accumulators, in phase 1, verifying, writing outputs
#define TICK(acc) do { if (stamp) { const long long t2_ = clock64(); acc += (int)(t2_ - tk); tk = t2_; } } while (0)
    int n_exact = 0, n_full = 0;
    bool t_ready = !p.t_smem;
    const int rl = q * 32 + lane;
    for (int it = set; it < my_tiles; it += VT_SETS) {
      const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
      const int64_t row = tile * 128 + rl;
      const bool ok = row < p.n;
      const int slot = it % ZS;
      if (stamp) tk = clock64();
      mbar_wait(BAR(BAR_ZFULL + slot), (it / ZS) & 1);
      TICK(acc_z);
      float xr[DS];
#pragma unroll
      for (int d = 0; d < DS; ++d) xr[d] = ok ? z_sm[(slot * 128 + rl) * DS + d] : 0.f;
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(BAR_ZEMPTY + slot));
      // margin = 2 * (score error) + (rounding of the reference distance), see header
      //   cross terms: |z.e - (z1e1 + z1e2 + z2e1)| <= 3 * 2^-18 sum_d |z_d||e_d|     -> budget 2^-14 P,  P = sum_d |z_d| max_k |e_kd|
      //   norm term  : fp32 evaluation of n (D+1 roundings) + exact 3-way bf16 split    -> budget 2^-20 N,  N = max_k |n_k|
      //   accumulation of <= 16 KB exact products in fp32 inside the tensor core        -> budget KB * 2^-19 (P + N)
      //   reference  : fl32 direct form, (D + 2) roundings of at most R = sum_d (|z_d| + max|e_d|)^2, halved in score units
      float pz = 0.f, rz = 0.f;
#pragma unroll
      for (int d = 0; d < DS; ++d) {
        const float az = fabsf(xr[d]);
        pz = fmaf(az, emax[d], pz);
        rz = fmaf(az + emax[d], az + emax[d], rz);
      }
      const float eps = 6.103515625e-05f * pz + 9.5367431640625e-07f * nmax + (float)KB * 1.9073486328125e-06f * (pz + nmax);
      const float margin = 2.f * eps + (float)(DS + 2) * 5.9604645e-08f * rz;
      const bool scan_all = !(margin < INFINITY);
      float best = INFINITY, runmax = -INFINITY;
      int bi = -1;
      // chunks are consumed in pairs (one per accumulator stage); the threshold uses the maximum seen so far, which
      // only admits more candidates than the final maximum would
      for (int c0 = 0; c0 < p.nch; c0 += 2) {
        float gm[2 * NG];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = c0 + h;
          if (c < p.nch) {
            TICK(acc_v);
            mbar_wait(BAR(BAR_FULL + set * 2 + h), (fph >> h) & 1u);
            fph ^= 1u << h;
            TICK(acc_w);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * CH);
            // software pipeline over the eight 32-column chunks of the stage: the TMEM load of chunk j + 1 is in flight
            // while chunk j is reduced (tcgen05.wait::ld waits for everything outstanding, so the wait sits after the
            // reduction of the previous chunk)
            constexpr int NB = SETS == 2 ? 2 : 1;       // (the three-set layout has no registers for a second buffer)
            uint32_t rb[NB][32];
            tmem_ld32(taddr, rb[0]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CH / 32; ++j) {
              uint32_t* r0 = rb[j % NB];
              if (NB == 2 && j + 1 < CH / 32) tmem_ld32(taddr + (j + 1) * 32, rb[(j + 1) % NB]);
              if (DBG && p.dbg && ok) {
                float* drow = p.dbg + (size_t)row * p.dbg_ld + c * CH + j * 32;
#pragma unroll
                for (int t = 0; t < 32; ++t) drow[t] = __uint_as_float(r0[t]);
              }
#pragma unroll
              for (int t = 0; t < 32 / GROUP; ++t) gm[h * NG + j * (32 / GROUP) + t] = maxg<GROUP>(r0 + t * GROUP);
              if (NB == 1 && j + 1 < CH / 32) tmem_ld32(taddr + (j + 1) * 32, rb[0]);
              if (j + 1 < CH / 32) tmem_ld_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(BAR_EMPTY + set * 2 + h));
            TICK(acc_p1);
          } else {
#pragma unroll
            for (int t = 0; t < NG; ++t) gm[h * NG + t] = -INFINITY;
          }
        }
        // maximum of the group maxima: four independent 3-input chains instead of one long dependent chain
        float cm[4] = {gm[0], gm[1], gm[2], gm[3]};
#pragma unroll
        for (int t = 4; t + 7 < 2 * NG; t += 8) {
          cm[0] = fmax3(cm[0], gm[t], gm[t + 1]);
          cm[1] = fmax3(cm[1], gm[t + 2], gm[t + 3]);
          cm[2] = fmax3(cm[2], gm[t + 4], gm[t + 5]);
          cm[3] = fmax3(cm[3], gm[t + 6], gm[t + 7]);
        }
        float cmax = fmax3(cm[0], cm[1], fmaxf(cm[2], cm[3]));
#pragma unroll
        for (int t = 4 + (2 * NG - 4) / 8 * 8; t < 2 * NG; ++t) cmax = fmaxf(cmax, gm[t]);
        runmax = fmaxf(runmax, cmax);
        const float thr = runmax - margin;
        // candidate masks: four independent accumulators per half instead of one 32-long LOP3 chain
        uint32_t ml[4] = {0, 0, 0, 0}, mh[4] = {0, 0, 0, 0};
#pragma unroll
        for (int t = 0; t < NG; ++t) {
          ml[t & 3] = mask_ge(ml[t & 3], gm[t], thr, 1u << t);
          mh[t & 3] = mask_ge(mh[t & 3], gm[NG + t], thr, 1u << t);
        }
        uint32_t mlo = (ml[0] | ml[1]) | (ml[2] | ml[3]), mhi = (mh[0] | mh[1]) | (mh[2] | mh[3]);
        if (scan_all || !(cmax == cmax)) { mlo = mhi = GMASK; ++n_full; }            // NaN / inf / overflow: scan all
        if (!ok) mlo = mhi = 0u;
        TICK(acc_m);
        // ascending code order + strict '<' = lowest index on ties
        while (mlo | mhi) {
          int t;
          if (mlo) { t = __ffs(mlo) - 1; mlo &= mlo - 1; }
          else { t = NG + __ffs(mhi) - 1; mhi &= mhi - 1; }
          ++n_exact;
          const int k0 = c0 * CH + t * GROUP;
          // all eight codes of the group: loads first, eight independent distance chains, then a tournament that keeps
          // the lower index on ties (padding codes hold +inf coordinates, so their distance is +inf / NaN and never wins)
#pragma unroll 1
          for (int sub = 0; sub < GROUP; sub += 8) {       // eight codes at a time (ascending: strict '<' keeps the lowest index)
            float dd[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) dd[j] = dist_fixed<DS>(xr, cb + cb_off(k0 + sub + j, CS));
            int ki[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ki[j] = k0 + sub + j;
#pragma unroll
            for (int w = 1; w < 8; w <<= 1)
#pragma unroll
              for (int j = 0; j + w < 8; j += 2 * w)
                if (dd[j + w] < dd[j]) { dd[j] = dd[j + w]; ki[j] = ki[j + w]; }
            if (dd[0] < best) { best = dd[0]; bi = ki[0]; }
          }
        }
      }
      TICK(acc_v);
      if (bi < 0) bi = 0;                          // every distance inf / NaN: torch.argmin returns index 0
      if (ok) {
        float zs[DS];
        p.loss[row] = vq_loss_fixed<DS>(xr, cb + cb_off(bi, CS), p.beta, p.squared, p.zsg ? zs : nullptr);
        if (p.zsg) {
#pragma unroll
          for (int d = 0; d < DS; ++d) p.zsg[row * DS + d] = zs[d];
        }
        p.idx[row] = bi;
        if (p.counts) atomicAdd(&hist[bi], 1);
      }
      // y = T[idx] only for the rows of the range (the index path needs the upscaled vector of the LEFT blocks alone: the
      // context GRU reads them; right / negative blocks are scored from their indices).  Warp-uniform early-out.
      if (p.y && tile * 128 + q * 32 < p.y_row0 + p.y_rows && tile * 128 + q * 32 + 32 > p.y_row0) {
        const int64_t row0 = tile * 128 + q * 32;
        const int64_t yend = min(p.n, p.y_row0 + p.y_rows);
        float* ybase = p.y - p.y_row0 * (int64_t)p.zd;
        const int zd = p.zd;
        if (!t_ready) {
          mbar_wait(BAR(BAR_T), 0);
          t_ready = true;
        }
        if (zd == 32 && p.t_smem) {
          // 8 lanes cover one 128-byte row; 4 rows per pass, 8 passes, all loads issued before the stores
          float4 v[8];
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const int b = __shfl_sync(0xffffffffu, bi, ps * 4 + (lane >> 3));
            v[ps] = *(reinterpret_cast<const float4*>(t_sm + b * 32) + (lane & 7));
          }
          float4* yp = reinterpret_cast<float4*>(ybase + row0 * 32) + lane;
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const int64_t rr = row0 + ps * 4 + (lane >> 3);
            if (rr >= p.y_row0 && rr < yend) yp[ps * 32] = v[ps];
          }
        } else {
          for (int r = 0; r < 32; ++r) {
            const int b = __shfl_sync(0xffffffffu, bi, r);
            if (row0 + r >= p.y_row0 && row0 + r < yend)
              for (int j = lane; j < zd; j += 32)
                ybase[(row0 + r) * zd + j] = p.t_smem ? t_sm[(size_t)b * zd + j] : __ldg(p.T + (size_t)b * zd + j);
          }
        }
      }
      TICK(acc_o);
    }
#undef TICK
    if (stamp) { prof[2] = acc_m; prof[3] = acc_z; prof[4] = acc_w; prof[5] = acc_p1; prof[6] = acc_v; prof[7] = acc_o; }
    if (DBG && p.stats) {
      if (n_exact) atomicAdd(p.stats, n_exact);
      if (n_full) atomicAdd(p.stats + 1, n_full);
    }
  }
#undef BAR

  tcgen05_fence_before();
  __syncthreads();
  if (p.counts)
    for (int i = threadIdx.x; i < p.Kc; i += VT_THREADS)
      if (hist[i]) atomicAdd(p.counts + i, hist[i]);
  if (warp == 0) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <int DS>
size_t vt_smem_bytes(int Kc, int kpad, int zd, bool t_smem) {
  constexpr int KTOT = 3 * DS + 3, KB = (KTOT + 15) / 16;
  constexpr int CS = (DS == 3) ? 4 : DS;
  size_t s = 1024;                                   // alignment slack
  s += NABUF * (size_t)KB * 128 * 32;                // A buffers
  s += (size_t)KB * kpad * 32;                       // B slabs
  s += (size_t)ZS * 128 * DS * 4;                    // z ring
  s += ((size_t)kpad * CS + (kpad >> 1)) * 4;        // exact codebook (padded, see cb_off)
  if (t_smem) s += (((size_t)Kc * zd + 3) & ~(size_t)3) * 4;
  s += ((size_t)(Kc + 3) & ~(size_t)3) * 4;          // histogram
  s += 32 + BAR_COUNT * 8 + 16;                      // reduction scratch, barriers, tmem slot
  return s;
}

template <int DS>
int vt_launch(VtParams& p, int sets, cudaStream_t st) {
  const int64_t ntiles = (p.n + 127) / 128;
  const bool t_ok = p.T && p.zd >= 1 && ((reinterpret_cast<uintptr_t>(p.T) & 15) == 0) && (((size_t)p.Kc * p.zd * 4) % 16 == 0);
  const int kpad = (p.Kc + CH - 1) / CH * CH;
  constexpr size_t FULL = 227 * 1024;
  p.nch = kpad / CH;
  p.t_smem = (t_ok && vt_smem_bytes<DS>(p.Kc, kpad, p.zd, true) <= FULL) ? 1 : 0;
  p.x_bulk = ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0) ? 1 : 0;
  size_t smem = vt_smem_bytes<DS>(p.Kc, kpad, p.zd, p.t_smem);
  VQ_CHECK_ARG(smem <= FULL, "vq_tc: operands for Kc=%d D=%d (%zu B) do not fit in shared memory", p.Kc, DS, smem);
  // one CTA per SM by construction (each allocates all 512 TMEM columns): make co-residency impossible
  smem = std::max<size_t>(smem, 120 * 1024);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)vq_sm_count()));
  if (p.dbg || p.stats) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(vq_tc_fwd_kernel<DS, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_tc_fwd_kernel<DS, true, 2><<<grid, vt_threads(2), smem, st>>>(p);
  } else if (sets == 3) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(vq_tc_fwd_kernel<DS, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_tc_fwd_kernel<DS, false, 3><<<grid, vt_threads(3), smem, st>>>(p);
  } else {
    VQ_CUDA_CALL(cudaFuncSetAttribute(vq_tc_fwd_kernel<DS, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_tc_fwd_kernel<DS, false, 2><<<grid, vt_threads(2), smem, st>>>(p);
  }
  VQ_CUDA_LAUNCH_CHECK("vq_tc_fwd_kernel");
  return VQ_OK;
}

}  // namespace

bool vq_tc_supported(int D, int Kc) { return (D == 3 || D == 4) && Kc >= 1 && Kc <= 2048; }

int vq_tc_fused_fwd(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta, int squared,
                    int64_t* idx, float* loss, float* z_sg, float* y, int64_t y_row0, int64_t y_rows, int* counts,
                    float* dbg_scores, int* stats, int variant, cudaStream_t st) {
  const int sets = variant == 3 ? 3 : 2;              // VQCPC_VQTC_LAYOUT=3: three epilogue sets behind a single producer warp
  VQ_CHECK_ARG(vq_tc_supported(D, Kc), "vq_tc: unsupported D=%d / Kc=%d", D, Kc);
  VtParams p;
  p.x = x; p.E = E; p.T = T; p.n = n; p.Kc = Kc; p.nch = 0; p.zd = zd; p.beta = beta; p.squared = squared;
  p.idx = idx; p.loss = loss; p.zsg = z_sg; p.y = y; p.y_row0 = y_row0; p.y_rows = y_rows; p.counts = counts; p.dbg = dbg_scores; p.dbg_ld = (Kc + 255) / 256 * 256;
  p.t_smem = 0; p.x_bulk = 0; p.stats = stats;
  switch (D) {
    case 3: return vt_launch<3>(p, sets, st);
    default: return vt_launch<4>(p, sets, st);
  }
}
