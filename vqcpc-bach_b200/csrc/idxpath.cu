// Index-gather form of the quantizer -> upscaler -> InfoNCE path (table mode: no active upscaler dropout, one codebook).
//
// Every quantised + upscaled vector of the step is one of only Kc rows of the table T = upscaler(E) (Kc, zd), so the
// InfoNCE scores of VQCPCB/vqcpc_helper.py:79-98 / :5-29 (f_k = u_k . y, u_k = W_k c) and everything behind them can be
// computed from the code INDICES instead of from materialised 128-byte rows:
//
//   forward   s_n = u_k . T[idx_n]                 (n = 0: the positive, 1..N: the negatives of (b, k))  -> lse, nll, accuracy
//   backward  coef_n = g (softmax(s)_n - [n = 0]);  du_k = sum_n coef_n T[idx_n];  dc = W_k^T du_k;  dW_k += du_k (x) c
//             dT[code] += sum_{(b,k,n): idx = code} coef_n u_k                      (gradient of the upscaler table)
//             dz_p = coef_p J[idx_p]^T u_k  (+ the commitment term)                 (straight-through, vector_quantizer.py:148)
//
// Real traffic per sample drops from 36.7 KB (z_up rows written by the quantizer and re-read by the
// scorer, dy rows written and re-read in backward) to ~3 KB (indices, z, u, scores); the tables (T 64 KB, J 192 KB, W 32 KB)
// live in shared memory / registers of persistent CTAs.  The tensor pipe is NOT used here on purpose: scoring 16 candidates
// per row from a shared-memory table is 512 FMAs per row, a dense U.T^T product on tcgen05 would compute all Kc scores per
// row (32x the work) and TMEM cannot be gathered per lane (tcgen05.ld addresses are warp-uniform).
//
// No floating-point shared-memory atomics anywhere (CAS loops on sm_100): scatter-adds are turned into
//   * owner-computes sweeps (a warp owns a residue class of codes and adds into its rows of a CTA-private table), or
//   * a counting sort by code inside the CTA (native integer ATOMS) followed by per-code sequential sums,
// and the cross-CTA reduction goes through per-CTA partial tables summed by a second small kernel (fixed order).
#include "gemm_tc_common.cuh"
#include <algorithm>
#include <cstdlib>

using namespace vqtc;

namespace {

constexpr int IZ = 32;          // zd == cd == 32: one lane per feature (every shipped configuration)
constexpr int TST = 36;         // padded row stride (floats) of the shared-memory table: rows start 4 banks apart
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return smem_u32(p); }

// W (z, c, k) -> W_s[k][row][col]: row = z, col = c (ZMAJOR: u = W_k c) or row = c, col = z (dc = W_k^T du); padded rows.
// One (z, c) pair per thread step: KR consecutive floats in, KR scattered stores out.
template <bool ZMAJOR>
__device__ __forceinline__ void stage_w(float* W_s, const float* __restrict__ W, int KR) {
  for (int zc = threadIdx.x; zc < IZ * IZ; zc += blockDim.x) {
    const int z = zc >> 5, c = zc & 31;
    const int row = ZMAJOR ? z : c, col = ZMAJOR ? c : z;
    for (int k = 0; k < KR; ++k) W_s[(k * IZ + row) * TST + col] = __ldg(W + (size_t)zc * KR + k);
  }
}

// global loads that stay where they are written (asm volatile): used for software prefetching, where ptxas would otherwise
// sink the loads towards their first use to shorten live ranges
__device__ __forceinline__ float ldg_now(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ long long ldg_now(const int64_t* p) {
  long long v;
  asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// stage the (Kc, 32) table with padded rows
__device__ __forceinline__ void stage_table(float* T_s, const float* __restrict__ T, int Kc) {
  for (int i = threadIdx.x; i < Kc * 8; i += blockDim.x) {
    const int r = i >> 3, ch = i & 7;
    *reinterpret_cast<float4*>(T_s + r * TST + ch * 4) = __ldg(reinterpret_cast<const float4*>(T) + i);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
struct NceFwdP {
  const float *c, *W, *T;
  const int64_t *idx_pos, *idx_neg;
  int B, N, KR, Kc, SPI;
  float *u, *pos, *neg, *lse, *nll;
  int* correct;
};

__device__ long long g_idx_dbg[16];

// CTA = KR * SPI warps, one per SM; warp (s, k) always handles step k, so W[:, :, k] stays in its registers (no shared-memory
// traffic for W) and walks the samples two at a time: lanes = features while the two u = W_k c are formed, then lane (j, n)
// scores candidate n of row j with one 32-term dot product, so the log-sum-exp shuffles serve both rows at once.
// The table is ONE bulk copy in its global layout; the row gathers are bank-conflict free by construction: lane l visits the
// eight 16-byte chunks of its row in the order (q + l) & 7, so the eight lanes of a quarter warp always touch eight different
// bank groups.  The summation order therefore depends on the lane; a negative that drew the positive's code takes the
// positive's score (one shuffle), which keeps such a candidate bit-equal and never "correct", like the reference.
// Per-row global reads (candidate code, one element of c per lane) are fetched one step ahead.
__global__ void __launch_bounds__(512, 1) nce_idx_fwd_kernel(NceFwdP p) {
  extern __shared__ __align__(16) float sm[];
  float* T_s = sm;                                   // [Kc][32]
  float* u_s = T_s + (size_t)p.Kc * IZ;              // [warps][2][32]: the c vectors of the two rows, then their u vectors
  float* W_s = u_s + (blockDim.x >> 5) * 2 * IZ;     // [KR][32 z][TST] staging of W (coalesced global reads) on its way to registers
  uint64_t* bar = reinterpret_cast<uint64_t*>(W_s + (size_t)p.KR * IZ * TST);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_a = smem_u32(bar);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)p.Kc * IZ * 4;
    mbar_expect_tx(bar_a, bytes);
    for (uint32_t o = 0; o < bytes; o += 16384) {
      const uint32_t nb = min(16384u, bytes - o);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(T_s) + o),
                   "l"(reinterpret_cast<const char*>(p.T) + o), "r"(nb), "r"(bar_a)
                   : "memory");
    }
  }
  const long long t_begin = clock64();
  const bool dbg = blockIdx.x == 0 && threadIdx.x == 0;
  long long t_u = 0, t_sc = 0, t_sm = 0, t_top = 0;
  const int KR = p.KR, N = p.N, SPI = p.SPI;
  const int k = warp % KR, s = warp / KR;
  stage_w<true>(W_s, p.W, KR);
  __syncthreads();                                   // (also: mbarrier initialised before anybody waits on it)
  float Wc[IZ];                                      // Wc[c] = W[z = lane][c][k]
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 w = *reinterpret_cast<const float4*>(W_s + (k * IZ + lane) * TST + 4 * q);
    Wc[4 * q] = w.x; Wc[4 * q + 1] = w.y; Wc[4 * q + 2] = w.z; Wc[4 * q + 3] = w.w;
  }
  float* us = u_s + warp * 2 * IZ;
  const int n = lane & 15, j = lane >> 4;            // candidate n (0 = positive) of row j of the step
  const bool act = n <= N;
  const int B = p.B, bstep = gridDim.x * SPI;
  // step i of this warp: samples b0 = base + 2 i bstep and b1 = b0 + bstep
  auto fetch = [&](int b0, int& code, float& c0, float& c1) {
    code = 0;
    c0 = c1 = 0.f;
    const int b1 = b0 + bstep;
    if (b0 < B) c0 = ldg_now(p.c + b0 * IZ + lane);
    if (b1 < B) c1 = ldg_now(p.c + b1 * IZ + lane);
    const int bj = j ? b1 : b0;
    if (act && bj < B) code = (int)ldg_now(n == 0 ? p.idx_pos + bj * KR + k : p.idx_neg + (bj * N + (n - 1)) * KR + k);
  };
  int b0 = blockIdx.x * SPI + s;
  int code_n;
  float c0_n, c1_n;
  fetch(b0, code_n, c0_n, c1_n);
  bool t_ready = false;
  const long long t_setup = clock64();
  long long t_wait = 0;
  int nsteps = 0;
  for (; b0 < B; b0 += 2 * bstep) {
    long long t0 = clock64();
    ++nsteps;
    const int code = min(max(code_n, 0), p.Kc - 1);
    us[lane] = c0_n;
    us[IZ + lane] = c1_n;
    fetch(b0 + 2 * bstep, code_n, c0_n, c1_n);       // next step's reads are in flight during this step
    __syncwarp();
    { const long long t1 = clock64(); t_top += t1 - t0; t0 = t1; }
    const int b1 = b0 + bstep;
    const bool v1 = b1 < B;
    // ---- u of both rows, lanes = features
    float uu[2];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const float4* c4 = reinterpret_cast<const float4*>(us + jj * IZ);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = c4[q];                      // broadcast
        a0 = fmaf(Wc[4 * q], t.x, a0);
        a1 = fmaf(Wc[4 * q + 1], t.y, a1);
        a2 = fmaf(Wc[4 * q + 2], t.z, a2);
        a3 = fmaf(Wc[4 * q + 3], t.w, a3);
      }
      uu[jj] = (a0 + a1) + (a2 + a3);
    }
    p.u[(b0 * KR + k) * IZ + lane] = uu[0];
    if (v1) p.u[(b1 * KR + k) * IZ + lane] = uu[1];
    __syncwarp();                                    // the u vectors replace the c vectors in the warp's buffer
    us[lane] = uu[0];
    us[IZ + lane] = uu[1];
    __syncwarp();
    { const long long t1 = clock64(); t_u += t1 - t0; t0 = t1; }
    if (!t_ready) {
      mbar_wait(bar_a, 0);
      t_ready = true;
      const long long t1 = clock64(); t_wait = t1 - t0; t0 = t1;
    }
    // ---- scores
    const float4* t4 = reinterpret_cast<const float4*>(T_s + code * IZ);
    const float4* u4 = reinterpret_cast<const float4*>(us + j * IZ);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 a = t4[(q + lane) & 7], v = u4[(q + lane) & 7];
      a0 = fmaf(a.x, v.x, a0);
      a1 = fmaf(a.y, v.y, a1);
      a2 = fmaf(a.z, v.z, a2);
      a3 = fmaf(a.w, v.w, a3);
    }
    float sc = (a0 + a1) + (a2 + a3);
    {
      const int code0 = __shfl_sync(FULL, code, lane & 16);          // the positive of this lane's row
      const float sc0 = __shfl_sync(FULL, sc, lane & 16);
      if (code == code0) sc = sc0;
    }
    if (!act) sc = -INFINITY;
    { const long long t1 = clock64(); t_sc += t1 - t0; t0 = t1; }
    float m = sc, nm = (n == 0) ? -INFINITY : sc;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
      nm = fmaxf(nm, __shfl_xor_sync(FULL, nm, o));
    }
    float e = act ? expf(sc - m) : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) e += __shfl_xor_sync(FULL, e, o);
    const float lse = m + logf(e);
    const int bj = j ? b1 : b0;
    if (bj < B) {
      const int rj = bj * KR + k;
      if (n == 0) {
        p.pos[rj] = sc;
        p.lse[rj] = lse;
        p.nll[rj] = lse - sc;
        p.correct[rj] = sc > nm ? 1 : 0;
      } else if (act) {
        p.neg[rj * N + n - 1] = sc;
      }
    }
    __syncwarp();
    { const long long t1 = clock64(); t_sm += t1 - t0; t0 = t1; }
  }
  if (!t_ready) mbar_wait(bar_a, 0);                 // never leave with a bulk copy in flight
  if (dbg) {
    const long long t_end = clock64();
    g_idx_dbg[0] = t_setup - t_begin; g_idx_dbg[1] = t_wait; g_idx_dbg[2] = t_end - t_setup; g_idx_dbg[3] = nsteps;
    g_idx_dbg[4] = t_top; g_idx_dbg[5] = t_u; g_idx_dbg[6] = t_sc; g_idx_dbg[7] = t_sm;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward of the scorer: coefficients, dc, per-CTA partial dW and dT
// ---------------------------------------------------------------------------------------------------------------------
struct NceBwdP {
  const float *c, *W, *T;
  const int64_t *idx_pos, *idx_neg;
  const float *u, *pos, *neg, *lse, *g_dev;
  float scale;
  int B, N, KR, Kc, SPI;
  float *dc, *coef_pos, *coef_neg, *part;            // part: [gridDim][Kc * 32 + 32 * 32 * KR]
  float* dcp;                                        // [B * KR][32] per-row W_k^T du (tensor-core kernel; summed over k later)
  size_t scratch_floats;                             // size of the T_s / scratch region
};

__global__ void __launch_bounds__(512, 1) nce_idx_bwd_kernel(NceBwdP p) {
  extern __shared__ __align__(16) float sm[];
  const int NW = blockDim.x >> 5, R = NW;            // rows (b, k) per iteration == warps
  float* T_s = sm;                                   // [Kc][TST]; reused as [NW][32][32] scratch at the end
  float* dT_s = T_s + p.scratch_floats;              // [Kc][32]  CTA-private gradient of the table
  float* u_s = dT_s + (size_t)p.Kc * IZ;             // [R][32]
  float* du_s = u_s + R * IZ;                        // [R][32]
  float* dcp_s = du_s + R * IZ;                      // [R][32]
  float* coef_s = dcp_s + R * IZ;                    // [R][16]
  uint32_t* code_s = reinterpret_cast<uint32_t*>(coef_s + R * 16);   // [R][16]: code | owner warp << 16 (0xFFFF....: none)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_table(T_s, p.T, p.Kc);
  for (int i = threadIdx.x; i < p.Kc * IZ; i += blockDim.x) dT_s[i] = 0.f;
  const int k = warp % p.KR, s = warp / p.KR;
  float Wz[IZ], acc[IZ];                             // Wz[z] = W[z][c = lane][k];  acc[z] = dW[z][c = lane][k]
#pragma unroll
  for (int z = 0; z < IZ; ++z) {
    Wz[z] = __ldg(p.W + ((size_t)z * IZ + lane) * p.KR + k);
    acc[z] = 0.f;
  }
  const float gs = p.scale * (p.g_dev ? __ldg(p.g_dev) : 1.f);
  __syncthreads();
  const int64_t nit = (p.B + (int64_t)gridDim.x * p.SPI - 1) / ((int64_t)gridDim.x * p.SPI);
  for (int64_t it = 0; it < nit; ++it) {
    const int64_t b = (it * gridDim.x + blockIdx.x) * p.SPI + s;
    const bool valid = b < p.B;
    const int64_t r = b * p.KR + k;
    // ---- phase a: this warp's row
    if (lane < 16) {
      const int n = lane;
      uint32_t cw = 0xFFFF0000u;
      float cf = 0.f;
      if (valid && n <= p.N) {
        const float sc = (n == 0) ? __ldg(p.pos + r) : __ldg(p.neg + r * p.N + n - 1);
        int code = (int)(n == 0 ? p.idx_pos[r] : p.idx_neg[(b * p.N + (n - 1)) * p.KR + k]);
        code = min(max(code, 0), p.Kc - 1);
        cf = gs * (expf(sc - __ldg(p.lse + r)) - (n == 0 ? 1.f : 0.f));
        if (n == 0) p.coef_pos[r] = cf; else p.coef_neg[(b * p.N + (n - 1)) * p.KR + k] = cf;
        cw = (uint32_t)code | ((uint32_t)(code % NW) << 16);
      }
      code_s[warp * 16 + n] = cw;
      coef_s[warp * 16 + n] = cf;
    }
    u_s[warp * IZ + lane] = valid ? __ldg(p.u + r * IZ + lane) : 0.f;
    const float cb = valid ? __ldg(p.c + b * IZ + lane) : 0.f;         // lane = c
    __syncwarp();
    float du = 0.f;                                                     // lane = z
#pragma unroll
    for (int n4 = 0; n4 < 4; ++n4) {
      const uint4 cd = *reinterpret_cast<const uint4*>(code_s + warp * 16 + n4 * 4);
      const float4 cf = *reinterpret_cast<const float4*>(coef_s + warp * 16 + n4 * 4);
      du = fmaf(cf.x, T_s[(cd.x & 0xFFFFu) * TST + lane], du);
      du = fmaf(cf.y, T_s[(cd.y & 0xFFFFu) * TST + lane], du);
      du = fmaf(cf.z, T_s[(cd.z & 0xFFFFu) * TST + lane], du);
      du = fmaf(cf.w, T_s[(cd.w & 0xFFFFu) * TST + lane], du);
    }
    du_s[warp * IZ + lane] = du;
    __syncwarp();
    float dck = 0.f;
#pragma unroll
    for (int z4 = 0; z4 < 8; ++z4) {
      const float4 d4 = *reinterpret_cast<const float4*>(du_s + warp * IZ + z4 * 4);
      dck = fmaf(Wz[4 * z4], d4.x, dck);
      dck = fmaf(Wz[4 * z4 + 1], d4.y, dck);
      dck = fmaf(Wz[4 * z4 + 2], d4.z, dck);
      dck = fmaf(Wz[4 * z4 + 3], d4.w, dck);
      acc[4 * z4] = fmaf(d4.x, cb, acc[4 * z4]);
      acc[4 * z4 + 1] = fmaf(d4.y, cb, acc[4 * z4 + 1]);
      acc[4 * z4 + 2] = fmaf(d4.z, cb, acc[4 * z4 + 2]);
      acc[4 * z4 + 3] = fmaf(d4.w, cb, acc[4 * z4 + 3]);
    }
    dcp_s[warp * IZ + lane] = dck;
    __syncthreads();
    // ---- phase b: dc of the sample (fixed order over k), then the owner-computes sweep for dT
    if (k == 0 && valid) {
      float t = 0.f;
      for (int kk = 0; kk < p.KR; ++kk) t += dcp_s[(s * p.KR + kk) * IZ + lane];
      p.dc[b * IZ + lane] = t;
    }
    for (int t0 = 0; t0 < R * 16; t0 += 32) {
      const uint32_t v = (t0 + lane < R * 16) ? code_s[t0 + lane] : 0xFFFF0000u;
      uint32_t m = __ballot_sync(FULL, (v >> 16) == (uint32_t)warp);
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t vv = __shfl_sync(FULL, v, j);
        const int t = t0 + j;
        float* d = dT_s + (vv & 0xFFFFu) * IZ + lane;
        *d = fmaf(coef_s[t], u_s[(t >> 4) * IZ + lane], *d);
      }
    }
    __syncthreads();
  }
  // ---- per-CTA partials: dT as is, dW combined over the warps that share k (fixed order), stored in (z, c, k) order
  float* part = p.part + (size_t)blockIdx.x * ((size_t)p.Kc * IZ + (size_t)IZ * IZ * p.KR);
  for (int i = threadIdx.x; i < p.Kc * 8; i += blockDim.x)
    reinterpret_cast<float4*>(part)[i] = reinterpret_cast<const float4*>(dT_s)[i];
  float* scr = T_s;                                  // all reads of T_s are behind the loop's last barrier
#pragma unroll
  for (int z = 0; z < IZ; ++z) scr[(warp * IZ + z) * IZ + lane] = acc[z];
  __syncthreads();
  float* pw = part + (size_t)p.Kc * IZ;
  for (int i = threadIdx.x; i < IZ * IZ * p.KR; i += blockDim.x) {
    const int kk = i % p.KR, zc = i / p.KR;
    float t = 0.f;
    for (int ss = 0; ss < p.SPI; ++ss) t += scr[(ss * p.KR + kk) * IZ * IZ + zc];
    pw[i] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Same contract, production version: the table gradient  dT = C^T U  on the tensor cores.
//
//   C (rows (b, k) x Kc codes) holds the 1 + N coefficients of every row at its code columns, U stacks the u vectors, so
//   dT[code][z] = sum_rows C[row][code] u[row][z] is a (Kc x rows) . (rows x 32) product.  One tcgen05.mma (kind::f16,
//   M = 128 codes, N = 32 features, K = 16 rows) per 128-code tile and operand-split pass, accumulated in TMEM (Kc/128 x 32
//   columns) over ALL rows of the CTA and read back once at the end: no scatter, no sort, no floating-point atomics.
//   Operands are bf16 (hi, lo) pairs of coef and u; passes hi.hi + hi.lo + lo.hi, fp32 accumulation (the same 16-bit
//   significand arithmetic as every tcgen05 GEMM of the encoder; |error| <= 2^-16 |coef| |u| per product).
//
// CTA = PW = KR * SPI producer warps + 1 MMA warp.  Producer warp w owns K-slot w: per step it turns ONE row (b, k = w % KR)
// into its column of the operands - the <= 16 non-zero entries of C^T (2-byte stores into a zero-initialised SWIZZLE_32B
// K-major tile, duplicates of a code merged first; the entries are zeroed again when the stage is recycled) and column w of
// U^T - and does the small dense parts on the CUDA cores (du = sum_n coef_n T[idx_n], dc = W_k^T du, dW_k += du (x) c).
// Stages are handed to the MMA warp through mbarriers (full: PW arrivals; empty: tcgen05.commit), so warps run up to NST
// steps apart and no CTA-wide barrier sits in the loop.
constexpr int NST = 3;

struct NceBwdTcP {
  NceBwdP b;
  int kpad;                   // Kc rounded up to 128
  int PW;                     // producer warps
  int tmem_cols;              // power of two >= kpad / 128 * 32
  uint32_t stage_region;      // bytes of the operand stages (>= PW * 4 KB: reused as the dW scratch at the end)
};

// byte offset of element (row r, k) inside a SWIZZLE_32B K-major slab [rows][16 bf16]: 16-byte chunk index ^= bit 2 of r
__device__ __forceinline__ uint32_t kelem_off(int r, int k) { return (uint32_t)(r * 32 + (((k >> 3) ^ ((r >> 2) & 1)) << 4) + (k & 7) * 2); }

__device__ __forceinline__ void split_store(unsigned char* hi, unsigned char* lo, uint32_t off, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
  *reinterpret_cast<__nv_bfloat16*>(lo + off) = l;
}

__global__ void __launch_bounds__(544, 1) nce_idx_bwd_tc_kernel(NceBwdTcP q) {
  const NceBwdP& p = q.b;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int PW = q.PW, kpad = q.kpad, MT = kpad >> 7;
  const uint32_t A_BYTES = (uint32_t)kpad * 32, STAGE = 2 * A_BYTES + 2048;       // A_hi | A_lo | B_hi (1 KB) | B_lo (1 KB)
  uint8_t* stages = smem;                                                           // [NST][STAGE]
  float* T_s = reinterpret_cast<float*>(stages + q.stage_region);                   // [Kc][32] (read lane = feature: no padding needed)
  float* W_s = T_s + (size_t)p.Kc * IZ;                                             // [KR][32 c][TST]  W_s[k][c][z] = W[z][c][k]
  float* coef_s = W_s + (size_t)p.KR * IZ * TST;                                    // [PW][16]
  int* code_s = reinterpret_cast<int*>(coef_s + PW * 16);                           // [PW][16]
  float* du_s = reinterpret_cast<float*>(code_s + PW * 16);                         // [PW][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(du_s + PW * IZ);                     // full[NST], empty[NST], done, table
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 2);
  const uint32_t bar0 = smem_u32(bars);
#define BARF(i) (bar0 + 8u * (uint32_t)(i))
#define BARE(i) (bar0 + 8u * (uint32_t)(NST + (i)))
#define BARD (bar0 + 8u * (uint32_t)(2 * NST))
#define BART (bar0 + 8u * (uint32_t)(2 * NST + 1))
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_begin = clock64();

  // ------------------------------------------------------------------ set-up
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      mbar_init(BARF(i), PW);
      mbar_init(BARE(i), 1);
    }
    mbar_init(BARD, 1);
    mbar_init(BART, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {                            // the table in its global layout: a few large bulk copies, overlapping
    const uint32_t bytes = (uint32_t)p.Kc * IZ * 4;  // the zero fill, the W staging and the first operand columns
    mbar_expect_tx(BART, bytes);
    for (uint32_t o = 0; o < bytes; o += 16384) {
      const uint32_t nb = min(16384u, bytes - o);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(T_s) + o),
                   "l"(reinterpret_cast<const char*>(p.T) + o), "r"(nb), "r"(BART)
                   : "memory");
    }
  }
  if (warp == PW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(q.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (int)(NST * STAGE / 16); i += blockDim.x) reinterpret_cast<uint4*>(stages)[i] = make_uint4(0, 0, 0, 0);
  stage_w<false>(W_s, p.W, p.KR);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t nit = (p.B + (int64_t)gridDim.x * p.SPI - 1) / ((int64_t)gridDim.x * p.SPI);
  const long long t_setup_end = clock64();
  long long tp_coef = 0, tp_wait = 0, tp_build = 0, tp_du = 0, tp_dc = 0;

  if (warp == PW) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc<32>(true, true);
      constexpr uint32_t hi_c = (256u >> 4) | (1u << 14) | (6u << 29);      // SBO 256 B, descriptor version 1, SWIZZLE_32B
      constexpr uint32_t lo_c = 1u << 16;
      const uint32_t sbase = smem_u32(stages) & 0x3FFFFu;
      for (int64_t it = 0; it < nit; ++it) {
        const int st = (int)(it % NST);
        mbar_wait(BARF(st), (uint32_t)((it / NST) & 1));
        tcgen05_fence_after();
        const uint32_t a_hi = sbase + (uint32_t)st * STAGE, a_lo = a_hi + A_BYTES, b_hi = a_lo + A_BYTES, b_lo = b_hi + 1024;
        const uint64_t dbh = pack_desc(lo_c | (b_hi >> 4), hi_c), dbl = pack_desc(lo_c | (b_lo >> 4), hi_c);
        for (int m = 0; m < MT; ++m) {
          const uint64_t dah = pack_desc(lo_c | ((a_hi + (uint32_t)m * 4096) >> 4), hi_c);
          const uint64_t dal = pack_desc(lo_c | ((a_lo + (uint32_t)m * 4096) >> 4), hi_c);
          const uint32_t d = tmem_base + (uint32_t)(m * 32);
          umma_bf16(d, dah, dbh, idesc, it > 0 ? 1u : 0u);
          umma_bf16(d, dah, dbl, idesc, 1u);
          umma_bf16(d, dal, dbh, idesc, 1u);
        }
        umma_commit(BARE(st));
      }
      umma_commit(BARD);
    }
    __syncwarp();
  } else {
    // ============================== producers ==============================
    const int k = warp % p.KR, s = warp / p.KR, KR = p.KR, N = p.N;
    float acc[IZ];                                   // acc[z] = dW[z][c = lane][k]
#pragma unroll
    for (int z = 0; z < IZ; ++z) acc[z] = 0.f;
    const float gs = p.scale * (p.g_dev ? __ldg(p.g_dev) : 1.f);
    int old_code[NST];                               // lane n: the code whose entry it may have written into stage i
#pragma unroll
    for (int i = 0; i < NST; ++i) old_code[i] = -1;
    float* cfs = coef_s + warp * 16;
    int* cds = code_s + warp * 16;
    float* dus = du_s + warp * IZ;
    const float4* w4 = reinterpret_cast<const float4*>(W_s + (k * IZ + lane) * TST);
    const bool cand = lane < 16 && lane <= N;        // lane n < 16 handles candidate n (0 = the positive)
    // global reads of one row, fetched one step ahead: score, code, lse (candidate lanes), u[z = lane], c[c = lane]
    float sc_n = 0.f, lse_n = 0.f, uz_n = 0.f, cb_n = 0.f;
    int code_n = 0;
    auto fetch = [&](int64_t it) {
      const int64_t b64 = (it * gridDim.x + blockIdx.x) * p.SPI + s;
      sc_n = lse_n = uz_n = cb_n = 0.f;
      code_n = 0;
      if (it >= nit || b64 >= p.B) return;
      const int b = (int)b64, r = b * KR + k;
      uz_n = ldg_now(p.u + r * IZ + lane);
      cb_n = ldg_now(p.c + b * IZ + lane);
      if (cand) {
        const int n = lane;
        sc_n = ldg_now((n == 0) ? p.pos + r : p.neg + r * N + n - 1);
        code_n = (int)ldg_now(n == 0 ? p.idx_pos + r : p.idx_neg + (b * N + (n - 1)) * KR + k);
        lse_n = ldg_now(p.lse + r);
      }
    };
    fetch(0);
    bool t_ready = false;
#pragma unroll 1
    for (int64_t it = 0; it < nit; ++it) {
      long long tq = clock64();
      const int st = (int)(it % NST);
      const int64_t b64 = (it * gridDim.x + blockIdx.x) * p.SPI + s;
      const bool valid = b64 < p.B;
      const int b = (int)b64, r = b * KR + k;
      const float sc = sc_n, lse_r = lse_n, uz = uz_n, cb = cb_n;
      const int code_raw = code_n;
      fetch(it + 1);                                 // next row's reads are in flight during this step
      float cf = 0.f;
      int code = -1 - lane;                          // inactive lanes: distinct keys
      if (cand && valid) {
        const int n = lane;
        code = min(max(code_raw, 0), p.Kc - 1);
        cf = gs * (expf(sc - lse_r) - (n == 0 ? 1.f : 0.f));
        if (n == 0) p.coef_pos[r] = cf; else p.coef_neg[(b * N + (n - 1)) * KR + k] = cf;
      }
      if (lane < 16) {
        cfs[lane] = cf;
        cds[lane] = code >= 0 ? code : 0;
      }
      __syncwarp();
      { const long long t1 = clock64(); tp_coef += t1 - tq; tq = t1; }
      // ---- stage hand-over: wait until the MMAs that read this stage NST steps ago have retired, un-write our old entries
      uint8_t* sg = stages + (size_t)st * STAGE;
      if (it >= NST) {
        mbar_wait(BARE(st), (uint32_t)(((it / NST) - 1) & 1));
        { const long long t1 = clock64(); tp_wait += t1 - tq; tq = t1; }
        int oc = old_code[0];
#pragma unroll
        for (int i = 1; i < NST; ++i) oc = (st == i) ? old_code[i] : oc;
        if (oc >= 0) {
          const uint32_t off = kelem_off(oc, warp);
          *reinterpret_cast<uint16_t*>(sg + off) = 0;
          *reinterpret_cast<uint16_t*>(sg + A_BYTES + off) = 0;
        }
      }
      __syncwarp();
      // ---- column `warp` of C^T: merge the candidates that share a code, one 2-byte store pair per distinct code
      if (lane < 16) {
        const uint32_t mm = __match_any_sync(0x0000FFFFu, code);
        float tot = cf;
        if (code >= 0 && (mm & (mm - 1)) && (__ffs(mm) - 1 == lane)) {
          uint32_t rest = mm & ~(1u << lane);
          while (rest) {
            const int j = __ffs(rest) - 1;
            rest &= rest - 1;
            tot += cfs[j];
          }
        }
        if (code >= 0 && (__ffs(mm) - 1 == lane)) split_store(sg, sg + A_BYTES, kelem_off(code, warp), tot);
      }
#pragma unroll
      for (int i = 0; i < NST; ++i) old_code[i] = (st == i) ? code : old_code[i];
      // ---- column `warp` of U^T (row = feature = lane); zeros for rows past the batch
      split_store(sg + 2 * A_BYTES, sg + 2 * A_BYTES + 1024, kelem_off(lane, warp), uz);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(BARF(st));
      { const long long t1 = clock64(); tp_build += t1 - tq; tq = t1; }
      if (!t_ready) {
        mbar_wait(BART, 0);
        t_ready = true;
      }
      // ---- du[z = lane] = sum_n coef_n T[code_n][z]
      float du0 = 0.f, du1 = 0.f;
#pragma unroll
      for (int n4 = 0; n4 < 4; ++n4) {
        const int4 cd = *reinterpret_cast<const int4*>(cds + n4 * 4);
        const float4 c4 = *reinterpret_cast<const float4*>(cfs + n4 * 4);
        du0 = fmaf(c4.x, T_s[cd.x * IZ + lane], du0);
        du1 = fmaf(c4.y, T_s[cd.y * IZ + lane], du1);
        du0 = fmaf(c4.z, T_s[cd.z * IZ + lane], du0);
        du1 = fmaf(c4.w, T_s[cd.w * IZ + lane], du1);
      }
      dus[lane] = du0 + du1;
      __syncwarp();
      { const long long t1 = clock64(); tp_du += t1 - tq; tq = t1; }
      // ---- dc_k[c = lane] = sum_z W[z][c][k] du[z]  (summed over k by the reduction kernel);  dW[z][c][k] += du[z] c_b[c]
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
      for (int z4 = 0; z4 < 8; ++z4) {
        const float4 d4 = *reinterpret_cast<const float4*>(dus + z4 * 4);
        const float4 w = w4[z4];
        d0 = fmaf(w.x, d4.x, d0);
        d1 = fmaf(w.y, d4.y, d1);
        d2 = fmaf(w.z, d4.z, d2);
        d3 = fmaf(w.w, d4.w, d3);
        acc[4 * z4] = fmaf(d4.x, cb, acc[4 * z4]);
        acc[4 * z4 + 1] = fmaf(d4.y, cb, acc[4 * z4 + 1]);
        acc[4 * z4 + 2] = fmaf(d4.z, cb, acc[4 * z4 + 2]);
        acc[4 * z4 + 3] = fmaf(d4.w, cb, acc[4 * z4 + 3]);
      }
      if (valid) p.dcp[r * IZ + lane] = (d0 + d1) + (d2 + d3);
      __syncwarp();
      { const long long t1 = clock64(); tp_dc += t1 - tq; tq = t1; }
    }
    if (!t_ready) mbar_wait(BART, 0);
    const long long t_loop_end = clock64();
    // ---- all MMAs retired: TMEM holds dT of this CTA
    mbar_wait(BARD, 0);
    tcgen05_fence_after();
    float* part = p.part + (size_t)blockIdx.x * ((size_t)p.Kc * IZ + (size_t)IZ * IZ * p.KR);
    {
      // warp w reads lane quadrant w & 3 of the code tiles w >> 2, w >> 2 + PW / 4, ...
      for (int m = warp >> 2; m < MT; m += max(1, PW >> 2)) {
        if ((warp >> 2) >= max(1, PW >> 2)) break;
        uint32_t rr[32];
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(m * 32), rr);
        tmem_ld_wait();
        const int code = m * 128 + (warp & 3) * 32 + lane;
        if (code < p.Kc) {
          float4* o = reinterpret_cast<float4*>(part + (size_t)code * IZ);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        }
      }
    }
    // ---- dW: combine the warps that share k through the (now idle) operand stages
    float* scr = reinterpret_cast<float*>(stages);
    asm volatile("bar.sync 9, %0;" ::"r"(PW * 32) : "memory");
#pragma unroll
    for (int z = 0; z < IZ; ++z) scr[(warp * IZ + z) * IZ + lane] = acc[z];
    asm volatile("bar.sync 9, %0;" ::"r"(PW * 32) : "memory");
    float* pw = part + (size_t)p.Kc * IZ;
    for (int zc = warp * 32 + lane; zc < IZ * IZ; zc += PW * 32)
      for (int kk = 0; kk < p.KR; ++kk) {
        float t = 0.f;
        for (int ss = 0; ss < p.SPI; ++ss) t += scr[(ss * p.KR + kk) * IZ * IZ + zc];
        pw[zc * p.KR + kk] = t;
      }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      g_idx_dbg[8] = t_setup_end - t_begin; g_idx_dbg[9] = t_loop_end - t_setup_end; g_idx_dbg[10] = clock64() - t_loop_end;
      g_idx_dbg[11] = tp_coef; g_idx_dbg[12] = tp_wait; g_idx_dbg[13] = tp_build; g_idx_dbg[14] = tp_du; g_idx_dbg[15] = tp_dc;
    }
  }
#undef BARF
#undef BARE
#undef BARD
#undef BART
  tcgen05_fence_before();
  __syncthreads();
  if (warp == PW) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(q.tmem_cols) : "memory");
  }
}

// out_a[i] += sum_p part[p][i] (i < na);  out_b[i] += sum_p part[p][na + i] (i < nb);  na, nb multiples of 4.
// Blocks [0, nblk_tab): 512 threads = 32 float4 columns x 16 groups of partials (<= RP_MAX loads per thread, all in flight
// at once), fixed summation order.  Blocks from nblk_tab on: dc[b][c] = sum_k dcp[(b * KR + k)][c] (one float4 per thread).
constexpr int RP_MAX = 12;    // partial tables per thread: up to 16 * 12 = 192 CTAs
__global__ void __launch_bounds__(512) reduce_partials_kernel(const float* __restrict__ part, int nparts, size_t stride,
                                                              float* __restrict__ out_a, int na, float* __restrict__ out_b, int nb,
                                                              int nblk_tab, const float* __restrict__ dcp, float* __restrict__ dc,
                                                              int B, int KR) {
  if ((int)blockIdx.x >= nblk_tab) {
    const int i = (blockIdx.x - nblk_tab) * 512 + threadIdx.x;      // (b, c4)
    if (i < B * 8) {
      const int b = i >> 3, c4 = i & 7;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < KR; ++k) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dcp + ((size_t)b * KR + k) * IZ) + c4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      reinterpret_cast<float4*>(dc + (size_t)b * IZ)[c4] = a;
    }
    return;
  }
  __shared__ float4 red[16][32];
  const int cx = threadIdx.x & 31, pg = threadIdx.x >> 5;
  const int col4 = blockIdx.x * 32 + cx;
  const int ncol4 = (na + nb) >> 2;
  float4 v[RP_MAX];
#pragma unroll
  for (int j = 0; j < RP_MAX; ++j) {
    const int q = pg + 16 * j;
    v[j] = (col4 < ncol4 && q < nparts) ? __ldg(reinterpret_cast<const float4*>(part + (size_t)q * stride) + col4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 a = v[0];
#pragma unroll
  for (int j = 1; j < RP_MAX; ++j) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
  red[pg][cx] = a;
  __syncthreads();
  if (pg == 0 && col4 < ncol4) {
#pragma unroll
    for (int q = 1; q < 16; ++q) {
      const float4 t = red[q][cx];
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    float4* o = (col4 * 4 < na) ? reinterpret_cast<float4*>(out_a) + col4 : reinterpret_cast<float4*>(out_b) + (col4 - (na >> 2));
    float4 c = *o;
    c.x += a.x; c.y += a.y; c.z += a.z; c.w += a.w;
    *o = c;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// out[idx[i]] += rows[i]  without atomics: CTA j owns the codes [j * cpc, (j + 1) * cpc), every warp scans its slice of
// the index list and keeps warp-private sums, combined in warp order (deterministic).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) table_scatter_add_kernel(const int64_t* __restrict__ idx, int64_t n, const float* __restrict__ rows,
                                                                int zd, int Kc, int cpc, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];        // [warps][cpc][zd]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, NW = blockDim.x >> 5;
  const int c0 = blockIdx.x * cpc;
  for (int i = threadIdx.x; i < NW * cpc * zd; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float* mine = sm + (size_t)warp * cpc * zd;
  const int64_t per = ((n + NW - 1) / NW + 31) / 32 * 32;
  const int64_t beg = warp * per, end = min(n, beg + per);
  for (int64_t i0 = beg; i0 < end; i0 += 32 * 8) {
    int codes[8];                                    // eight independent index loads in flight per lane
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int64_t i = i0 + q * 32 + lane;
      codes[q] = i < end ? (int)idx[i] : -1;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t m = __ballot_sync(FULL, codes[q] >= c0 && codes[q] < c0 + cpc);
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const int cj = __shfl_sync(FULL, codes[q], j) - c0;
        for (int f = lane; f < zd; f += 32) mine[cj * zd + f] += __ldg(rows + (i0 + q * 32 + j) * zd + f);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cpc * zd; e += blockDim.x) {
    if (c0 + e / zd >= Kc) break;
    float t = 0.f;
    for (int w = 0; w < NW; ++w) t += sm[(size_t)w * cpc * zd + e];
    out[(size_t)c0 * zd + e] += t;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward of the quantizer in index form: dz for every vector, dE
// ---------------------------------------------------------------------------------------------------------------------
struct IdxSeg {
  int64_t row0, rows;
  int kind;                   // 1: dy rows a[(p - row0) * 32 ..];  2: a[p - row0] * b[((p - row0) / per_b * KR + (p - row0) % KR) * 32 ..]
  const float *a, *b;
  int per_b, KR;
  unsigned long long m_per_b, m_kr;   // ceil(2^40 / d): x / d == (x * m) >> 40 for x < 2^30, d < 1024 (two instructions)
};
__device__ __forceinline__ uint32_t fast_div(uint32_t x, unsigned long long m) { return (uint32_t)(((unsigned long long)x * m) >> 40); }
struct VqIdxBwdP {
  const float *x, *E, *J, *gl;
  const int64_t* idx;
  int64_t n;
  int Kc;
  float beta;
  int squared, nseg, j_smem;
  uint32_t kind1_mask;        // bit s: segment s holds dense rows (kind 1)
  IdxSeg seg[4];
  float *dx, *dE, *dT;
};

constexpr int VCH = 512;      // vectors per CTA pass == threads

// per-vector inputs of one pass, fetched one pass ahead (their global-load latency hides behind phases 2 .. J)
template <int DS>
struct VqVec {
  int code;
  uint32_t src;
  float cf, gl, x[DS];
};

template <int DS>
__device__ __forceinline__ VqVec<DS> vq_vec_load(const VqIdxBwdP& p, int64_t v) {
  VqVec<DS> r;
  r.code = 0xFFFF;
  r.src = 0xFFFFFFFFu;
  r.cf = 0.f;
  r.gl = 0.f;
#pragma unroll
  for (int d = 0; d < DS; ++d) r.x[d] = 0.f;
  if (v < p.n) {
    r.code = (int)ldg_now(p.idx + v);          // (clamped by the consumer: nothing here may wait for a load)
    r.gl = p.gl ? ldg_now(p.gl + v) : 0.f;
#pragma unroll
    for (int d = 0; d < DS; ++d) r.x[d] = ldg_now(p.x + v * DS + d);
    const uint32_t v32 = (uint32_t)v;                  // (n < 2^31, checked by the host side)
#pragma unroll
    for (int sgi = 0; sgi < 4; ++sgi) {                // static indices: the segment descriptors stay in the constant bank
      const IdxSeg& sg = p.seg[sgi];
      const uint32_t pl = v32 - (uint32_t)sg.row0;
      if (sgi < p.nseg && pl < (uint32_t)sg.rows) {
        if (sg.kind == 1) {
          r.src = pl | ((uint32_t)sgi << 30);
          r.cf = 1.f;
        } else {
          r.src = ((pl / (uint32_t)sg.per_b) * (uint32_t)sg.KR + pl % (uint32_t)sg.KR) | ((uint32_t)sgi << 30);
          r.cf = ldg_now(sg.a + pl);
        }
      }
    }
  }
  return r;
}

template <int DS>
__global__ void __launch_bounds__(VCH, 1) vq_idx_bwd_kernel(VqIdxBwdP p) {
  constexpr int JROW = IZ * DS;                      // floats of J per code
  constexpr int JST = JROW + 4;                      // padded stride: consecutive codes start 4 banks apart
  constexpr int ES = 4;                              // padded code stride of the shared-memory codebook
  extern __shared__ __align__(16) unsigned char smraw[];
  float* J_s = reinterpret_cast<float*>(smraw);                                        // [Kc][JST] (if j_smem)
  float* E_s = J_s + (p.j_smem ? (size_t)p.Kc * JST : 0);                                // [Kc][4]
  float* gdq_s = E_s + (size_t)p.Kc * ES;                                                // [VCH][DS]  g_loss * dloss/dq
  float* coef_s = gdq_s + VCH * DS;                                                      // [VCH]      scale of the source row
  uint32_t* src_s = reinterpret_cast<uint32_t*>(coef_s + VCH);                           // [VCH]      source row | segment << 30
  int* hist_s = reinterpret_cast<int*>(src_s + VCH);                                     // [Kc]
  int* start_s = hist_s + p.Kc;                                                          // [Kc + 1]
  int* cur_s = start_s + p.Kc + 1;                                                       // [Kc]
  int* wsum_s = cur_s + p.Kc;                                                            // [32]
  uint16_t* code_s = reinterpret_cast<uint16_t*>(wsum_s + 32);                           // [VCH]
  uint16_t* perm_s = code_s + VCH;                                                       // [VCH]
  uint64_t* bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(perm_s + VCH) + 7) & ~(uintptr_t)7);
  const float** segbase_s = reinterpret_cast<const float**>(bar + 1);                    // [4] source matrix of each segment
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_a = s_u32(bar);
  if (tid < 4) segbase_s[tid] = tid < p.nseg ? (p.seg[tid].kind == 1 ? p.seg[tid].a : p.seg[tid].b) : nullptr;
  if (p.j_smem) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(p.Kc * JROW * 4)) : "memory");
    __syncthreads();
    // one bulk copy per code row (JROW * 4 bytes, a multiple of 16) into the padded layout; overlaps the first pass
    for (int kx = tid; kx < p.Kc; kx += VCH)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(J_s + (size_t)kx * JST)),
                   "l"(p.J + (size_t)kx * JROW), "r"((uint32_t)(JROW * 4)), "r"(bar_a)
                   : "memory");
  }
  for (int i = tid; i < p.Kc; i += VCH) {
    hist_s[i] = 0;
#pragma unroll
    for (int d = 0; d < ES; ++d) E_s[i * ES + d] = d < DS ? __ldg(p.E + i * DS + d) : 0.f;
  }
  float accE[2][DS];                                 // thread t owns the codes t and t + VCH (Kc <= 2 * VCH)
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int d = 0; d < DS; ++d) accE[a][d] = 0.f;
  bool j_ready = !p.j_smem;
  const int64_t nchunk = (p.n + VCH - 1) / VCH;
  const long long tb0 = clock64();
  long long tb_p1 = 0, tb_scan = 0, tb_sc = 0, tb_ld = 0, tb_j = 0, tb_p4 = 0;
  VqVec<DS> cur = vq_vec_load<DS>(p, (int64_t)blockIdx.x * VCH + tid);
  __syncthreads();
  const long long tb1 = clock64();
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    long long tq = clock64();
    // ---- phase 1: thread == vector: loss gradients, source row of the upscaled-side gradient, histogram
    {
      float g[DS];
#pragma unroll
      for (int d = 0; d < DS; ++d) g[d] = 0.f;
      const bool have = ch * VCH + tid < p.n;
      const int code1 = have ? min(max(cur.code, 0), p.Kc - 1) : 0xFFFF;
      if (have) {
        float df[DS];
        float ssq = 0.f;
#pragma unroll
        for (int d = 0; d < DS; ++d) {
          df[d] = E_s[code1 * ES + d] - cur.x[d];
          const float t = df[d] + 1e-5f;
          ssq = fmaf(t, t, ssq);
        }
        const float inv = p.squared ? 0.f : 1.f / sqrtf(ssq);
#pragma unroll
        for (int d = 0; d < DS; ++d) g[d] = cur.gl * (p.squared ? 2.f * df[d] : (df[d] + 1e-5f) * inv);
        atomicAdd(&hist_s[code1], 1);
      }
#pragma unroll
      for (int d = 0; d < DS; ++d) gdq_s[tid * DS + d] = g[d];
      coef_s[tid] = cur.cf;
      src_s[tid] = cur.src;
      code_s[tid] = (uint16_t)code1;
    }
    // next pass's inputs: in flight until the next phase 1
    cur = vq_vec_load<DS>(p, (ch + gridDim.x) * VCH + tid);
    __syncthreads();
    { const long long t1 = clock64(); tb_p1 += t1 - tq; tq = t1; }
    // ---- phase 2: exclusive scan of the histogram (Kc <= 2 * VCH bins, two per thread)
    {
      const int b0 = 2 * tid, b1 = 2 * tid + 1;
      const int h0 = b0 < p.Kc ? hist_s[b0] : 0, h1 = b1 < p.Kc ? hist_s[b1] : 0;
      int incl = h0 + h1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) wsum_s[warp] = incl;
      __syncthreads();
      int base = 0;
      for (int w = 0; w < warp; ++w) base += wsum_s[w];
      const int ex = base + incl - (h0 + h1);
      if (b0 < p.Kc) { start_s[b0] = ex; cur_s[b0] = ex; hist_s[b0] = 0; }
      if (b1 < p.Kc) { start_s[b1] = ex + h0; cur_s[b1] = ex + h0; hist_s[b1] = 0; }
    }
    __syncthreads();
    { const long long t1 = clock64(); tb_scan += t1 - tq; tq = t1; }
    // ---- phase 3: scatter the vector slots into code order
    {
      const int code = code_s[tid];
      if (code != 0xFFFF) perm_s[atomicAdd(&cur_s[code], 1)] = (uint16_t)tid;
    }
    { const long long t1 = clock64(); tb_sc += t1 - tq; tq = t1; }
    // ---- phase J (needs phase 1 only): a quad of lanes per vector: dz = coef * J[code]^T v - beta * g_loss * dloss/dq.
    // All source rows of the four passes are requested first, then consumed.
    const int ql = tid & 3, qid = tid >> 2;
    float4 va[4], vb[4];
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const uint32_t src = src_s[pass * (VCH / 4) + qid];
      va[pass] = vb[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src != 0xFFFFFFFFu) {
        const float4* row = reinterpret_cast<const float4*>(segbase_s[src >> 30] + (size_t)(src & 0x3FFFFFFFu) * IZ + ql * 8);
        va[pass] = __ldg(row);
        vb[pass] = __ldg(row + 1);
        if (p.dT && ((p.kind1_mask >> (src >> 30)) & 1u)) {
          // gradient of the table gather y = T[idx] of the dense-row (left) blocks: dT[code] += dy, 32 bytes per lane as two
          // fire-and-forget vector reductions (these rows are ~6 % of the vectors; the L2 atomic units work in the background)
          float* t = p.dT + (size_t)min((int)code_s[pass * (VCH / 4) + qid], p.Kc - 1) * IZ + ql * 8;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(t), "f"(va[pass].x), "f"(va[pass].y), "f"(va[pass].z), "f"(va[pass].w) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(t + 4), "f"(vb[pass].x), "f"(vb[pass].y), "f"(vb[pass].z), "f"(vb[pass].w) : "memory");
        }
      }
    }
    if (!j_ready) {
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
      j_ready = true;
    }
    { const long long t1 = clock64(); tb_ld += t1 - tq; tq = t1; }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int i = pass * (VCH / 4) + qid;
      const int64_t v = ch * VCH + i;
      const float vv[8] = {va[pass].x, va[pass].y, va[pass].z, va[pass].w, vb[pass].x, vb[pass].y, vb[pass].z, vb[pass].w};
      const int code = min((int)code_s[i], p.Kc - 1);
      const float* jr = p.j_smem ? (J_s + (size_t)code * JST + ql * 8 * DS) : (p.J + (size_t)code * JROW + ql * 8 * DS);
      float gz[DS];
#pragma unroll
      for (int d = 0; d < DS; ++d) gz[d] = 0.f;
#pragma unroll
      for (int q4 = 0; q4 < 2 * DS; ++q4) {
        const float4 t = *reinterpret_cast<const float4*>(jr + q4 * 4);
        const float jv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int f = q4 * 4 + e;                  // flat index (feature t, component d) = t * DS + d
          gz[f % DS] = fmaf(vv[f / DS], jv[e], gz[f % DS]);
        }
      }
#pragma unroll
      for (int d = 0; d < DS; ++d) {
        gz[d] += __shfl_xor_sync(FULL, gz[d], 1);
        gz[d] += __shfl_xor_sync(FULL, gz[d], 2);
      }
      if (v < p.n && ql < DS) {
        float mine = gz[0];
#pragma unroll
        for (int d = 1; d < DS; ++d) mine = (ql == d) ? gz[d] : mine;
        p.dx[v * DS + ql] = fmaf(coef_s[i], mine, -p.beta * gdq_s[i * DS + ql]);
      }
    }
    __syncthreads();
    { const long long t1 = clock64(); tb_j += t1 - tq; tq = t1; }
    // ---- phase 4: thread == code: sequential sum over the code's vectors of this pass (no atomics)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int code = tid + a * VCH;
      if (code < p.Kc) {
        const int e = cur_s[code];
        for (int t = start_s[code]; t < e; ++t) {
          const int i = perm_s[t];
#pragma unroll
          for (int d = 0; d < DS; ++d) accE[a][d] += gdq_s[i * DS + d];
        }
      }
    }
    __syncthreads();
    { const long long t1 = clock64(); tb_p4 += t1 - tq; tq = t1; }
  }
  if (p.dE) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int code = tid + a * VCH;
      if (code < p.Kc)
#pragma unroll
        for (int d = 0; d < DS; ++d)
          if (accE[a][d] != 0.f) atomicAdd(p.dE + code * DS + d, accE[a][d]);
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    g_idx_dbg[0] = tb1 - tb0; g_idx_dbg[1] = clock64() - tb1; g_idx_dbg[2] = tb_p1; g_idx_dbg[3] = tb_scan; g_idx_dbg[4] = tb_sc;
    g_idx_dbg[5] = tb_ld; g_idx_dbg[6] = tb_j; g_idx_dbg[7] = tb_p4;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Production split of the quantizer backward (the single kernel above is kept as VQCPC_VQ_BWD_ALGO=1 / fallback):
//   vq_idx_dz_kernel : pure stream, a quad of lanes per vector, no shared-memory staging and no CTA barrier after set-up:
//                      dz = coef * J[code]^T v - beta * g_loss * dloss/dq, plus the vector reductions of the dense rows into dT.
//                      Everything a vector needs from global memory (code, z, g_loss, coefficient, source row) is fetched one
//                      iteration ahead.
//   vq_idx_de_kernel : dE[code] += g_loss * dloss/dq by a counting sort of a large slab of vectors per CTA (8 per thread) and
//                      per-code sequential sums - small shared-memory footprint, so the slab is big enough (4096 vectors
//                      over <= 1024 codes) for the sort to pay.
// ---------------------------------------------------------------------------------------------------------------------
template <int DS, int NV4>
struct DzIn {                                        // everything one vector needs from global memory
  int code;
  uint32_t seg;                                      // 4 = no source row
  float gl, cf, x[DS];
  float4 vr[NV4];
};

template <int DS, bool JS, int LPV>      // JS: J staged in shared memory (compile-time, so that its reads are LDS, not generic
__global__ void __launch_bounds__(512, 1) vq_idx_dz_kernel(VqIdxBwdP p) {   // loads); LPV: lanes per vector (2 or 4)
  constexpr int JROW = IZ * DS, JST = JROW + 4, ES = 4;
  constexpr int FP = IZ / LPV;                       // features per lane
  constexpr int NV4 = FP / 4;                        // float4 loads of the source row per lane
  constexpr int NJ4 = FP * DS / 4;                   // float4 loads of J per lane
  constexpr int VPW = 32 / LPV;                      // vectors per warp and iteration
  extern __shared__ __align__(16) unsigned char smraw[];
  float* J_s = reinterpret_cast<float*>(smraw);                                        // [Kc][JST] (if JS)
  float* E_s = J_s + (JS ? (size_t)p.Kc * JST : 0);                                      // [Kc][4]
  uint64_t* bar = reinterpret_cast<uint64_t*>(E_s + (size_t)p.Kc * ES);
  IdxSeg* seg_s = reinterpret_cast<IdxSeg*>(bar + 1);                                    // [4] shared-memory copy (dynamic index)
  const int tid = threadIdx.x;
  const long long tz0 = clock64();
  const uint32_t bar_a = s_u32(bar);
  if (tid < 4) seg_s[tid] = p.seg[tid];
  if (JS) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)(p.Kc * JROW * 4)) : "memory");
    __syncthreads();
    for (int kx = tid; kx < p.Kc; kx += blockDim.x)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(J_s + (size_t)kx * JST)),
                   "l"(p.J + (size_t)kx * JROW), "r"((uint32_t)(JROW * 4)), "r"(bar_a)
                   : "memory");
  }
  for (int i = tid; i < p.Kc; i += blockDim.x) {
#pragma unroll
    for (int d = 0; d < ES; ++d) E_s[i * ES + d] = d < DS ? __ldg(p.E + i * DS + d) : 0.f;
  }
  __syncthreads();
  const int ql = tid & (LPV - 1);
  const int64_t ngroups = (int64_t)gridDim.x * (blockDim.x / LPV);
  int64_t v = (int64_t)blockIdx.x * (blockDim.x / LPV) + (tid / LPV);
  // The loop is unrolled by two with two register sets: a set is refilled right after it has been consumed and used again
  // two iterations later, so every global read has two iterations (> the DRAM latency) to land.
  auto fetch = [&](DzIn<DS, NV4>& in, int64_t vv) {
    in.seg = 4;
    in.cf = in.gl = 0.f;
    in.code = 0;
#pragma unroll
    for (int q = 0; q < NV4; ++q) in.vr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int d = 0; d < DS; ++d) in.x[d] = 0.f;
    if (vv >= p.n) return;
    in.code = (int)ldg_now(p.idx + vv);
    in.gl = p.gl ? ldg_now(p.gl + vv) : 0.f;
#pragma unroll
    for (int d = 0; d < DS; ++d) in.x[d] = ldg_now(p.x + vv * DS + d);
    const uint32_t v32 = (uint32_t)vv;
    // which segment (at most one) holds this vector: a short compare chain on constant-bank operands ...
    uint32_t sgi = 4, pl = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint32_t d = v32 - (uint32_t)p.seg[t].row0;
      if (t < p.nseg && d < (uint32_t)p.seg[t].rows) { sgi = t; pl = d; }
    }
    // ... then ONE copy of the source-row code, its descriptor read from shared memory
    if (sgi < 4) {
      const IdxSeg& sg = seg_s[sgi];
      uint32_t row = pl;
      in.cf = 1.f;
      const float* base = sg.a;
      if (sg.kind != 1) {
        const uint32_t bq = fast_div(pl, sg.m_per_b), kq = fast_div(pl, sg.m_kr);
        row = bq * (uint32_t)sg.KR + (pl - kq * (uint32_t)sg.KR);
        in.cf = ldg_now(sg.a + pl);
        base = sg.b;
      }
      in.seg = sgi;
      const float* src = base + row * IZ + ql * FP;
#pragma unroll
      for (int q = 0; q < NV4; ++q)
        asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(in.vr[q].x), "=f"(in.vr[q].y), "=f"(in.vr[q].z), "=f"(in.vr[q].w) : "l"(src + 4 * q));
    }
  };
  auto process = [&](const DzIn<DS, NV4>& in, int64_t vv) {
    const bool ok = vv < p.n;
    const int code = min(max(in.code, 0), p.Kc - 1);
    const uint32_t sgi = in.seg;
    float vf[FP];
#pragma unroll
    for (int q = 0; q < NV4; ++q) { vf[4 * q] = in.vr[q].x; vf[4 * q + 1] = in.vr[q].y; vf[4 * q + 2] = in.vr[q].z; vf[4 * q + 3] = in.vr[q].w; }
    // loss-side gradient of this vector (every lane of the group computes it: cheaper than a shuffle)
    float g[DS];
    {
      float df[DS];
      float ssq = 0.f;
#pragma unroll
      for (int d = 0; d < DS; ++d) {
        df[d] = E_s[code * ES + d] - in.x[d];
        const float t = df[d] + 1e-5f;
        ssq = fmaf(t, t, ssq);
      }
      const float inv = p.squared ? 0.f : 1.f / sqrtf(ssq);
#pragma unroll
      for (int d = 0; d < DS; ++d) g[d] = in.gl * (p.squared ? 2.f * df[d] : (df[d] + 1e-5f) * inv);
    }
    float gz[DS];
#pragma unroll
    for (int d = 0; d < DS; ++d) gz[d] = 0.f;
    if (sgi < 4) {
#pragma unroll
      for (int q4 = 0; q4 < NJ4; ++q4) {
        const float4 t = JS ? *reinterpret_cast<const float4*>(J_s + code * JST + ql * FP * DS + q4 * 4)
                            : __ldg(reinterpret_cast<const float4*>(p.J + (size_t)code * JROW + ql * FP * DS + q4 * 4));
        const float jv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int f = q4 * 4 + e;                  // flat index (feature t, component d) = t * DS + d
          gz[f % DS] = fmaf(vf[f / DS], jv[e], gz[f % DS]);
        }
      }
      if (ok && p.dT && ((p.kind1_mask >> sgi) & 1u)) {
        float* t = p.dT + (size_t)code * IZ + ql * FP;
#pragma unroll
        for (int q = 0; q < NV4; ++q)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(t + 4 * q), "f"(vf[4 * q]), "f"(vf[4 * q + 1]), "f"(vf[4 * q + 2]), "f"(vf[4 * q + 3]) : "memory");
      }
    }
#pragma unroll
    for (int d = 0; d < DS; ++d) {
#pragma unroll
      for (int o = 1; o < LPV; o <<= 1) gz[d] += __shfl_xor_sync(FULL, gz[d], o);
    }
    if (ok) {
#pragma unroll
      for (int d = 0; d < DS; ++d)
        if ((d & (LPV - 1)) == ql) p.dx[vv * DS + d] = fmaf(in.cf, gz[d], -p.beta * g[d]);
    }
  };
  DzIn<DS, NV4> s0, s1;
  fetch(s0, v);
  fetch(s1, v + ngroups);
  const long long tz1 = clock64();
  if (JS) {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
  }
  const long long tz2 = clock64();
  const int qw = (tid / LPV) & (VPW - 1);             // whole warps stay in the loop together (shuffles inside process):
  while (v - qw < p.n) {                              // v - qw is the first vector of the warp, the same on all 32 lanes
    process(s0, v);
    fetch(s0, v + 2 * ngroups);
    v += ngroups;
    if (!(v - qw < p.n)) break;
    process(s1, v);
    fetch(s1, v + 2 * ngroups);
    v += ngroups;
  }
  if (blockIdx.x == 0 && tid == 0) { g_idx_dbg[0] = tz1 - tz0; g_idx_dbg[1] = tz2 - tz1; g_idx_dbg[2] = clock64() - tz2; }
}

constexpr int DE_T = 512, DE_V = 8, DE_CH = DE_T * DE_V;     // dE kernel: threads, vectors per thread, slab
template <int DS>
__global__ void __launch_bounds__(DE_T, 2) vq_idx_de_kernel(VqIdxBwdP p) {
  constexpr int ES = 4;
  extern __shared__ __align__(16) unsigned char smraw[];
  float* E_s = reinterpret_cast<float*>(smraw);                                          // [Kc][4]
  float* gdq_s = E_s + (size_t)p.Kc * ES;                                                // [DE_CH][DS]
  int* hist_s = reinterpret_cast<int*>(gdq_s + DE_CH * DS);                              // [Kc]
  int* start_s = hist_s + p.Kc;                                                          // [Kc]
  int* cur_s = start_s + p.Kc;                                                           // [Kc]
  int* wsum_s = cur_s + p.Kc;                                                            // [32]
  uint16_t* code_s = reinterpret_cast<uint16_t*>(wsum_s + 32);                           // [DE_CH]
  uint16_t* perm_s = code_s + DE_CH;                                                     // [DE_CH]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < p.Kc; i += DE_T) {
    hist_s[i] = 0;
#pragma unroll
    for (int d = 0; d < ES; ++d) E_s[i * ES + d] = d < DS ? __ldg(p.E + i * DS + d) : 0.f;
  }
  float accE[2][DS];                                 // thread t owns the codes t and t + DE_T
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int d = 0; d < DS; ++d) accE[a][d] = 0.f;
  __syncthreads();
  const int64_t nchunk = (p.n + DE_CH - 1) / DE_CH;
  for (int64_t ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
    // ---- all global reads of the slab first (8 vectors per thread, coalesced), then the loss gradients
    int codes[DE_V];
    float gl[DE_V], x[DE_V][DS];
#pragma unroll
    for (int j = 0; j < DE_V; ++j) {
      const int64_t v = ch * DE_CH + j * DE_T + tid;
      codes[j] = -1;
      gl[j] = 0.f;
#pragma unroll
      for (int d = 0; d < DS; ++d) x[j][d] = 0.f;
      if (v < p.n) {
        codes[j] = (int)ldg_now(p.idx + v);
        gl[j] = p.gl ? ldg_now(p.gl + v) : 0.f;
#pragma unroll
        for (int d = 0; d < DS; ++d) x[j][d] = ldg_now(p.x + v * DS + d);
      }
    }
#pragma unroll
    for (int j = 0; j < DE_V; ++j) {
      const int i = j * DE_T + tid;
      int code = 0xFFFF;
      float g[DS];
#pragma unroll
      for (int d = 0; d < DS; ++d) g[d] = 0.f;
      if (ch * DE_CH + i < p.n) {
        code = min(max(codes[j], 0), p.Kc - 1);
        float df[DS];
        float ssq = 0.f;
#pragma unroll
        for (int d = 0; d < DS; ++d) {
          df[d] = E_s[code * ES + d] - x[j][d];
          const float t = df[d] + 1e-5f;
          ssq = fmaf(t, t, ssq);
        }
        const float inv = p.squared ? 0.f : 1.f / sqrtf(ssq);
#pragma unroll
        for (int d = 0; d < DS; ++d) g[d] = gl[j] * (p.squared ? 2.f * df[d] : (df[d] + 1e-5f) * inv);
        atomicAdd(&hist_s[code], 1);
      }
#pragma unroll
      for (int d = 0; d < DS; ++d) gdq_s[i * DS + d] = g[d];
      code_s[i] = (uint16_t)code;
    }
    __syncthreads();
    // ---- exclusive scan of the histogram (Kc <= 2 * DE_T bins, two per thread)
    {
      const int b0 = 2 * tid, b1 = 2 * tid + 1;
      const int h0 = b0 < p.Kc ? hist_s[b0] : 0, h1 = b1 < p.Kc ? hist_s[b1] : 0;
      int incl = h0 + h1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) wsum_s[warp] = incl;
      __syncthreads();
      int base = 0;
      for (int w = 0; w < warp; ++w) base += wsum_s[w];
      const int ex = base + incl - (h0 + h1);
      if (b0 < p.Kc) { start_s[b0] = ex; cur_s[b0] = ex; hist_s[b0] = 0; }
      if (b1 < p.Kc) { start_s[b1] = ex + h0; cur_s[b1] = ex + h0; hist_s[b1] = 0; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < DE_V; ++j) {
      const int i = j * DE_T + tid;
      const int code = code_s[i];
      if (code != 0xFFFF) perm_s[atomicAdd(&cur_s[code], 1)] = (uint16_t)i;
    }
    __syncthreads();
    // ---- thread == code: sequential sum over the code's vectors of the slab (no atomics)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int code = tid + a * DE_T;
      if (code < p.Kc) {
        const int e = cur_s[code];
        for (int t = start_s[code]; t < e; ++t) {
          const int i = perm_s[t];
#pragma unroll
          for (int d = 0; d < DS; ++d) accE[a][d] += gdq_s[i * DS + d];
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int code = tid + a * DE_T;
    if (code < p.Kc)
#pragma unroll
      for (int d = 0; d < DS; ++d)
        if (accE[a][d] != 0.f) atomicAdd(p.dE + code * DS + d, accE[a][d]);
  }
}

template <int DS>
size_t vq_idx_bwd_smem(int Kc, bool j_smem) {
  size_t s = j_smem ? (size_t)Kc * (IZ * DS + 4) * 4 : 0;
  s += (size_t)Kc * 4 * 4;                           // codebook
  s += (size_t)VCH * DS * 4 + VCH * 4 + VCH * 4;     // gdq, coef, src
  s += ((size_t)3 * Kc + 1 + 32) * 4;                // hist, start, cur, wsum
  s += VCH * 2 * 2;                                  // code, perm
  s = (s + 7) & ~(size_t)7;
  return s + 8 + 4 * 8 + 8;                          // mbarrier, segment bases, alignment slack
}

int nce_spi(int KR) { return KR >= 9 ? 1 : 16 / KR; }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
// kernel selection (test / tuning hooks; -1 = take the environment variable, else the default 0)
static int g_nce_bwd_algo = -1;     // 0: tcgen05 dT (production), 1: CUDA-core owner-computes kernel
static int g_vq_bwd_algo = -1;      // 0: streaming dz + counting-sort dE kernels (production), 1: single kernel
static int env_or(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

VQ_API int vqcpc_infonce_idx_set_bwd_algo(int algo) {
  VQ_CHECK_ARG(algo >= -1 && algo <= 1, "infonce_idx_set_bwd_algo: algo must be -1, 0 or 1");
  g_nce_bwd_algo = algo;
  return VQ_OK;
}
VQ_API int vqcpc_vq_idx_set_bwd_algo(int algo) {
  VQ_CHECK_ARG(algo >= -1 && algo <= 1, "vq_idx_set_bwd_algo: algo must be -1, 0 or 1");
  g_vq_bwd_algo = algo;
  return VQ_OK;
}

static int idx_path_check(const char* who, int N, int KR, int zd, int cd, int Kc) {
  VQ_CHECK_ARG(zd == IZ && cd == IZ, "%s: the index path is specialised for zd == cd == 32 (got %d, %d)", who, zd, cd);
  VQ_CHECK_ARG(N >= 1 && N <= 15, "%s: 1 <= N <= 15 negatives supported (got %d)", who, N);
  VQ_CHECK_ARG(KR >= 1 && KR <= 16, "%s: 1 <= KR <= 16 (got %d)", who, KR);
  VQ_CHECK_ARG(Kc >= 1 && Kc <= 65535 && (size_t)Kc * (TST + IZ) * 4 <= 190 * 1024, "%s: Kc=%d tables do not fit in shared memory", who, Kc);
  return VQ_OK;
}

// debugging aid: the clock64 phase timers the forward kernel's block 0 / thread 0 recorded at its last launch
VQ_API int vqcpc_idx_debug_read(long long* out16 /*host*/) {
  VQ_CUDA_CALL(cudaMemcpyFromSymbol(out16, g_idx_dbg, sizeof(long long) * 16));
  return VQ_OK;
}

// 1 if the shapes are handled by the index-path kernels below (callers fall back to the per-vector path otherwise)
VQ_API int vqcpc_idxpath_supported(int N, int KR, int zd, int cd, int Kc, int D) {
  return (zd == IZ && cd == IZ && N >= 1 && N <= 15 && KR >= 1 && KR <= 16 && Kc >= 1 && Kc <= 2 * VCH &&
          (size_t)Kc * (TST + IZ) * 4 <= 190 * 1024 && (D == 3 || D == 4)) ? 1 : 0;
}

// number of per-CTA partial tables vqcpc_infonce_idx_bwd writes
VQ_API int vqcpc_infonce_idx_parts(int B, int KR) {
  const int spi = nce_spi(KR);
  return (int)std::max<int64_t>(1, std::min<int64_t>(vq_sm_count(), vq_cdiv(B, spi)));
}

// floats of scratch vqcpc_infonce_idx_bwd needs: per-CTA partial tables (dT | dW) followed by the per-row dc contributions
VQ_API int64_t vqcpc_infonce_idx_scratch(int B, int KR, int Kc) {
  return (int64_t)vqcpc_infonce_idx_parts(B, KR) * ((int64_t)Kc * IZ + (int64_t)IZ * IZ * KR) + (int64_t)B * KR * IZ;
}

VQ_API int vqcpc_infonce_idx_fwd(const float* c, const float* W, const float* T, const int64_t* idx_pos, const int64_t* idx_neg,
                                 int B, int N, int KR, int zd, int cd, int Kc, float* u, float* pos, float* neg, float* lse,
                                 float* nll, int* correct, void* stream) {
  if (int rc = idx_path_check("infonce_idx_fwd", N, KR, zd, cd, Kc)) return rc;
  if (B == 0) return VQ_OK;
  NceFwdP p{c, W, T, idx_pos, idx_neg, B, N, KR, Kc, nce_spi(KR), u, pos, neg, lse, nll, correct};
  const int warps = KR * p.SPI;
  const size_t smem = ((size_t)Kc * IZ + (size_t)warps * 2 * IZ + (size_t)KR * IZ * TST) * 4 + 16;
  VQ_CHECK_ARG(smem <= 227 * 1024, "infonce_idx_fwd: %zu bytes of shared memory needed", smem);
  VQ_CHECK_ARG((((uintptr_t)T | (uintptr_t)c) & 15) == 0, "infonce_idx_fwd: T and c must be 16-byte aligned");
  VQ_CHECK_ARG((int64_t)B * (N + 1) * KR < (1ll << 31) && (int64_t)B * KR * IZ < (1ll << 31), "infonce_idx_fwd: problem too large for 32-bit indexing");
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)vq_sm_count(), vq_cdiv(B, 2 * p.SPI)));
  VQ_CUDA_CALL(cudaFuncSetAttribute(nce_idx_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nce_idx_fwd_kernel<<<grid, warps * 32, smem, vq_stream(stream)>>>(p);
  VQ_CUDA_LAUNCH_CHECK("nce_idx_fwd_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_infonce_idx_bwd(const float* c, const float* W, const float* T, const int64_t* idx_pos, const int64_t* idx_neg,
                                 const float* u, const float* pos, const float* neg, const float* lse, const float* g_dev,
                                 float scale, int B, int N, int KR, int zd, int cd, int Kc, float* d_c, float* coef_pos,
                                 float* coef_neg, float* part, float* d_W, float* d_T, void* stream) {
  if (int rc = idx_path_check("infonce_idx_bwd", N, KR, zd, cd, Kc)) return rc;
  if (B == 0) return VQ_OK;
  NceBwdP p{};
  p.c = c; p.W = W; p.T = T; p.idx_pos = idx_pos; p.idx_neg = idx_neg; p.u = u; p.pos = pos; p.neg = neg; p.lse = lse;
  p.g_dev = g_dev; p.scale = scale; p.B = B; p.N = N; p.KR = KR; p.Kc = Kc; p.SPI = nce_spi(KR);
  p.dc = d_c; p.coef_pos = coef_pos; p.coef_neg = coef_neg; p.part = part;
  const int warps = KR * p.SPI;
  const int grid = vqcpc_infonce_idx_parts(B, KR);
  VQ_CHECK_ARG(grid <= 16 * RP_MAX, "infonce_idx_bwd: %d partial tables exceed the reduction kernel's limit", grid);
  VQ_CHECK_ARG((int64_t)B * (N + 1) * KR < (1ll << 31) && (int64_t)B * KR * IZ < (1ll << 31), "infonce_idx_bwd: problem too large for 32-bit indexing");
  VQ_CHECK_ARG((((uintptr_t)T | (uintptr_t)part) & 15) == 0, "infonce_idx_bwd: T and the scratch must be 16-byte aligned");
  p.dcp = part + (size_t)grid * ((size_t)Kc * IZ + (size_t)IZ * IZ * KR);
  bool dc_in_reduce = false;
  cudaStream_t st = vq_stream(stream);
  // production kernel: dT on the tensor cores (tcgen05, TMEM accumulators); VQCPC_NCE_BWD_ALGO=1 forces the CUDA-core
  // owner-computes kernel (also the fallback when the operand stages do not fit next to the table)
  const int force_simt = g_nce_bwd_algo >= 0 ? g_nce_bwd_algo : env_or("VQCPC_NCE_BWD_ALGO", 0);
  NceBwdTcP q{};
  q.kpad = (Kc + 127) / 128 * 128;
  q.PW = warps;
  q.tmem_cols = 32;
  while (q.tmem_cols < q.kpad / 128 * 32) q.tmem_cols *= 2;
  q.stage_region = (uint32_t)std::max<size_t>((size_t)NST * (2 * (size_t)q.kpad * 32 + 2048), (size_t)warps * IZ * IZ * 4);
  q.stage_region = (q.stage_region + 1023u) & ~1023u;
  const size_t tc_smem = 1024 + (size_t)q.stage_region +
                         ((size_t)Kc * IZ + (size_t)KR * IZ * TST + (size_t)warps * (32 + IZ)) * 4 + (2 * NST + 2) * 8 + 16;
  if (force_simt != 1 && tc_smem <= 227 * 1024 && q.tmem_cols <= 512) {
    q.b = p;
    VQ_CUDA_CALL(cudaFuncSetAttribute(nce_idx_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
    nce_idx_bwd_tc_kernel<<<grid, (warps + 1) * 32, tc_smem, st>>>(q);
    VQ_CUDA_LAUNCH_CHECK("nce_idx_bwd_tc_kernel");
    dc_in_reduce = true;
  } else {
    p.scratch_floats = std::max<size_t>((size_t)Kc * TST, (size_t)warps * IZ * IZ);
    p.scratch_floats = (p.scratch_floats + 3) & ~(size_t)3;
    const size_t smem = (p.scratch_floats + (size_t)Kc * IZ + (size_t)warps * (3 * IZ + 32)) * 4;
    VQ_CHECK_ARG(smem <= 227 * 1024, "infonce_idx_bwd: %zu bytes of shared memory needed", smem);
    VQ_CUDA_CALL(cudaFuncSetAttribute(nce_idx_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nce_idx_bwd_kernel<<<grid, warps * 32, smem, st>>>(p);
    VQ_CUDA_LAUNCH_CHECK("nce_idx_bwd_kernel");
  }
  const int na = Kc * IZ, nb = IZ * IZ * KR;
  const int ncol4 = (na + nb) / 4;
  const int nblk_tab = (ncol4 + 31) / 32, nblk_dc = dc_in_reduce ? (B * 8 + 511) / 512 : 0;
  reduce_partials_kernel<<<nblk_tab + nblk_dc, 512, 0, st>>>(part, grid, (size_t)na + nb, d_T, na, d_W, nb, nblk_tab, p.dcp, d_c, B, KR);
  VQ_CUDA_LAUNCH_CHECK("reduce_partials_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_table_scatter_add(const int64_t* idx, int64_t n, const float* rows, int zd, int Kc, float* out, void* stream) {
  VQ_CHECK_ARG(zd >= 1 && Kc >= 1, "table_scatter_add: bad shape zd=%d Kc=%d", zd, Kc);
  if (n == 0) return VQ_OK;
  const int grid = std::min(vq_sm_count(), Kc);
  const int cpc = (Kc + grid - 1) / grid;
  const size_t smem = (size_t)16 * cpc * zd * 4;
  VQ_CHECK_ARG(smem <= 200 * 1024, "table_scatter_add: %zu bytes of shared memory needed", smem);
  VQ_CUDA_CALL(cudaFuncSetAttribute(table_scatter_add_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  table_scatter_add_kernel<<<(Kc + cpc - 1) / cpc, 512, smem, vq_stream(stream)>>>(idx, n, rows, zd, Kc, cpc, out);
  VQ_CUDA_LAUNCH_CHECK("table_scatter_add_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_vq_idx_bwd(const float* x, int64_t n, int D, const float* E, const float* J, int Kc, int zd, const int64_t* idx,
                            float beta, int squared, const float* g_loss, int nseg, const int64_t* seg_row0, const int64_t* seg_rows,
                            const int* seg_kind, const float* const* seg_a, const float* const* seg_b, const int* seg_per_b, int KR,
                            float* d_x, float* d_E, float* d_T, void* stream) {
  VQ_CHECK_ARG(D == 3 || D == 4, "vq_idx_bwd: D must be 3 or 4 (got %d)", D);
  VQ_CHECK_ARG(zd == IZ, "vq_idx_bwd: zd must be 32 (got %d)", zd);
  VQ_CHECK_ARG(Kc >= 1 && Kc <= 2 * VCH, "vq_idx_bwd: Kc <= %d (got %d)", 2 * VCH, Kc);
  VQ_CHECK_ARG(nseg >= 0 && nseg <= 4, "vq_idx_bwd: at most 4 segments (got %d)", nseg);
  VQ_CHECK_ARG(n < (1ll << 31), "vq_idx_bwd: n must fit in 31 bits");
  if (n == 0) return VQ_OK;
  VqIdxBwdP p{};
  p.x = x; p.E = E; p.J = J; p.gl = g_loss; p.idx = idx; p.n = n; p.Kc = Kc; p.beta = beta; p.squared = squared; p.nseg = nseg;
  p.dx = d_x; p.dE = d_E; p.dT = d_T;
  VQ_CHECK_ARG(d_T == nullptr || ((uintptr_t)d_T & 15) == 0, "vq_idx_bwd: d_T must be 16-byte aligned");
  for (int i = 0; i < nseg; ++i) {
    VQ_CHECK_ARG(seg_kind[i] == 1 || seg_kind[i] == 2, "vq_idx_bwd: segment kind must be 1 (rows) or 2 (coef * u)");
    VQ_CHECK_ARG(seg_row0[i] >= 0 && seg_rows[i] >= 0 && seg_row0[i] + seg_rows[i] <= n && seg_rows[i] < (1 << 30),
                 "vq_idx_bwd: segment %d out of range", i);
    VQ_CHECK_ARG(seg_kind[i] == 1 || (seg_per_b[i] >= 1 && KR >= 1 && seg_per_b[i] % KR == 0), "vq_idx_bwd: segment %d: bad per_b / KR", i);
    VQ_CHECK_ARG((((uintptr_t)seg_a[i] | (uintptr_t)(seg_kind[i] == 2 ? seg_b[i] : nullptr)) & 15) == 0, "vq_idx_bwd: segment %d: sources must be 16-byte aligned", i);
    const int pb = seg_kind[i] == 2 ? seg_per_b[i] : 1;
    VQ_CHECK_ARG(pb < 1024 && KR < 1024, "vq_idx_bwd: per_b / KR too large");
    p.seg[i] = IdxSeg{seg_row0[i], seg_rows[i], seg_kind[i], seg_a[i], seg_kind[i] == 2 ? seg_b[i] : nullptr, pb, KR,
                      ((1ull << 40) + pb - 1) / pb, ((1ull << 40) + KR - 1) / KR};
    if (seg_kind[i] == 1) p.kind1_mask |= 1u << i;
  }
  cudaStream_t st = vq_stream(stream);
  const int algo = g_vq_bwd_algo >= 0 ? g_vq_bwd_algo : env_or("VQCPC_VQ_BWD_ALGO", 0);          // 1: the single-kernel version
  static const int lpv = getenv("VQCPC_VQ_BWD_LPV") ? atoi(getenv("VQCPC_VQ_BWD_LPV")) : 2;       // lanes per vector of the dz kernel
  if (algo != 1) {
    // production: streaming dz kernel + counting-sort dE kernel
#define VQ_IB2(DSV)                                                                                                           \
  do {                                                                                                                        \
    const size_t js_bytes = (size_t)Kc * (IZ * DSV + 4) * 4, tail = (size_t)Kc * 16 + 8 + 4 * sizeof(IdxSeg) + 8;                            \
    const bool js = (((uintptr_t)J & 15) == 0) && js_bytes + tail <= 227 * 1024;                                              \
    p.j_smem = js ? 1 : 0;                                                                                                    \
    const size_t smem = (js ? js_bytes : 0) + tail;                                                                           \
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(n, 128), vq_sm_count()));                            \
    if (js && lpv == 2) {                                                                                                     \
      VQ_CUDA_CALL(cudaFuncSetAttribute(vq_idx_dz_kernel<DSV, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      vq_idx_dz_kernel<DSV, true, 2><<<grid, 512, smem, st>>>(p);                                                             \
    } else if (js) {                                                                                                          \
      VQ_CUDA_CALL(cudaFuncSetAttribute(vq_idx_dz_kernel<DSV, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      vq_idx_dz_kernel<DSV, true, 4><<<grid, 512, smem, st>>>(p);                                                             \
    } else {                                                                                                                  \
      VQ_CUDA_CALL(cudaFuncSetAttribute(vq_idx_dz_kernel<DSV, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      vq_idx_dz_kernel<DSV, false, 4><<<grid, 512, smem, st>>>(p);                                                            \
    }                                                                                                                         \
    VQ_CUDA_LAUNCH_CHECK("vq_idx_dz_kernel");                                                                                 \
    if (d_E) {                                                                                                                \
      const size_t smem2 = (size_t)Kc * 16 + (size_t)DE_CH * DSV * 4 + ((size_t)3 * Kc + 32) * 4 + (size_t)DE_CH * 4 + 16;    \
      const int grid2 = (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(n, DE_CH), 2 * (int64_t)vq_sm_count()));          \
      VQ_CUDA_CALL(cudaFuncSetAttribute(vq_idx_de_kernel<DSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));     \
      vq_idx_de_kernel<DSV><<<grid2, DE_T, smem2, st>>>(p);                                                                   \
      VQ_CUDA_LAUNCH_CHECK("vq_idx_de_kernel");                                                                               \
    }                                                                                                                         \
  } while (0)
    if (D == 3) VQ_IB2(3); else VQ_IB2(4);
#undef VQ_IB2
    return VQ_OK;
  }
  const int64_t nchunk = vq_cdiv(n, VCH);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nchunk, vq_sm_count()));
#define VQ_IB(DSV)                                                                                                            \
  do {                                                                                                                        \
    const bool js = (((uintptr_t)J & 15) == 0) && vq_idx_bwd_smem<DSV>(Kc, true) <= 227 * 1024;                               \
    p.j_smem = js ? 1 : 0;                                                                                                    \
    const size_t smem = vq_idx_bwd_smem<DSV>(Kc, js);                                                                         \
    VQ_CUDA_CALL(cudaFuncSetAttribute(vq_idx_bwd_kernel<DSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    vq_idx_bwd_kernel<DSV><<<grid, VCH, smem, st>>>(p);                                                                       \
  } while (0)
  if (D == 3) VQ_IB(3); else VQ_IB(4);
#undef VQ_IB
  VQ_CUDA_LAUNCH_CHECK("vq_idx_bwd_kernel");
  return VQ_OK;
}
