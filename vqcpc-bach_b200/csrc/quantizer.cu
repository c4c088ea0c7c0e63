// Fused vector-quantisation bottleneck (ProductVectorQuantizer, VQCPCB/quantizer/vector_quantizer.py:85-159).
//
//   search : idx[n, cb] = argmin_k sum_d (x[n, cb*Ds + d] - E_cb[k, d])^2          (:104-116)
//            direct form in fp32, no FMA contraction, and the SAME summation order torch's CPU sum uses for a
//            contiguous inner dimension (8 interleaved partial sums combined in order; 32 when Ds % 32 == 0), so the
//            distances - and therefore the argmin, ties included (lowest index wins) - are bit-identical to the
//            reference for Ds <= 8 and Ds in {16, 24, 32, 64, 128} (checked against torch 2.11 CPU).
//   gather : q = E[idx]; loss = q_latent + beta * e_latent (:72-83); z_sg = x + (q - x) (:148); usage counts
//   bwd    : dL/dx = g_sg + g_loss * d(loss)/dx ; dL/dE[idx] += g_loss * d(loss)/dq   (straight-through :148)
//
// Algorithmic traffic per vector (D=3, one codebook): 12 B in, 12 + 8 + 4 B out -> HBM-bound; the codebook tile is
// staged once per CTA in shared memory and read as warp-wide broadcasts.
#include "common.cuh"
#include "vq_dist.cuh"
#include "vq_tc.h"
#include <cstdlib>

namespace {

constexpr int MAX_CB = 8;
struct Books {
  const float* e[MAX_CB];
  float* g[MAX_CB];
};

__device__ __forceinline__ float dist_generic(const float* __restrict__ x, const float* __restrict__ e, int DS) {
  const int W = (DS % 32 == 0) ? 32 : 8;
  float acc[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) acc[l] = 0.f;
  for (int d0 = 0; d0 < DS; d0 += W) {
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      if (l < W && d0 + l < DS) {
        float df = __fsub_rn(x[d0 + l], e[d0 + l]);
        float sq = __fmul_rn(df, df);
        acc[l] = (d0 == 0) ? sq : __fadd_rn(acc[l], sq);
      }
    }
  }
  if (W == 32) {
#pragma unroll
    for (int l = 0; l < 8; ++l) acc[l] = __fadd_rn(__fadd_rn(__fadd_rn(acc[l], acc[8 + l]), acc[16 + l]), acc[24 + l]);
  }
  float s = acc[0];
#pragma unroll
  for (int l = 1; l < 8; ++l)
    if (l < DS) s = __fadd_rn(s, acc[l]);
  return s;
}

// blockIdx.y = sub-codebook; one thread per vector; codebook in dynamic smem
template <int DS>
__global__ void __launch_bounds__(256)
vq_search_kernel(const float* __restrict__ x, Books books, int Kc, int D, int DSrt, int ncb, int64_t n, int64_t* __restrict__ idx) {
  extern __shared__ __align__(16) float cb[];
  const int ds = DS > 0 ? DS : DSrt;
  const int c = blockIdx.y;
  const float* __restrict__ E = books.e[c];
  for (int i = threadIdx.x; i < Kc * ds; i += blockDim.x) cb[i] = E[i];
  __syncthreads();
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const float* xp = x + v * D + c * ds;
    float best = INFINITY;
    int bi = 0;
    if (DS > 0) {
      float xr[DS > 0 ? DS : 1];
#pragma unroll
      for (int d = 0; d < DS; ++d) xr[d] = xp[d];
      for (int k = 0; k < Kc; ++k) {
        float dd = dist_fixed<(DS > 0 ? DS : 1)>(xr, cb + k * DS);
        if (dd < best || (k == 0)) { best = dd; bi = k; }     // strict '<' keeps the lowest index on ties; NaN never wins
      }
    } else {
      for (int k = 0; k < Kc; ++k) {
        float dd = dist_generic(xp, cb + k * ds, ds);
        if (dd < best || (k == 0)) { best = dd; bi = k; }
      }
    }
    idx[v * ncb + c] = bi;
  }
}

// one thread per vector: gather, loss, straight-through output, usage histogram
__global__ void __launch_bounds__(256)
vq_gather_kernel(const float* __restrict__ x, Books books, int Kc, int D, int ds, int ncb, int64_t n,
                 const int64_t* __restrict__ idx, float beta, int squared, float* __restrict__ zsg,
                 float* __restrict__ loss, int* __restrict__ counts) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) acc[l] = 0.f;
    for (int c = 0; c < ncb; ++c) {
      int k = (int)idx[v * ncb + c];
      k = k < 0 ? 0 : (k >= Kc ? Kc - 1 : k);
      if (counts) atomicAdd(counts + c * Kc + k, 1);
      const float* __restrict__ e = books.e[c] + (size_t)k * ds;
      for (int d = 0; d < ds; ++d) {
        int dd = c * ds + d;
        float xv = x[v * D + dd];
        float df = __fsub_rn(e[d], xv);                       // q - x
        zsg[v * D + dd] = __fadd_rn(xv, df);                  // x + (q - x).detach()
        if (!squared) df = __fadd_rn(df, 1e-5f);
        float sq = __fmul_rn(df, df);
        acc[dd & 7] = __fadd_rn(acc[dd & 7], sq);
      }
    }
    float s = acc[0];
#pragma unroll
    for (int l = 1; l < 8; ++l) s = __fadd_rn(s, acc[l]);
    float ql = squared ? s : sqrtf(s);
    loss[v] = __fadd_rn(ql, __fmul_rn(beta, ql));             // q_latent + beta * e_latent (same value, two graphs)
  }
}

__global__ void __launch_bounds__(256)
vq_bwd_kernel(const float* __restrict__ x, Books books, int Kc, int D, int ds, int ncb, int64_t n,
              const int64_t* __restrict__ idx, float beta, int squared, const float* __restrict__ g_zsg,
              const float* __restrict__ g_loss, float* __restrict__ dx) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    float gl = g_loss ? g_loss[v] : 0.f;
    float inv_norm = 0.f;
    if (!squared) {
      float s = 0.f;
      for (int c = 0; c < ncb; ++c) {
        int k = (int)idx[v * ncb + c];
        const float* __restrict__ e = books.e[c] + (size_t)k * ds;
        for (int d = 0; d < ds; ++d) {
          float df = e[d] - x[v * D + c * ds + d] + 1e-5f;
          s += df * df;
        }
      }
      inv_norm = 1.f / sqrtf(s);
    }
    for (int c = 0; c < ncb; ++c) {
      int k = (int)idx[v * ncb + c];
      const float* __restrict__ e = books.e[c] + (size_t)k * ds;
      float* ge = books.g[c] ? books.g[c] + (size_t)k * ds : nullptr;
      for (int d = 0; d < ds; ++d) {
        int dd = c * ds + d;
        float df = e[d] - x[v * D + dd];                       // q - x
        float dq = squared ? 2.f * df : (df + 1e-5f) * inv_norm;   // d q_latent / d q ; d e_latent / d x = -dq
        float gs = g_zsg ? g_zsg[v * D + dd] : 0.f;
        dx[v * D + dd] = gs - gl * beta * dq;
        if (ge && gl != 0.f) atomicAdd(ge + d, gl * dq);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Fused bottleneck for the common single-codebook case with a dropout-free upscaler:
//   one pass over z: exact search -> idx, loss, usage histogram, and y = T[idx], where T = upscaler(codebook) is a
//   (Kc, zd) table computed once per call on the Kc code vectors (the upscaled output of every vector assigned to code
//   k is the same row).  Algorithmic traffic per vector: 4*D in, 8 (idx) + 4 (loss) + 4*zd (y) out = 152 B for D=3,
//   zd=32 (SURVEY §8d stage A).  The direct-form exact search costs ~12 fp32 instructions per (vector, code) and is the
//   binding limit for Kc = 512 (fp32-ALU bound, not HBM bound).
template <int DS>
__global__ void __launch_bounds__(256)
vq_fused_fwd_kernel(const float* __restrict__ x, const float* __restrict__ E, const float* __restrict__ T, int Kc, int zd,
                    int64_t n, float beta, int squared, int64_t* __restrict__ idx, float* __restrict__ loss,
                    float* __restrict__ zsg, float* __restrict__ y, int64_t y_row0, int64_t y_rows, int* __restrict__ counts) {
  extern __shared__ __align__(16) float sm[];
  constexpr int CS = (DS == 3) ? 4 : DS;          // padded code stride (one LDS.128 per code for D = 3)
  float* cb = sm;                                  // [Kc][CS]
  int* hist = reinterpret_cast<int*>(cb + Kc * CS);   // [Kc]
  // (the (Kc, zd) table T stays in global memory: 64 KB, L1/L2 resident, read as whole 128-byte rows)
  for (int i = threadIdx.x; i < Kc * CS; i += blockDim.x) {
    int k = i / CS, dd = i % CS;
    cb[i] = dd < DS ? E[k * DS + dd] : 0.f;
  }
  for (int i = threadIdx.x; i < Kc; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // whole warps stay in the loop together (the y write-out below is warp-cooperative)
  for (int64_t v0 = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); v0 < n; v0 += stride) {
    const int64_t v = v0 + lane;
    const bool ok = v < n;
    float xr[DS];
#pragma unroll
    for (int dd = 0; dd < DS; ++dd) xr[dd] = ok ? x[v * DS + dd] : 0.f;
    float best = dist_fixed<DS>(xr, cb);        // code 0 seeds the running minimum; strict '<' keeps the lowest index
    int bi = 0;
#pragma unroll 4
    for (int k = 1; k < Kc; ++k) {
      float dd = dist_fixed<DS>(xr, cb + k * CS);
      if (dd < best) { best = dd; bi = k; }
    }
    if (ok) {
      // loss / straight-through output exactly as vq_gather_kernel
      float acc[8];
#pragma unroll
      for (int l = 0; l < 8; ++l) acc[l] = 0.f;
#pragma unroll
      for (int dd = 0; dd < DS; ++dd) {
        float df = __fsub_rn(cb[bi * CS + dd], xr[dd]);
        if (zsg) zsg[v * DS + dd] = __fadd_rn(xr[dd], df);
        if (!squared) df = __fadd_rn(df, 1e-5f);
        acc[dd & 7] = __fadd_rn(acc[dd & 7], __fmul_rn(df, df));
      }
      float s = acc[0];
#pragma unroll
      for (int l = 1; l < 8; ++l) s = __fadd_rn(s, acc[l]);
      float ql = squared ? s : sqrtf(s);
      loss[v] = __fadd_rn(ql, __fmul_rn(beta, ql));
      idx[v] = bi;
      if (counts) atomicAdd(&hist[bi], 1);
    }
    if (y && v0 < y_row0 + y_rows && v0 + 32 > y_row0) {         // y covers the rows [y_row0, y_row0 + y_rows) only
      for (int sft = 0; sft < 32; ++sft) {
        const int b = __shfl_sync(0xffffffffu, bi, sft);
        const int64_t r = v0 + sft;
        if (r < n && r >= y_row0 && r < y_row0 + y_rows)
          for (int j = lane; j < zd; j += 32) y[(r - y_row0) * zd + j] = __ldg(T + (size_t)b * zd + j);
      }
    }
  }
  __syncthreads();
  if (counts)
    for (int i = threadIdx.x; i < Kc; i += blockDim.x)
      if (hist[i]) atomicAdd(counts + i, hist[i]);
}

// backward of the fused path.  J (Kc, zd, D) = d upscaler / d input evaluated at each code vector.
//   dx[v]      = J[idx]^T dy[v] + g_loss[v] * d loss / d x           (straight-through + commitment)
//   dT[idx]   += dy[v]   (CTA-private smem table, one atomic flush)  -> upscaler weight gradients via the Kc-row MLP
//   dE[idx]   += g_loss[v] * d loss / d q
template <int DS>
__global__ void __launch_bounds__(256)
vq_fused_bwd_kernel(const float* __restrict__ x, const float* __restrict__ E, const float* __restrict__ J, int Kc, int zd,
                    int64_t n, const int64_t* __restrict__ idx, float beta, int squared, const float* __restrict__ dy,
                    const float* __restrict__ g_loss, float* __restrict__ dx, float* __restrict__ dT, float* __restrict__ dE) {
  extern __shared__ __align__(16) float sm[];
  float* dTs = sm;                                 // [Kc][zd]
  float* dEs = dTs + Kc * zd;                      // [Kc][DS]
  for (int i = threadIdx.x; i < Kc * (zd + DS); i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v0 = blockIdx.x * (int64_t)blockDim.x + (threadIdx.x & ~31); v0 < n; v0 += stride) {
    const int64_t v = v0 + lane;
    const bool ok = v < n;
    const int k = ok ? (int)idx[v] : 0;
    float gx[DS];
#pragma unroll
    for (int dd = 0; dd < DS; ++dd) gx[dd] = 0.f;
    // warp-cooperative: for each vector of the warp, lanes span the zd features
    for (int sft = 0; sft < 32; ++sft) {
      const int kb = __shfl_sync(0xffffffffu, k, sft);
      const bool vok = (v0 + sft) < n;
      float part[DS];
#pragma unroll
      for (int dd = 0; dd < DS; ++dd) part[dd] = 0.f;
      if (vok) {
        for (int j = lane; j < zd; j += 32) {
          const float g = dy[(v0 + sft) * zd + j];
          atomicAdd(&dTs[kb * zd + j], g);
#pragma unroll
          for (int dd = 0; dd < DS; ++dd) part[dd] = fmaf(g, __ldg(J + ((size_t)kb * zd + j) * DS + dd), part[dd]);
        }
      }
#pragma unroll
      for (int dd = 0; dd < DS; ++dd) {
        float t = warp_sum(part[dd]);
        if (lane == sft) gx[dd] = t;
      }
    }
    if (ok) {
      const float gl = g_loss ? g_loss[v] : 0.f;
      float inv_norm = 0.f;
      if (!squared) {
        float s = 0.f;
#pragma unroll
        for (int dd = 0; dd < DS; ++dd) {
          float df = E[k * DS + dd] - x[v * DS + dd] + 1e-5f;
          s += df * df;
        }
        inv_norm = 1.f / sqrtf(s);
      }
#pragma unroll
      for (int dd = 0; dd < DS; ++dd) {
        const float df = E[k * DS + dd] - x[v * DS + dd];
        const float dq = squared ? 2.f * df : (df + 1e-5f) * inv_norm;
        dx[v * DS + dd] = gx[dd] - gl * beta * dq;
        if (gl != 0.f) atomicAdd(&dEs[k * DS + dd], gl * dq);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Kc * zd; i += blockDim.x)
    if (dTs[i] != 0.f) atomicAdd(dT + i, dTs[i]);
  if (dE)
    for (int i = threadIdx.x; i < Kc * DS; i += blockDim.x)
      if (dEs[i] != 0.f) atomicAdd(dE + i, dEs[i]);
}

// Streaming backward of the fused path for zd == 32 (the upscaler width every configuration uses).
// dT[idx_v] += dy[v] is a scatter-add of 128-byte rows into a Kc x 32 table.  Floating-point atomics on shared memory are
// CAS loops on sm_100 (ATOMS.CAST.SPIN: measured 325 us for 557 056 rows), vector reductions straight to the L2-resident
// table saturate the L2 atomic units (170 us), and warp-cooperative segmented sums drown in shuffles (157 us).  So the
// scatter is turned into a gather: every CTA takes a contiguous block of at most VB_SUB vectors, counting-sorts their
// indices by code in shared memory (native 32-bit ATOMS.ADD), then
// four threads (adjacent lanes) per task - a code's rows, at most 16 at a time, prefetched four deep - walk those rows, each with a quarter of the features and of the code's
// 32 x DS upscaler Jacobian in registers:
//   dT_smem[k] += sum rows;   dx[v] = J[k]^T dy[v] + g_loss[v] dloss/dx (two shuffle steps inside the quad);
//   dE_smem[k] += sum g_loss dloss/dq.
// Table words have one writer, except for the codes popular enough to be split over several tasks (shared CAS atomics).  The per-CTA tables are summed across the
// 4 CTAs of a thread-block cluster through distributed shared memory before the global atomics (4x fewer red.global).
constexpr int VB_SUB = 4096;          // vectors per sorted sub-block (local indices fit in 16 bits)
constexpr int VB_THREADS = 512;
constexpr int VB_ROWS = 16;           // rows per task: popular codes are split so that no quad walks a long segment alone
constexpr int VB_TS = 33;             // padded row of the shared dT table (conflict-free per-thread row writes)

template <int DS>
__global__ void __launch_bounds__(VB_THREADS, 1)
vq_fused_bwd_stream_kernel(const float* __restrict__ x, const float* __restrict__ E, const float* __restrict__ J, int Kc,
                           int64_t n, const int64_t* __restrict__ idx, float beta, int squared, const float* __restrict__ dy,
                           const float* __restrict__ g_loss, float* __restrict__ dx, float* __restrict__ dT,
                           float* __restrict__ dE) {
  constexpr int ZD = 32;
  extern __shared__ __align__(16) float sm[];
  float* dTs = sm;                                 // [Kc][33]
  float* dEs = dTs + Kc * VB_TS;                   // [Kc][DS]
  int* hist = reinterpret_cast<int*>(dEs + Kc * DS);      // [Kc + 1] counts, then exclusive offsets
  int* cursor = hist + Kc + 1;                            // [Kc]
  int* toff = cursor + Kc;                                // [Kc + 1] first task of each code
  int* task = toff + Kc + 1;                              // [Kc + VB_SUB / VB_ROWS] (code | chunk << 16)
  unsigned short* perm = reinterpret_cast<unsigned short*>(task + Kc + VB_SUB / VB_ROWS);     // [VB_SUB] local vector index, sorted by code
  const int tsz = Kc * (VB_TS + DS);
  for (int i = threadIdx.x; i < tsz; i += VB_THREADS) sm[i] = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // contiguous range of this CTA, walked in sub-blocks
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo_v = min(n, per * blockIdx.x), hi_v = min(n, lo_v + per);
  for (int64_t base = lo_v; base < hi_v; base += VB_SUB) {
    const int cnt = (int)min((int64_t)VB_SUB, hi_v - base);
    __syncthreads();                               // previous sub-block fully consumed (perm / hist reuse)
    for (int i = threadIdx.x; i <= Kc; i += VB_THREADS) hist[i] = 0;
    __syncthreads();
    constexpr int PER_T = VB_SUB / VB_THREADS;
    int kreg[PER_T];
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
      const int i = threadIdx.x + j * VB_THREADS;
      int k = i < cnt ? (int)idx[base + i] : 0;
      kreg[j] = k < 0 ? 0 : (k >= Kc ? Kc - 1 : k);
    }
#pragma unroll
    for (int j = 0; j < PER_T; ++j)
      if (threadIdx.x + j * VB_THREADS < cnt) atomicAdd(&hist[kreg[j]], 1);
    __syncthreads();
    if (warp == 0) {                               // exclusive scans: rows -> hist (offsets), cursor; tasks -> toff
      int carry = 0, tcarry = 0;
      for (int k0 = 0; k0 < Kc; k0 += 32) {
        const int c = (k0 + lane < Kc) ? hist[k0 + lane] : 0;
        const int nt = (c + VB_ROWS - 1) / VB_ROWS;
        int incl = c, tincl = nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          const int tt = __shfl_up_sync(0xffffffffu, tincl, o);
          if (lane >= o) { incl += t; tincl += tt; }
        }
        if (k0 + lane < Kc) {
          hist[k0 + lane] = carry + incl - c;
          cursor[k0 + lane] = carry + incl - c;
          toff[k0 + lane] = tcarry + tincl - nt;
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
        tcarry += __shfl_sync(0xffffffffu, tincl, 31);
      }
      if (lane == 0) { hist[Kc] = carry; toff[Kc] = tcarry; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
      const int i = threadIdx.x + j * VB_THREADS;
      if (i < cnt) perm[atomicAdd(&cursor[kreg[j]], 1)] = (unsigned short)i;
    }
    for (int k = threadIdx.x; k < Kc; k += VB_THREADS)
      for (int t = toff[k], c = 0; t < toff[k + 1]; ++t, ++c) task[t] = k | (c << 16);
    __syncthreads();
    // ---- four threads per code (adjacent lanes; a quarter of the features and of the Jacobian each -> one 128-byte line
    // per row and quad), rows prefetched four deep.  Lane q < DS of the quad also owns dimension q of x / dx / dE.
    const int ntask = toff[Kc];
    for (int t = threadIdx.x >> 2; t < ntask; t += VB_THREADS / 4) {
      const int k = task[t] & 0xFFFF, qd = threadIdx.x & 3;
      const int s0 = hist[k] + (task[t] >> 16) * VB_ROWS, s1 = min(hist[k + 1], s0 + VB_ROWS);
      const bool shared_code = hist[k + 1] - hist[k] > VB_ROWS;        // several tasks add into the same table row
      const unsigned qmask = 0xFu << (lane & ~3);
      const float ek = (qd < DS) ? __ldg(E + (size_t)k * DS + qd) : 0.f;
      float jq[8 * DS];                             // J[k][8 qd .. 8 qd + 7][0 .. DS-1]
      {
        const float4* jr = reinterpret_cast<const float4*>(J + ((size_t)k * ZD + qd * 8) * DS);     // 8 * DS floats, 16-byte aligned
#pragma unroll
        for (int q = 0; q < 2 * DS; ++q) {
          const float4 t4 = __ldg(jr + q);
          jq[4 * q] = t4.x; jq[4 * q + 1] = t4.y; jq[4 * q + 2] = t4.z; jq[4 * q + 3] = t4.w;
        }
      }
      float acc[8], de = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      constexpr int PF = 4;
      float4 buf[PF][2];
      float bx[PF], bg[PF];
      int bv[PF];
      auto fetch = [&](int slot, int i) {
        if (i < s1) {
          const int pl = perm[i];
          const int64_t v = base + pl;
          const float4* r = reinterpret_cast<const float4*>(dy + v * ZD) + qd * 2;
          buf[slot][0] = r[0];
          buf[slot][1] = r[1];
          bv[slot] = pl;
          if (qd < DS) {
            bx[slot] = x[v * DS + qd];
            bg[slot] = g_loss ? g_loss[v] : 0.f;
          }
        }
      };
#pragma unroll
      for (int u = 0; u < PF; ++u) fetch(u, s0 + u);
      for (int i0 = s0; i0 < s1; i0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          if (i0 + u < s1) {
            const float4 a = buf[u][0], c = buf[u][1];
            const float row[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
            float gx[DS];
#pragma unroll
            for (int d = 0; d < DS; ++d) gx[d] = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              acc[j] += row[j];
#pragma unroll
              for (int d = 0; d < DS; ++d) gx[d] = fmaf(row[j], jq[j * DS + d], gx[d]);
            }
            const float df = (qd < DS) ? ek - bx[u] : 0.f;
            float s2 = (qd < DS) ? (df + 1e-5f) * (df + 1e-5f) : 0.f;
#pragma unroll
            for (int o = 1; o < 4; o <<= 1) {
#pragma unroll
              for (int d = 0; d < DS; ++d) gx[d] += __shfl_xor_sync(qmask, gx[d], o);
              if (!squared) s2 += __shfl_xor_sync(qmask, s2, o);
            }
            if (qd < DS) {
              float gsel = gx[0];
#pragma unroll
              for (int d = 1; d < DS; ++d) gsel = (qd == d) ? gx[d] : gsel;
              const float dq = squared ? 2.f * df : (df + 1e-5f) / sqrtf(s2);
              dx[(base + bv[u]) * DS + qd] = gsel - bg[u] * beta * dq;
              de += bg[u] * dq;
            }
          }
          fetch(u, i0 + PF + u);
        }
      }
      if (shared_code) {                            // rare: (CAS) atomics only where a popular code was split
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&dTs[k * VB_TS + qd * 8 + j], acc[j]);
        if (qd < DS) atomicAdd(&dEs[k * DS + qd], de);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dTs[k * VB_TS + qd * 8 + j] += acc[j];
        if (qd < DS) dEs[k * DS + qd] += de;
      }
    }
  }
  __syncthreads();
  // ---- cluster reduction of the tables, then one global atomic per word and cluster
  unsigned crank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int nwords = Kc * (ZD + DS);                // logical words: dT (Kc x 32) then dE (Kc x DS)
  const int slice = (nwords + (int)csize - 1) / (int)csize;
  const int lo = (int)crank * slice, hi = min(nwords, lo + slice);
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
  for (int i = lo + threadIdx.x; i < hi; i += VB_THREADS) {
    const int off = i < Kc * ZD ? (i / ZD) * VB_TS + (i % ZD) : Kc * VB_TS + (i - Kc * ZD);
    float acc = 0.f;
    for (unsigned r = 0; r < csize; ++r) {
      uint32_t ra;
      float t;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sbase + 4u * (uint32_t)off), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t) : "r"(ra));
      acc += t;
    }
    if (acc != 0.f) {
      if (i < Kc * ZD) atomicAdd(dT + i, acc);
      else if (dE) atomicAdd(dE + (i - Kc * ZD), acc);
    }
  }
  // nobody leaves while a peer may still read its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int fill_books(Books& b, int ncb, const float* const* e, float* const* g) {
  if (ncb < 1 || ncb > MAX_CB) return -1;
  for (int i = 0; i < ncb; ++i) {
    b.e[i] = e[i];
    b.g[i] = g ? g[i] : nullptr;
  }
  return 0;
}

inline int vec_grid(int64_t n) {
  int64_t g = vq_cdiv(n, 256);
  int64_t cap = (int64_t)vq_sm_count() * 8;
  return (int)std::max<int64_t>(1, std::min(g, cap));
}

}  // namespace

// Nearest-code search.  x: (n, D) fp32; books[cb]: (Kc, D/ncb) fp32; idx: (n, ncb) int64 (the reference's dtype).
VQ_API int vqcpc_vq_search(const float* x, int64_t n, int D, const float* const* books, int ncb, int Kc, int64_t* idx,
                           void* stream) {
  Books b;
  VQ_CHECK_ARG(fill_books(b, ncb, books, nullptr) == 0, "vq_search: 1..%d sub-codebooks supported", MAX_CB);
  VQ_CHECK_ARG(D % ncb == 0 && Kc >= 1, "vq_search: D=%d not divisible by num_codebooks=%d", D, ncb);
  if (n == 0) return VQ_OK;
  int ds = D / ncb;
  size_t smem = (size_t)Kc * ds * sizeof(float);
  VQ_CHECK_ARG(smem <= 200 * 1024, "vq_search: codebook tile (%zu B) does not fit in shared memory", smem);
  dim3 grid(vec_grid(n), ncb);
  cudaStream_t st = vq_stream(stream);
#define VQ_LAUNCH(DSV)                                                                                         \
  do {                                                                                                         \
    if (smem > 48 * 1024)                                                                                      \
      VQ_CUDA_CALL(cudaFuncSetAttribute(vq_search_kernel<DSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    vq_search_kernel<DSV><<<grid, 256, smem, st>>>(x, b, Kc, D, ds, ncb, n, idx);                              \
  } while (0)
  switch (ds) {
    case 1: VQ_LAUNCH(1); break;
    case 2: VQ_LAUNCH(2); break;
    case 3: VQ_LAUNCH(3); break;
    case 4: VQ_LAUNCH(4); break;
    case 8: VQ_LAUNCH(8); break;
    case 16: VQ_LAUNCH(16); break;
    case 32: VQ_LAUNCH(32); break;
    default: VQ_LAUNCH(0); break;
  }
#undef VQ_LAUNCH
  VQ_CUDA_LAUNCH_CHECK("vq_search_kernel");
  return VQ_OK;
}

// Gather + loss + straight-through output (+ optional usage histogram counts[ncb*Kc], accumulated).
VQ_API int vqcpc_vq_gather(const float* x, int64_t n, int D, const float* const* books, int ncb, int Kc, const int64_t* idx,
                           float beta, int squared, float* z_sg, float* loss, int* counts, void* stream) {
  Books b;
  VQ_CHECK_ARG(fill_books(b, ncb, books, nullptr) == 0, "vq_gather: 1..%d sub-codebooks supported", MAX_CB);
  VQ_CHECK_ARG(D % ncb == 0, "vq_gather: D=%d not divisible by num_codebooks=%d", D, ncb);
  if (n == 0) return VQ_OK;
  vq_gather_kernel<<<vec_grid(n), 256, 0, vq_stream(stream)>>>(x, b, Kc, D, D / ncb, ncb, n, idx, beta, squared, z_sg, loss, counts);
  VQ_CUDA_LAUNCH_CHECK("vq_gather_kernel");
  return VQ_OK;
}

// Backward: d_x overwritten; d_books[cb] ACCUMULATED (may be NULL).  g_zsg / g_loss may be NULL (treated as zero).
VQ_API int vqcpc_vq_bwd(const float* x, int64_t n, int D, const float* const* books, float* const* d_books, int ncb, int Kc,
                        const int64_t* idx, float beta, int squared, const float* g_zsg, const float* g_loss, float* d_x,
                        void* stream) {
  Books b;
  VQ_CHECK_ARG(fill_books(b, ncb, books, d_books) == 0, "vq_bwd: 1..%d sub-codebooks supported", MAX_CB);
  if (n == 0) return VQ_OK;
  vq_bwd_kernel<<<vec_grid(n), 256, 0, vq_stream(stream)>>>(x, b, Kc, D, D / ncb, ncb, n, idx, beta, squared, g_zsg, g_loss, d_x);
  VQ_CUDA_LAUNCH_CHECK("vq_bwd_kernel");
  return VQ_OK;
}

// Fused single-codebook bottleneck with the upscaler folded into a table: see vq_fused_fwd_kernel / vq_tc.cu.
//   x (n, D); E (Kc, D); T (Kc, zd) = upscaler(E) or NULL (then y is not written); idx (n) int64; loss (n);
//   z_sg (n, D) optional; y (n, zd) optional; counts (Kc) int32 ACCUMULATED, optional.
//   algo: 0 = automatic, 1 = direct-form scan on the CUDA cores, 2 = tcgen05 score + exact verification (vq_tc.cu).
//   Both algorithms return bit-identical indices.  dbg_scores / stats: test hooks of algo 2 (NULL in production).
static int vq_fused_fwd_direct(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta,
                               int squared, int64_t* idx, float* loss, float* z_sg, float* y, int64_t y_row0, int64_t y_rows,
                               int* counts, cudaStream_t st) {
  const int cs = D == 3 ? 4 : D;
  size_t smem = ((size_t)Kc * cs + Kc) * sizeof(float);
  VQ_CHECK_ARG(smem <= 200 * 1024, "vq_fused_fwd: codebook (%zu B) does not fit in shared memory", smem);
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(n, 256), (int64_t)vq_sm_count() * per_sm));
#define VQ_FL(DSV)                                                                                                          \
  do {                                                                                                                      \
    VQ_CUDA_CALL(cudaFuncSetAttribute(vq_fused_fwd_kernel<DSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    vq_fused_fwd_kernel<DSV><<<grid, 256, smem, st>>>(x, E, T, Kc, zd, n, beta, squared, idx, loss, z_sg, y, y_row0, y_rows, counts);        \
  } while (0)
  switch (D) {
    case 3: VQ_FL(3); break;
    case 4: VQ_FL(4); break;
    case 8: VQ_FL(8); break;
    case 16: VQ_FL(16); break;
    default: VQ_FL(32); break;
  }
#undef VQ_FL
  VQ_CUDA_LAUNCH_CHECK("vq_fused_fwd_kernel");
  return VQ_OK;
}

static int vq_fused_fwd_any(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta,
                            int squared, int64_t* idx, float* loss, float* z_sg, float* y, int64_t y_row0, int64_t y_rows,
                            int* counts, int algo, float* dbg_scores, int* stats, void* stream) {
  VQ_CHECK_ARG(D == 3 || D == 4 || D == 8 || D == 16 || D == 32, "vq_fused_fwd: D must be 3, 4, 8, 16 or 32 (got %d)", D);
  VQ_CHECK_ARG((T == nullptr) == (y == nullptr), "vq_fused_fwd: T and y go together");
  VQ_CHECK_ARG(algo >= 0 && algo <= 2, "vq_fused_fwd: algo must be 0 (auto), 1 (direct) or 2 (tensor) (got %d)", algo);
  VQ_CHECK_ARG(y_row0 >= 0 && y_rows >= 0 && y_row0 + y_rows <= n, "vq_fused_fwd: y row range [%lld, +%lld) outside [0, %lld)",
               (long long)y_row0, (long long)y_rows, (long long)n);
  if (n == 0) return VQ_OK;
  cudaStream_t st = vq_stream(stream);
  if (algo == 0) {
    // the tensor-core search wins from about a thousand vectors on (measured: 13 us vs 25 us at n = 4352, Kc = 512 - the
    // direct scan is one long dependent chain per thread); tiny problems keep the simpler kernel
    algo = (vq_tc_supported(D, Kc) && Kc >= 64 && n >= 1024) ? 2 : 1;
  }
  if (algo == 2) {
    VQ_CHECK_ARG(vq_tc_supported(D, Kc), "vq_fused_fwd: tensor-core search supports D in {3, 4}, Kc <= 2048 (got D=%d Kc=%d)", D, Kc);
    static const int variant = getenv("VQCPC_VQTC_LAYOUT") ? atoi(getenv("VQCPC_VQTC_LAYOUT")) : 0;
    return vq_tc_fused_fwd(x, n, D, E, T, Kc, zd, beta, squared, idx, loss, z_sg, y, y_row0, y_rows, counts, dbg_scores, stats,
                           variant, st);
  }
  return vq_fused_fwd_direct(x, n, D, E, T, Kc, zd, beta, squared, idx, loss, z_sg, y, y_row0, y_rows, counts, st);
}

VQ_API int vqcpc_vq_fused_fwd_ex(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta,
                                 int squared, int64_t* idx, float* loss, float* z_sg, float* y, int* counts, int algo,
                                 float* dbg_scores, int* stats, void* stream) {
  return vq_fused_fwd_any(x, n, D, E, T, Kc, zd, beta, squared, idx, loss, z_sg, y, 0, n, counts, algo, dbg_scores, stats, stream);
}

// Index-path form: the upscaled rows y = T[idx] are produced only for the rows [y_row0, y_row0 + y_rows) (y has y_rows rows).
VQ_API int vqcpc_vq_fused_fwd_rows(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta,
                                   int squared, int64_t* idx, float* loss, float* z_sg, float* y, int64_t y_row0, int64_t y_rows,
                                   int* counts, void* stream) {
  return vq_fused_fwd_any(x, n, D, E, T, Kc, zd, beta, squared, idx, loss, z_sg, y, y_row0, y_rows, counts, 0, nullptr, nullptr,
                          stream);
}

VQ_API int vqcpc_vq_fused_fwd(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta,
                              int squared, int64_t* idx, float* loss, float* z_sg, float* y, int* counts, void* stream) {
  return vqcpc_vq_fused_fwd_ex(x, n, D, E, T, Kc, zd, beta, squared, idx, loss, z_sg, y, counts, 0, nullptr, nullptr, stream);
}

// Backward of the fused path: d_x (n, D) overwritten; d_T (Kc, zd) and d_E (Kc, D) ACCUMULATED.
// J (Kc, zd, D): Jacobian of the upscaler at each code vector; dy (n, zd); g_loss (n) or NULL.
VQ_API int vqcpc_vq_fused_bwd(const float* x, int64_t n, int D, const float* E, const float* J, int Kc, int zd,
                              const int64_t* idx, float beta, int squared, const float* dy, const float* g_loss, float* d_x,
                              float* d_T, float* d_E, void* stream) {
  VQ_CHECK_ARG(D == 3 || D == 4 || D == 8 || D == 16 || D == 32, "vq_fused_bwd: D must be 3, 4, 8, 16 or 32 (got %d)", D);
  if (n == 0) return VQ_OK;
  size_t smem = (size_t)Kc * (zd + D) * sizeof(float);
  VQ_CHECK_ARG(smem <= 200 * 1024, "vq_fused_bwd: gradient tables (%zu B) do not fit in shared memory", smem);
  cudaStream_t st = vq_stream(stream);
  if (zd == 32 && (D == 3 || D == 4) && n >= 4096 && (((uintptr_t)dy | (uintptr_t)J) & 15) == 0) {
    // streaming kernel: one CTA per SM, clusters of 4 (148 = 4 x 37 on B200; any SM count that is a multiple of 4)
    int sms = vq_sm_count();
    int grid = (int)std::min<int64_t>(vq_cdiv(n, 512), sms) / 4 * 4;
    VQ_CHECK_ARG(Kc <= 65535, "vq_fused_bwd: Kc=%d too large", Kc);
    if (grid >= 4) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(VB_THREADS);
      const size_t smem2 = (size_t)Kc * (VB_TS + D) * 4 + (size_t)(4 * Kc + 2 + VB_SUB / VB_ROWS) * 4 + VB_SUB * 2 + 16;
      cfg.dynamicSmemBytes = smem2;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 4;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (D == 3) {
        VQ_CUDA_CALL(cudaFuncSetAttribute(vq_fused_bwd_stream_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        VQ_CUDA_CALL(cudaLaunchKernelEx(&cfg, vq_fused_bwd_stream_kernel<3>, x, E, J, Kc, n, idx, beta, squared, dy, g_loss, d_x, d_T, d_E));
      } else {
        VQ_CUDA_CALL(cudaFuncSetAttribute(vq_fused_bwd_stream_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        VQ_CUDA_CALL(cudaLaunchKernelEx(&cfg, vq_fused_bwd_stream_kernel<4>, x, E, J, Kc, n, idx, beta, squared, dy, g_loss, d_x, d_T, d_E));
      }
      VQ_CUDA_LAUNCH_CHECK("vq_fused_bwd_stream_kernel");
      return VQ_OK;
    }
  }
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(n, 256), (int64_t)vq_sm_count() * (smem > 100 * 1024 ? 1 : 2)));
#define VQ_BL(DSV)                                                                                                          \
  do {                                                                                                                      \
    VQ_CUDA_CALL(cudaFuncSetAttribute(vq_fused_bwd_kernel<DSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    vq_fused_bwd_kernel<DSV><<<grid, 256, smem, st>>>(x, E, J, Kc, zd, n, idx, beta, squared, dy, g_loss, d_x, d_T, d_E);   \
  } while (0)
  switch (D) {
    case 3: VQ_BL(3); break;
    case 4: VQ_BL(4); break;
    case 8: VQ_BL(8); break;
    case 16: VQ_BL(16); break;
    default: VQ_BL(32); break;
  }
#undef VQ_BL
  VQ_CUDA_LAUNCH_CHECK("vq_fused_bwd_kernel");
  return VQ_OK;
}
