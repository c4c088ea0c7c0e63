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

namespace {

constexpr int MAX_CB = 8;
struct Books {
  const float* e[MAX_CB];
  float* g[MAX_CB];
};

// sum of (x_d - e_d)^2 over d in torch-CPU order (see header); x in registers when DS is a compile-time constant
template <int DS>
__device__ __forceinline__ float dist_fixed(const float* x, const float* __restrict__ e) {
  constexpr int W = (DS % 32 == 0) ? 32 : 8;
  float acc[W];
#pragma unroll
  for (int l = 0; l < W; ++l) acc[l] = 0.f;
#pragma unroll
  for (int d = 0; d < DS; ++d) {
    float df = __fsub_rn(x[d], e[d]);
    float sq = __fmul_rn(df, df);
    acc[d % W] = (d < W) ? sq : __fadd_rn(acc[d % W], sq);
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
