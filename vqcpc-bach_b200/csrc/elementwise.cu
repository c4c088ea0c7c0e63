// HBM-bound elementwise / gather / reduction kernels of the encoder path (everything that is not a GEMM,
// the per-block attention, the quantizer or InfoNCE).  All fp32; coalesced float4 accesses; grids sized from the
// SM count.  Reference semantics are cited at each entry point (paths relative to the reference repo).
#include "common.cuh"

namespace {

// (hi, lo) bf16 split of 4 floats, stored as two 8-byte vectors
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, float4 v) {
  __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                h3 = __float2bfloat16_rn(v.w);
  __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1)),
                l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
  uint2 hv, lv;
  hv.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  hv.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
  lv.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  lv.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
  *reinterpret_cast<uint2*>(hi) = hv;
  *reinterpret_cast<uint2*>(lo) = lv;
}

constexpr int MAX_VOICES = 8;
struct TablePtrs {
  const float* t[MAX_VOICES];
  float* g[MAX_VOICES];
  int rows[MAX_VOICES];
  int offset[MAX_VOICES + 1];   // row offset of each voice inside the packed smem copy
};

// ----------------------------------------------------------------------------------------------
// embedding gather: out[i, :] = table[i % V][tok[i], :]
// ----------------------------------------------------------------------------------------------
__global__ void embed_fwd_kernel(const int64_t* __restrict__ tok, TablePtrs tp, int V, int E4, float* __restrict__ out,
                                 int64_t n) {
  int64_t total = n * E4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / E4;
    int c = (int)(i % E4);
    int v = (int)(r % V);
    int64_t t = tok[r];
    t = t < 0 ? 0 : (t >= tp.rows[v] ? tp.rows[v] - 1 : t);
    st4(out + r * (int64_t)E4 * 4 + c * 4, ld4(tp.t[v] + t * (int64_t)E4 * 4 + c * 4));
  }
}

// embedding backward with a CTA-private copy of all tables in shared memory (red.shared), flushed once per CTA
__global__ void embed_bwd_smem_kernel(const int64_t* __restrict__ tok, TablePtrs tp, int V, int E, const float* __restrict__ dout,
                                      int64_t n, int total_rows) {
  extern __shared__ float acc[];
  for (int i = threadIdx.x; i < total_rows * E; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  int64_t total = n * E;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / E;
    int c = (int)(i % E);
    int v = (int)(r % V);
    int64_t t = tok[r];
    t = t < 0 ? 0 : (t >= tp.rows[v] ? tp.rows[v] - 1 : t);
    atomicAdd(&acc[(tp.offset[v] + (int)t) * E + c], dout[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < total_rows * E; i += blockDim.x) {
    float a = acc[i];
    if (a != 0.f) {
      int row = i / E, c = i % E;
      int v = 0;
      while (v + 1 < V && row >= tp.offset[v + 1]) ++v;
      atomicAdd(tp.g[v] + (size_t)(row - tp.offset[v]) * E + c, a);
    }
  }
}

__global__ void embed_bwd_global_kernel(const int64_t* __restrict__ tok, TablePtrs tp, int V, int E, const float* __restrict__ dout,
                                        int64_t n) {
  int64_t total = n * E;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / E;
    int c = (int)(i % E);
    int v = (int)(r % V);
    int64_t t = tok[r];
    t = t < 0 ? 0 : (t >= tp.rows[v] ? tp.rows[v] - 1 : t);
    atomicAdd(tp.g[v] + (size_t)t * E + c, dout[i]);
  }
}

// ----------------------------------------------------------------------------------------------
// positional embeddings: h[m, d-2pe + j] = chan[(m%S)%V][j] (j<pe) | evt[(m%S)/V][j-pe]
// ----------------------------------------------------------------------------------------------
__global__ void posemb_fill_kernel(float* __restrict__ h, int ld, int col0, const float* __restrict__ chan,
                                   const float* __restrict__ evt, int64_t M, int S, int V, int pe) {
  int w = 2 * pe;
  int64_t total = M * w;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = i / w;
    int j = (int)(i % w);
    int t = (int)(m % S);
    float v = j < pe ? chan[(t % V) * pe + j] : evt[(t / V) * pe + (j - pe)];
    h[m * ld + col0 + j] = v;
  }
}

// blockDim = S * 2pe threads: thread (t, j) sums column j over all rows with position t, then one atomic.
__global__ void posemb_bwd_kernel(const float* __restrict__ dh, int ld, int col0, float* __restrict__ dchan,
                                  float* __restrict__ devt, int64_t nblocks, int S, int V, int pe) {
  int w = 2 * pe;
  int t = threadIdx.x / w, j = threadIdx.x % w;
  float acc = 0.f;
  for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x) acc += dh[(b * S + t) * ld + col0 + j];
  if (j < pe) atomicAdd(dchan + (t % V) * pe + j, acc);
  else atomicAdd(devt + (t / V) * pe + (j - pe), acc);
}

// ----------------------------------------------------------------------------------------------
// LayerNorm (eps inside the sqrt, biased variance), one warp per row, row kept in registers
// ----------------------------------------------------------------------------------------------
template <int VPL>   // float4 per lane; d = 128 * VPL
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd, int64_t M, float eps,
                     __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const int d = 128 * VPL;
  int lane = threadIdx.x & 31;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g[VPL], b[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    g[i] = ld4(gamma + (i * 32 + lane) * 4);
    b[i] = ld4(beta + (i * 32 + lane) * 4);
  }
  for (int64_t r = warp; r < M; r += nwarps) {
    float4 x[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      x[i] = ld4(y + r * d + (i * 32 + lane) * 4);
      s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    }
    float mu = warp_sum(s) * (1.f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float a0 = x[i].x - mu, a1 = x[i].y - mu, a2 = x[i].z - mu, a3 = x[i].w - mu;
      q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    float var = warp_sum(q) * (1.f / d);
    float rs = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 o;
      o.x = (x[i].x - mu) * rs * g[i].x + b[i].x;
      o.y = (x[i].y - mu) * rs * g[i].y + b[i].y;
      o.z = (x[i].z - mu) * rs * g[i].z + b[i].z;
      o.w = (x[i].w - mu) * rs * g[i].w + b[i].w;
      if (out) st4(out + r * d + (i * 32 + lane) * 4, o);
      if (out_hi) store_split4(out_hi + r * d + (i * 32 + lane) * 4, out_lo + r * d + (i * 32 + lane) * 4, o);
    }
    if (lane == 0) {
      mean[r] = mu;
      rstd[r] = rs;
    }
  }
}

template <int VPL, bool COLSUM>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, float* __restrict__ dy,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M,
                     __nv_bfloat16* __restrict__ dy_hi, __nv_bfloat16* __restrict__ dy_lo, float* __restrict__ dy_colsum) {
  const int d = 128 * VPL;
  __shared__ float red[2][8][128 * VPL];   // [dgamma|dbeta][warp][col]   (VPL=4: 32 KB)
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g[VPL], ag[VPL], ab[VPL], ad[COLSUM ? VPL : 1];   // ad: column sums of dy (the bias gradient of the layer feeding the LN)
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    if (!COLSUM) g[i] = ld4(gamma + (i * 32 + lane) * 4);
    ag[i] = make_float4(0, 0, 0, 0);
    ab[i] = make_float4(0, 0, 0, 0);
    if (COLSUM) ad[i] = make_float4(0, 0, 0, 0);
  }
  for (int64_t r = warp; r < M; r += nwarps) {
    float mu = mean[r], rs = rstd[r];
    float4 xh[VPL], go[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 yy = ld4(y + r * d + (i * 32 + lane) * 4);
      go[i] = ld4(dout + r * d + (i * 32 + lane) * 4);
      xh[i] = make_float4((yy.x - mu) * rs, (yy.y - mu) * rs, (yy.z - mu) * rs, (yy.w - mu) * rs);
      ag[i].x += go[i].x * xh[i].x; ag[i].y += go[i].y * xh[i].y; ag[i].z += go[i].z * xh[i].z; ag[i].w += go[i].w * xh[i].w;
      ab[i].x += go[i].x; ab[i].y += go[i].y; ab[i].z += go[i].z; ab[i].w += go[i].w;
      const float4 gm = COLSUM ? ld4(gamma + (i * 32 + lane) * 4) : g[i];   // (COLSUM: gamma re-read from L1, 16 registers freed)
      go[i].x *= gm.x; go[i].y *= gm.y; go[i].z *= gm.z; go[i].w *= gm.w;
      s1 += (go[i].x + go[i].y) + (go[i].z + go[i].w);
      s2 += (go[i].x * xh[i].x + go[i].y * xh[i].y) + (go[i].z * xh[i].z + go[i].w * xh[i].w);
    }
    float c1 = warp_sum(s1) * (1.f / d), c2 = warp_sum(s2) * (1.f / d);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 o;
      o.x = (go[i].x - c1 - xh[i].x * c2) * rs;
      o.y = (go[i].y - c1 - xh[i].y * c2) * rs;
      o.z = (go[i].z - c1 - xh[i].z * c2) * rs;
      o.w = (go[i].w - c1 - xh[i].w * c2) * rs;
      st4(dy + r * d + (i * 32 + lane) * 4, o);
      if (dy_hi) store_split4(dy_hi + r * d + (i * 32 + lane) * 4, dy_lo + r * d + (i * 32 + lane) * 4, o);
      if (COLSUM) { ad[i].x += o.x; ad[i].y += o.y; ad[i].z += o.z; ad[i].w += o.w; }
    }
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    st4(&red[0][wib][(i * 32 + lane) * 4], ag[i]);
    st4(&red[1][wib][(i * 32 + lane) * 4], ab[i]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += red[0][w][c];
      b += red[1][w][c];
    }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
  }
  if (COLSUM) {                              // second round through the same scratch
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i) st4(&red[0][wib][(i * 32 + lane) * 4], ad[i]);
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) a += red[0][w][c];
      atomicAdd(dy_colsum + c, a);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// strided row gather / scatter inside blocks (the downscaler's x[::factor])
// ----------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows_out, int Lout,
                                   int L, int factor, int d4) {
  int64_t total = rows_out * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / d4;
    int c = (int)(i % d4);
    int64_t p = r / Lout;
    int t = (int)(r % Lout);
    st4(out + i * 4, ld4(x + ((p * L + (int64_t)t * factor) * d4 + c) * 4));
  }
}

__global__ void scatter_rows_kernel(const float* __restrict__ dout, float* __restrict__ dx, int64_t rows_in, int Lout,
                                    int L, int factor, int d4) {
  int64_t total = rows_in * d4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / d4;
    int c = (int)(i % d4);
    int64_t p = r / L;
    int t = (int)(r % L);
    float4 v = make_float4(0, 0, 0, 0);
    if (t % factor == 0) v = ld4(dout + ((p * Lout + t / factor) * d4 + c) * 4);
    st4(dx + i * 4, v);
  }
}

// ----------------------------------------------------------------------------------------------
// column sums (bias gradients): out[n] += sum_m x[m, n]
// ----------------------------------------------------------------------------------------------
// each warp covers 128 columns (float4 per lane); 8 warps stride over the rows of the CTA's row range.
// SPLIT: x is given as bf16 (hi, lo) and summed as hi + lo.
template <bool SPLIT>
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl, int ld,
              float* __restrict__ out, int64_t M, int N, int64_t rows_per_cta) {
  __shared__ float4 red[8][32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int col = blockIdx.x * 128 + lane * 4;
  int64_t r0 = blockIdx.y * rows_per_cta;
  int64_t r1 = min(M, r0 + rows_per_cta);
  float4 a = make_float4(0, 0, 0, 0);
  auto load_row = [&](int64_t r) -> float4 {
    float4 v;
    if (SPLIT) {
      uint2 h = *reinterpret_cast<const uint2*>(xh + r * ld + col);
      uint2 l = *reinterpret_cast<const uint2*>(xl + r * ld + col);
      v.x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
      v.y = __uint_as_float(h.x & 0xFFFF0000u) + __uint_as_float(l.x & 0xFFFF0000u);
      v.z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
      v.w = __uint_as_float(h.y & 0xFFFF0000u) + __uint_as_float(l.y & 0xFFFF0000u);
    } else {
      v = ld4(x + r * ld + col);
    }
    return v;
  };
  if (col < N) {
    // eight independent row loads in flight per warp (a warp-wide row segment is 512 B): with ~32 warps per SM that is
    // well above the ~45 KB per SM an HBM-bound stream needs; the plain loop had one load in flight per warp
    int64_t r = r0 + w;
    for (; r + 56 < r1; r += 64) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = load_row(r + 8 * u);
#pragma unroll
      for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
    }
    for (; r < r1; r += 8) {
      const float4 v = load_row(r);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && col < N) {
    float4 s = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += red[i][lane].x; s.y += red[i][lane].y; s.z += red[i][lane].z; s.w += red[i][lane].w; }
    atomicAdd(out + col, s.x); atomicAdd(out + col + 1, s.y); atomicAdd(out + col + 2, s.z); atomicAdd(out + col + 3, s.w);
  }
}

// generic (any N / ld) scalar fallback
__global__ void __launch_bounds__(256)
colsum_scalar_kernel(const float* __restrict__ x, int ld, float* __restrict__ out, int64_t M, int N, int64_t rows_per_cta) {
  __shared__ float red[8][32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int col = blockIdx.x * 32 + lane;
  int64_t r0 = blockIdx.y * rows_per_cta;
  int64_t r1 = min(M, r0 + rows_per_cta);
  float a = 0.f;
  if (col < N)
    for (int64_t r = r0 + w; r < r1; r += 8) a += x[r * ld + col];
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && col < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][lane];
    atomicAdd(out + col, s);
  }
}

// ----------------------------------------------------------------------------------------------
// SELU
// ----------------------------------------------------------------------------------------------
constexpr float SELU_SCALE = 1.0507009873554804934193349852946f;
constexpr float SELU_ALPHA = 1.6732632423543772848170429916717f;

__global__ void selu_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    out[i] = v > 0.f ? SELU_SCALE * v : (SELU_SCALE * SELU_ALPHA) * expm1f(v);
  }
}
__global__ void selu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    dx[i] = dy[i] * (v > 0.f ? SELU_SCALE : (SELU_SCALE * SELU_ALPHA) * expf(v));
  }
}

// ----------------------------------------------------------------------------------------------
// inverted dropout with an explicit keep mask (nn.Dropout, mlp_upscaler.py:23): out = x * keep * scale, scale = 1/(1-p).
// The mask is an input (one byte per element) so forward and backward see the same draw and tests can replay it.
// ----------------------------------------------------------------------------------------------
__global__ void mask_scale_kernel(const float* __restrict__ x, const unsigned char* __restrict__ keep, float scale,
                                  float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = keep[i] ? x[i] * scale : 0.f;
}

// out = (keep ? x * scale : 0) + residual, four elements per thread; fp32 and / or bf16-split outputs (the dropout sites
// of the encoder layer, transformer_custom.py:278-290: dropout1 / dropout2 before the residual add, `dropout` after the
// activation, and the same mask applied to the incoming gradient in backward)
__global__ void mask_scale_add_kernel(const float* __restrict__ x, const unsigned char* __restrict__ keep, float scale,
                                      const float* __restrict__ residual, float* __restrict__ out,
                                      __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = ld4(x + i * 4);
    const uchar4 k = *reinterpret_cast<const uchar4*>(keep + i * 4);
    v.x = k.x ? v.x * scale : 0.f; v.y = k.y ? v.y * scale : 0.f; v.z = k.z ? v.z * scale : 0.f; v.w = k.w ? v.w * scale : 0.f;
    if (residual) {
      const float4 r = ld4(residual + i * 4);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (out) st4(out + i * 4, v);
    if (out_hi) {
      __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
      __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1)),
                    l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
      __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(out_hi + i * 4);
      __nv_bfloat162* lp = reinterpret_cast<__nv_bfloat162*>(out_lo + i * 4);
      hp[0] = __nv_bfloat162(h0, h1); hp[1] = __nv_bfloat162(h2, h3);
      lp[0] = __nv_bfloat162(l0, l1); lp[1] = __nv_bfloat162(l2, l3);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// fp32 -> (bf16 hi, bf16 lo) split for the BF16x3 tensor-core GEMM:  x ~= hi + lo  (16 significand bits)
// ----------------------------------------------------------------------------------------------
__global__ void split_bf16_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int ldo, int64_t rows, int cols4) {
  int64_t total = rows * cols4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols4;
    int c = (int)(i % cols4) * 4;
    float4 v = ld4(x + r * ldx + c);
    __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                  h3 = __float2bfloat16_rn(v.w);
    __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1)),
                  l2 = __float2bfloat16_rn(v.z - __bfloat162float(h2)), l3 = __float2bfloat16_rn(v.w - __bfloat162float(h3));
    __nv_bfloat162 hp0 = __halves2bfloat162(h0, h1), hp1 = __halves2bfloat162(h2, h3);
    __nv_bfloat162 lp0 = __halves2bfloat162(l0, l1), lp1 = __halves2bfloat162(l2, l3);
    uint2 hv, lv;
    hv.x = *reinterpret_cast<uint32_t*>(&hp0); hv.y = *reinterpret_cast<uint32_t*>(&hp1);
    lv.x = *reinterpret_cast<uint32_t*>(&lp0); lv.y = *reinterpret_cast<uint32_t*>(&lp1);
    *reinterpret_cast<uint2*>(hi + r * ldo + c) = hv;
    *reinterpret_cast<uint2*>(lo + r * ldo + c) = lv;
  }
}

// deterministic single-CTA sum: out[0] = scale * sum(x)
__global__ void __launch_bounds__(1024) reduce_sum_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out, float scale) {
  __shared__ double red[32];
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += (double)x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[0] = (float)(v * (double)scale);
  }
}

// Histogram of code indices: counts[idx[i]] += 1 (int32, ACCUMULATED).  Per-CTA privatised in shared memory with native
// integer atomics (ATOMS.ADD), one global atomic per (CTA, non-empty bin).  Out-of-range values are ignored.
__global__ void __launch_bounds__(512) code_hist_kernel(const int64_t* __restrict__ idx, int64_t n, int bins, int* __restrict__ counts) {
  extern __shared__ int hist_s[];
  for (int i = threadIdx.x; i < bins; i += blockDim.x) hist_s[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t k = idx[i];
    if (k >= 0 && k < bins) atomicAdd(&hist_s[(int)k], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x)
    if (hist_s[i]) atomicAdd(counts + i, hist_s[i]);
}

inline int grid_for(int64_t work_items, int threads, int per_sm = 8) {
  int64_t g = vq_cdiv(work_items, threads);
  int64_t cap = (int64_t)vq_sm_count() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

int fill_tables(TablePtrs& tp, int V, const float* const* tables, float* const* grads, const int* rows) {
  if (V < 1 || V > MAX_VOICES) return -1;
  int off = 0;
  for (int v = 0; v < V; ++v) {
    tp.t[v] = tables ? tables[v] : nullptr;
    tp.g[v] = grads ? grads[v] : nullptr;
    tp.rows[v] = rows[v];
    tp.offset[v] = off;
    off += rows[v];
  }
  tp.offset[V] = off;
  return off;
}

}  // namespace

// data_processor/bach_cpc_data_processor.py:42-68 (per-voice nn.Embedding lookup, voice = position % V)
VQ_API int vqcpc_embed_fwd(const int64_t* tokens, int64_t n, int V, const float* const* tables, const int* rows, int E,
                           float* out, void* stream) {
  VQ_CHECK_ARG(E % 4 == 0 && E > 0, "embed_fwd: embedding size must be a multiple of 4 (got %d)", E);
  TablePtrs tp;
  VQ_CHECK_ARG(fill_tables(tp, V, tables, nullptr, rows) > 0, "embed_fwd: 1..%d voices supported", MAX_VOICES);
  if (n == 0) return VQ_OK;
  embed_fwd_kernel<<<grid_for(n * (E / 4), 256), 256, 0, vq_stream(stream)>>>(tokens, tp, V, E / 4, out, n);
  VQ_CUDA_LAUNCH_CHECK("embed_fwd_kernel");
  return VQ_OK;
}

// gradient of the lookup: d_table[v][tok] += d_out  (accumulates into grads)
VQ_API int vqcpc_embed_bwd(const int64_t* tokens, int64_t n, int V, float* const* grads, const int* rows, int E,
                           const float* d_out, void* stream) {
  TablePtrs tp;
  int total_rows = fill_tables(tp, V, nullptr, grads, rows);
  VQ_CHECK_ARG(total_rows > 0, "embed_bwd: 1..%d voices supported", MAX_VOICES);
  if (n == 0) return VQ_OK;
  size_t smem = (size_t)total_rows * E * sizeof(float);
  cudaStream_t st = vq_stream(stream);
  if (smem <= 96 * 1024) {
    static bool attr_done = false;
    if (!attr_done) {
      VQ_CUDA_CALL(cudaFuncSetAttribute(embed_bwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_done = true;
    }
    int grid = grid_for(n * E, 512 * 16, 2);
    embed_bwd_smem_kernel<<<grid, 512, smem, st>>>(tokens, tp, V, E, d_out, n, total_rows);
    VQ_CUDA_LAUNCH_CHECK("embed_bwd_smem_kernel");
  } else {
    embed_bwd_global_kernel<<<grid_for(n * E, 256), 256, 0, st>>>(tokens, tp, V, E, d_out, n);
    VQ_CUDA_LAUNCH_CHECK("embed_bwd_global_kernel");
  }
  return VQ_OK;
}

// downscalers/relative_transformer_downscaler.py:106-112: the two 8-wide positional embeddings concatenated after
// the input projection; written into columns [col0, col0 + 2*pe) of h (leading dimension ld).
VQ_API int vqcpc_posemb_fill(float* h, int ld, int col0, const float* chan, const float* evt, int64_t M, int S, int V,
                             int pe, void* stream) {
  VQ_CHECK_ARG(S % V == 0 && M % S == 0, "posemb_fill: rows (%lld) must be whole blocks of S=%d tokens, V=%d", (long long)M, S, V);
  if (M == 0) return VQ_OK;
  posemb_fill_kernel<<<grid_for(M * 2 * pe, 256), 256, 0, vq_stream(stream)>>>(h, ld, col0, chan, evt, M, S, V, pe);
  VQ_CUDA_LAUNCH_CHECK("posemb_fill_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_posemb_bwd(const float* dh, int ld, int col0, float* dchan, float* devt, int64_t M, int S, int V, int pe,
                            void* stream) {
  VQ_CHECK_ARG(S % V == 0 && M % S == 0, "posemb_bwd: rows must be whole blocks");
  VQ_CHECK_ARG(S * 2 * pe <= 1024, "posemb_bwd: S*2*pe = %d exceeds one CTA", S * 2 * pe);
  if (M == 0) return VQ_OK;
  int64_t nblocks = M / S;
  int grid = (int)std::min<int64_t>(nblocks, (int64_t)vq_sm_count() * 4);
  posemb_bwd_kernel<<<grid, S * 2 * pe, 0, vq_stream(stream)>>>(dh, ld, col0, dchan, devt, nblocks, S, V, pe);
  VQ_CUDA_LAUNCH_CHECK("posemb_bwd_kernel");
  return VQ_OK;
}

// transformer/transformer_custom.py:283,289 (nn.LayerNorm(d_model), eps=1e-5); saves mean / rstd for backward
VQ_API int vqcpc_layernorm_fwd(const float* y, const float* gamma, const float* beta, float* out, void* out_hi, void* out_lo,
                               float* mean, float* rstd, int64_t M, int d, float eps, void* stream) {
  VQ_CHECK_ARG(d == 128 || d == 256 || d == 512 || d == 1024, "layernorm: d_model must be 128/256/512/1024 (got %d)", d);
  VQ_CHECK_ARG((out || out_hi) && ((out_hi == nullptr) == (out_lo == nullptr)), "layernorm: need out and/or (out_hi, out_lo)");
  if (M == 0) return VQ_OK;
  int grid = grid_for(M, 8, 8);
  cudaStream_t st = vq_stream(stream);
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
  switch (d / 128) {
    case 1: layernorm_fwd_kernel<1><<<grid, 256, 0, st>>>(y, gamma, beta, out, mean, rstd, M, eps, oh, ol); break;
    case 2: layernorm_fwd_kernel<2><<<grid, 256, 0, st>>>(y, gamma, beta, out, mean, rstd, M, eps, oh, ol); break;
    case 4: layernorm_fwd_kernel<4><<<grid, 256, 0, st>>>(y, gamma, beta, out, mean, rstd, M, eps, oh, ol); break;
    default: layernorm_fwd_kernel<8><<<grid, 256, 0, st>>>(y, gamma, beta, out, mean, rstd, M, eps, oh, ol); break;
  }
  VQ_CUDA_LAUNCH_CHECK("layernorm_fwd_kernel");
  return VQ_OK;
}

// d_y = LN backward wrt its input; d_gamma / d_beta are ACCUMULATED (+=).  d_y_colsum (d floats, may be NULL) additionally
// accumulates the column sums of d_y: the bias gradient of the linear layer whose output (+ residual) the LN normalises
// (out_proj.bias for norm1, linear2.bias for norm2; transformer_custom.py:283-289) - saves a separate pass over d_y.
VQ_API int vqcpc_layernorm_bwd_ex(const float* d_out, const float* y, const float* mean, const float* rstd, const float* gamma,
                                  float* d_y, void* d_y_hi, void* d_y_lo, float* d_gamma, float* d_beta, float* d_y_colsum,
                                  int64_t M, int d, void* stream) {
  VQ_CHECK_ARG(d == 128 || d == 256 || d == 512, "layernorm_bwd: d_model must be 128/256/512 (got %d)", d);
  VQ_CHECK_ARG((d_y_hi == nullptr) == (d_y_lo == nullptr), "layernorm_bwd: d_y_hi/d_y_lo must be given together");
  if (M == 0) return VQ_OK;
  int grid = grid_for(M, 8 * 16, 4);
  cudaStream_t st = vq_stream(stream);
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(d_y_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(d_y_lo);
#define VQ_LN_BWD(V, C) layernorm_bwd_kernel<V, C><<<grid, 256, 0, st>>>(d_out, y, mean, rstd, gamma, d_y, d_gamma, d_beta, M, oh, ol, d_y_colsum)
  if (d_y_colsum) {
    switch (d / 128) {
      case 1: VQ_LN_BWD(1, true); break;
      case 2: VQ_LN_BWD(2, true); break;
      default: VQ_LN_BWD(4, true); break;
    }
  } else {
    switch (d / 128) {
      case 1: VQ_LN_BWD(1, false); break;
      case 2: VQ_LN_BWD(2, false); break;
      default: VQ_LN_BWD(4, false); break;
    }
  }
#undef VQ_LN_BWD
  VQ_CUDA_LAUNCH_CHECK("layernorm_bwd_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_layernorm_bwd(const float* d_out, const float* y, const float* mean, const float* rstd, const float* gamma,
                               float* d_y, void* d_y_hi, void* d_y_lo, float* d_gamma, float* d_beta, int64_t M, int d,
                               void* stream) {
  return vqcpc_layernorm_bwd_ex(d_out, y, mean, rstd, gamma, d_y, d_y_hi, d_y_lo, d_gamma, d_beta, nullptr, M, d, stream);
}

// relative_transformer_downscaler.py:124  transfo_input = output[::downscaling]  (blocks-major layout here)
VQ_API int vqcpc_subsample_fwd(const float* x, float* out, int64_t P, int L, int factor, int d, void* stream) {
  VQ_CHECK_ARG(L % factor == 0 && d % 4 == 0, "subsample_fwd: L=%d factor=%d d=%d", L, factor, d);
  if (P == 0) return VQ_OK;
  int Lout = L / factor;
  gather_rows_kernel<<<grid_for(P * Lout * (d / 4), 256), 256, 0, vq_stream(stream)>>>(x, out, P * Lout, Lout, L, factor, d / 4);
  VQ_CUDA_LAUNCH_CHECK("gather_rows_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_subsample_bwd(const float* d_out, float* d_x, int64_t P, int L, int factor, int d, void* stream) {
  VQ_CHECK_ARG(L % factor == 0 && d % 4 == 0, "subsample_bwd: L=%d factor=%d d=%d", L, factor, d);
  if (P == 0) return VQ_OK;
  scatter_rows_kernel<<<grid_for(P * L * (d / 4), 256), 256, 0, vq_stream(stream)>>>(d_out, d_x, P * L, L / factor, L, factor, d / 4);
  VQ_CUDA_LAUNCH_CHECK("scatter_rows_kernel");
  return VQ_OK;
}

static int colsum_launch(const float* x, const void* xh, const void* xl, int ld, float* out, int64_t M, int N, cudaStream_t st) {
  if (M == 0) return VQ_OK;
  bool vec = (N % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x ? (const void*)x : xh) & 15) == 0 || xh);
  if (xh) VQ_CHECK_ARG(N % 4 == 0 && ld % 4 == 0, "colsum_split: N and ld must be multiples of 4");
  int gx = (int)vq_cdiv(N, vec ? 128 : 32);
  int64_t want_y = std::max<int64_t>(1, (int64_t)vq_sm_count() * 4 / gx);
  int64_t rows_per_cta = std::max<int64_t>(64, vq_cdiv(M, want_y));
  int gy = (int)vq_cdiv(M, rows_per_cta);
  VQ_CHECK_ARG(gy <= 65535, "colsum: grid too large");
  if (xh)
    colsum_kernel<true><<<dim3(gx, gy), 256, 0, st>>>(nullptr, reinterpret_cast<const __nv_bfloat16*>(xh),
                                                      reinterpret_cast<const __nv_bfloat16*>(xl), ld, out, M, N, rows_per_cta);
  else if (vec)
    colsum_kernel<false><<<dim3(gx, gy), 256, 0, st>>>(x, nullptr, nullptr, ld, out, M, N, rows_per_cta);
  else
    colsum_scalar_kernel<<<dim3(gx, gy), 256, 0, st>>>(x, ld, out, M, N, rows_per_cta);
  VQ_CUDA_LAUNCH_CHECK("colsum_kernel");
  return VQ_OK;
}

// bias gradients: out[n] += sum_m x[m, n]
VQ_API int vqcpc_colsum(const float* x, int ld, float* out, int64_t M, int N, void* stream) {
  return colsum_launch(x, nullptr, nullptr, ld, out, M, N, vq_stream(stream));
}
// same on a bf16 (hi, lo) split matrix: out[n] += sum_m (hi + lo)[m, n]
VQ_API int vqcpc_colsum_split(const void* x_hi, const void* x_lo, int ld, float* out, int64_t M, int N, void* stream) {
  VQ_CHECK_ARG(x_hi && x_lo, "colsum_split: null input");
  return colsum_launch(nullptr, x_hi, x_lo, ld, out, M, N, vq_stream(stream));
}

// upscalers/mlp_upscaler.py:21-26 nn.SELU
VQ_API int vqcpc_selu_fwd(const float* x, float* out, int64_t n, void* stream) {
  if (n == 0) return VQ_OK;
  selu_fwd_kernel<<<grid_for(n, 256), 256, 0, vq_stream(stream)>>>(x, out, n);
  VQ_CUDA_LAUNCH_CHECK("selu_fwd_kernel");
  return VQ_OK;
}
VQ_API int vqcpc_selu_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream) {
  if (n == 0) return VQ_OK;
  selu_bwd_kernel<<<grid_for(n, 256), 256, 0, vq_stream(stream)>>>(x, dy, dx, n);
  VQ_CUDA_LAUNCH_CHECK("selu_bwd_kernel");
  return VQ_OK;
}

// nn.Dropout with an explicit keep mask (u8, one per element): out = keep ? x * scale : 0
VQ_API int vqcpc_mask_scale(const float* x, const unsigned char* keep, float scale, float* out, int64_t n, void* stream) {
  if (n == 0) return VQ_OK;
  mask_scale_kernel<<<grid_for(n, 256), 256, 0, vq_stream(stream)>>>(x, keep, scale, out, n);
  VQ_CUDA_LAUNCH_CHECK("mask_scale_kernel");
  return VQ_OK;
}

// dropout site of the encoder layer: out = (keep ? x * scale : 0) (+ residual); contiguous (n elements, n % 4 == 0);
// fp32 `out` and / or the bf16 split (out_hi, out_lo) may be requested
VQ_API int vqcpc_mask_scale_add(const float* x, const unsigned char* keep, float scale, const float* residual, float* out,
                                void* out_hi, void* out_lo, int64_t n, void* stream) {
  VQ_CHECK_ARG(n % 4 == 0, "mask_scale_add: n must be a multiple of 4 (got %lld)", (long long)n);
  VQ_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr), "mask_scale_add: out_hi and out_lo go together");
  if (n == 0) return VQ_OK;
  mask_scale_add_kernel<<<grid_for(n / 4, 256), 256, 0, vq_stream(stream)>>>(
      x, keep, scale, residual, out, reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo), n / 4);
  VQ_CUDA_LAUNCH_CHECK("mask_scale_add_kernel");
  return VQ_OK;
}

// operand preparation for vqcpc_gemm_bf16x3: x (rows x cols, leading dim ldx) -> hi, lo (leading dim ldo)
VQ_API int vqcpc_split_bf16(const float* x, int ldx, void* hi, void* lo, int ldo, int64_t rows, int cols, void* stream) {
  VQ_CHECK_ARG(cols % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "split_bf16: cols/ld must be multiples of 4 (cols=%d ldx=%d ldo=%d)", cols, ldx, ldo);
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 &&
               (reinterpret_cast<uintptr_t>(lo) & 7) == 0, "split_bf16: misaligned pointer");
  if (rows == 0) return VQ_OK;
  split_bf16_kernel<<<grid_for(rows * (cols / 4), 256), 256, 0, vq_stream(stream)>>>(
      x, ldx, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), ldo, rows, cols / 4);
  VQ_CUDA_LAUNCH_CHECK("split_bf16_kernel");
  return VQ_OK;
}

// out[0] = scale * sum(x[0..n))   (fixed summation order -> deterministic)
VQ_API int vqcpc_reduce_sum(const float* x, int64_t n, float* out, float scale, void* stream) {
  reduce_sum_kernel<<<1, 1024, 0, vq_stream(stream)>>>(x, n, out, scale);
  VQ_CUDA_LAUNCH_CHECK("reduce_sum_kernel");
  return VQ_OK;
}

// Code-usage histogram (the per-code counts that ride in the data-parallel all-reduce; replaces torch.unique monitoring,
// VQCPCB/vqcpc_encoder_trainer.py:325-331): counts[idx[i]] += 1 for 0 <= idx[i] < bins.  counts int32, ACCUMULATED.
VQ_API int vqcpc_code_hist(const int64_t* idx, int64_t n, int bins, int* counts, void* stream) {
  VQ_CHECK_ARG(bins >= 1 && bins <= 12288, "vqcpc_code_hist: bins must be in [1, 12288] (got %d)", bins);
  if (n <= 0) return VQ_OK;
  code_hist_kernel<<<grid_for(n, 512, 2), 512, bins * sizeof(int), vq_stream(stream)>>>(idx, n, bins, counts);
  VQ_CUDA_LAUNCH_CHECK("code_hist_kernel");
  return VQ_OK;
}
