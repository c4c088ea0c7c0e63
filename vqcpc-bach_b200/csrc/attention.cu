// Per-block multi-head self-attention with the learned relative-position bias, forward and backward.
//
// Reference: VQCPCB/transformer/multihead_attention_custom.py:274-345 (head split, QK^T, +bias, softmax, PV) and
// VQCPCB/transformer/subsampled_relative_attention.py:30-122, whose pad/cat/view "skewing" reduces, for
// self-attention over one block of L tokens, to the closed form
//     bias[i][j] = q_i . e1[h][L-1-(i-j)]   (j <= i)         q already scaled by hd^-0.5
//                = q_i . e2[h][j-i]         (j >  i)
// so  S[i][j] = q_i . (k_j + E(i,j)).
//
// Blocks are independent length-L sequences (L = 16, then 4), rows are blocks-major: row = p*L + t.
// One warp owns a 16-row tile (= 16/L whole blocks) of one head (head_dim 64); attention is masked to the block.
// Everything is fp32 on CUDA cores (0.8 MFLOP per block-layer against ~100 MFLOP of GEMMs); the kernel is bound by
// the HBM read of QKV / write of O, so q/k/v/e are staged once in shared memory with float4 accesses.
#include "common.cuh"

namespace {

constexpr int HD = 64;        // head dim
constexpr int TS = 68;        // padded smem row stride (floats): 16-byte aligned, shifts rows by 4 banks
constexpr int TR = 16;        // rows per warp tile

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ void fma4(float4& acc, float s, float4 v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}

// stage a 16 x 64 tile (row stride ld in global) into smem [16][TS]; rows >= M are zero
__device__ __forceinline__ void load_tile(float* s, const float* __restrict__ g, int64_t row0, int64_t M, int64_t ld, int lane) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    int r = it * 2 + (lane >> 4);
    int c = (lane & 15) * 4;
    int64_t gr = row0 + r;
    float4 v = make_float4(0, 0, 0, 0);
    if (gr < M) v = ld4(g + gr * ld + c);
    st4(s + r * TS + c, v);
  }
}

// same, but the source holds only the rows with (token % keep == 0), compacted; other rows are zero
__device__ __forceinline__ void load_tile_compact(float* s, const float* __restrict__ g, int64_t row0, int64_t M, int64_t ld,
                                                  int lane, int L, int keep) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    int r = it * 2 + (lane >> 4);
    int c = (lane & 15) * 4;
    int64_t gr = row0 + r;
    int i = (int)(gr % L);
    float4 v = make_float4(0, 0, 0, 0);
    if (gr < M && (i % keep) == 0) v = ld4(g + ((gr / L) * (L / keep) + i / keep) * ld + c);
    st4(s + r * TS + c, v);
  }
}

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

template <int L>
__device__ __forceinline__ const float* bias_row(const float* e1s, const float* e2s, int i, int j) {
  return (j <= i) ? (e1s + (L - 1 - (i - j)) * TS) : (e2s + (j - i) * TS);
}

// ------------------------------------------------------------------------------------------------------------
template <int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ e1, const float* __restrict__ e2,
                float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                float* __restrict__ probs, int64_t M, int H, int64_t ntiles, int keep) {
  extern __shared__ __align__(16) float smem[];
  float* e1s = smem;
  float* e2s = e1s + L * TS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = e2s + L * TS + warp * (3 * TR * TS);
  float* ks = qs + TR * TS;
  float* vs = ks + TR * TS;
  const int h = blockIdx.y;
  const int d = H * HD;
  for (int i = threadIdx.x; i < L * 16; i += blockDim.x) {
    int r = i / 16, c = (i % 16) * 4;
    st4(e1s + r * TS + c, ld4(e1 + ((size_t)h * L + r) * HD + c));
    st4(e2s + r * TS + c, ld4(e2 + ((size_t)h * L + r) * HD + c));
  }
  __syncthreads();

  const int r = lane >> 1, jh = lane & 1;
  const int g = r / L, i = r % L;
  const bool kept = (i % keep) == 0;       // dead-row elimination: only these query rows produce output
  for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * WARPS) {
    const int64_t row0 = tile * TR;
    __syncwarp();
    load_tile(qs, qkv + h * HD, row0, M, 3 * d, lane);
    load_tile(ks, qkv + d + h * HD, row0, M, 3 * d, lane);
    load_tile(vs, qkv + 2 * d + h * HD, row0, M, 3 * d, lane);
    __syncwarp();

    float s[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) s[jj] = 0.f;
    if (kept)
#pragma unroll 4
    for (int d4 = 0; d4 < 16; ++d4) {
      float4 q4 = ld4(qs + r * TS + d4 * 4);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        int c = jh * 8 + jj;
        if (c / L == g) {
          float4 k4 = ld4(ks + c * TS + d4 * 4);
          float4 e4 = ld4(bias_row<L>(e1s, e2s, i, c % L) + d4 * 4);
          s[jj] += dot4(q4, add4(k4, e4));
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
      if ((jh * 8 + jj) / L == g) mx = fmaxf(mx, s[jj]);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      s[jj] = ((jh * 8 + jj) / L == g) ? expf(s[jj] - mx) : 0.f;
      sum += s[jj];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    const float inv = 1.f / sum;
    const int64_t gr = row0 + r;
    float a_all[16];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float a = s[jj] * inv;
      float o = __shfl_xor_sync(0xffffffffu, a, 1);
      a_all[jj] = jh == 0 ? a : o;     // selects, not dynamic register indexing
      a_all[8 + jj] = jh == 0 ? o : a;
      int c = jh * 8 + jj;
      if (probs != nullptr && kept && gr < M && c / L == g) {
        int64_t p = gr / L;
        probs[((p * H + h) * L + i) * L + (c % L)] = a;
      }
    }
    // O[r][jh*32 .. +32) = sum_c A[r][c] v[c][...]
    float4 o4[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) o4[x] = make_float4(0, 0, 0, 0);
    if (kept)
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float a = a_all[c];
      if (c / L == g) {
#pragma unroll
        for (int x = 0; x < 8; ++x) fma4(o4[x], a, ld4(vs + c * TS + jh * 32 + x * 4));
      }
    }
    if (kept && gr < M) {
      const int64_t orow = (gr / L) * (L / keep) + i / keep;
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const int64_t off = orow * d + h * HD + jh * 32 + x * 4;
        if (out) st4(out + off, o4[x]);
        if (out_hi) store_split4(out_hi + off, out_lo + off, o4[x]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward: given dO, recomputes nothing but reads the saved probabilities.
//   dV = A^T dO ; dA = dO V^T ; dS = A * (dA - rowsum(A*dA)) ; dq~ = dS (K + E) ; dK = dS^T q~ ;
//   dE1[h][L-1-(i-j)] += dS[i][j] q~_i (j<=i) ; dE2[h][j-i] += dS[i][j] q~_i (j>i) ; dq = q_scale * dq~.
template <int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ e1, const float* __restrict__ e2,
                const float* __restrict__ probs, const float* __restrict__ dout, float* __restrict__ dqkv,
                __nv_bfloat16* __restrict__ dqkv_hi, __nv_bfloat16* __restrict__ dqkv_lo,
                float* __restrict__ de1, float* __restrict__ de2, int64_t M, int H, int64_t ntiles, float q_scale, int keep) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PS = 17;   // padded stride of the 16x16 matrices
  constexpr int PER_WARP = 4 * TR * TS + 2 * TR * PS;
  float* e1s = smem;
  float* e2s = e1s + L * TS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = e2s + L * TS + warp * PER_WARP;
  float* ks = qs + TR * TS;
  float* vs = ks + TR * TS;
  float* gs = vs + TR * TS;          // dO tile
  float* As = gs + TR * TS;          // probabilities [16][17] (zero outside the block)
  float* Ds = As + TR * PS;          // dS
  const int h = blockIdx.y;
  const int d = H * HD;
  for (int i = threadIdx.x; i < L * 16; i += blockDim.x) {
    int r = i / 16, c = (i % 16) * 4;
    st4(e1s + r * TS + c, ld4(e1 + ((size_t)h * L + r) * HD + c));
    st4(e2s + r * TS + c, ld4(e2 + ((size_t)h * L + r) * HD + c));
  }
  __syncthreads();

  const int r = lane >> 1, jh = lane & 1;
  const int g = r / L, i = r % L;
  // relative-table gradient accumulators: lane (idx = lane/2, half = lane%2) owns 32 columns of row idx
  float4 acc1[8], acc2[8];
#pragma unroll
  for (int x = 0; x < 8; ++x) { acc1[x] = make_float4(0, 0, 0, 0); acc2[x] = make_float4(0, 0, 0, 0); }

  for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * WARPS) {
    const int64_t row0 = tile * TR;
    const int64_t gr = row0 + r;
    __syncwarp();
    load_tile(qs, qkv + h * HD, row0, M, 3 * d, lane);
    load_tile(ks, qkv + d + h * HD, row0, M, 3 * d, lane);
    load_tile(vs, qkv + 2 * d + h * HD, row0, M, 3 * d, lane);
    if (keep == 1) load_tile(gs, dout + h * HD, row0, M, d, lane);
    else load_tile_compact(gs, dout + h * HD, row0, M, d, lane, L, keep);     // dO exists for the kept query rows only
    const bool kept = (i % keep) == 0;
    float a[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      int c = jh * 8 + jj;
      float v = 0.f;
      if (kept && gr < M && c / L == g) v = probs[(((gr / L) * H + h) * L + i) * L + (c % L)];
      a[jj] = v;
      As[r * PS + c] = v;
    }
    __syncwarp();

    // dA[r][c] = dO[r] . v[c]
    float da[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) da[jj] = 0.f;
#pragma unroll 4
    for (int d4 = 0; d4 < 16; ++d4) {
      float4 g4 = ld4(gs + r * TS + d4 * 4);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        int c = jh * 8 + jj;
        if (c / L == g) da[jj] += dot4(g4, ld4(vs + c * TS + d4 * 4));
      }
    }
    float rd = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) rd += a[jj] * da[jj];
    rd += __shfl_xor_sync(0xffffffffu, rd, 1);
    float ds_all[16];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float v = a[jj] * (da[jj] - rd);
      float o = __shfl_xor_sync(0xffffffffu, v, 1);
      ds_all[jj] = jh == 0 ? v : o;
      ds_all[8 + jj] = jh == 0 ? o : v;
      Ds[r * PS + jh * 8 + jj] = v;
    }
    __syncwarp();

    // dq~[r][jh*32..+32) = sum_c dS[r][c] (k[c] + E(i,j))
    {
      float4 o4[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) o4[x] = make_float4(0, 0, 0, 0);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (c / L == g) {
          float w = ds_all[c];
          const float* er = bias_row<L>(e1s, e2s, i, c % L);
#pragma unroll
          for (int x = 0; x < 8; ++x)
            fma4(o4[x], w, add4(ld4(ks + c * TS + jh * 32 + x * 4), ld4(er + jh * 32 + x * 4)));
        }
      }
      if (gr < M) {
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          float4 v = o4[x];
          v.x *= q_scale; v.y *= q_scale; v.z *= q_scale; v.w *= q_scale;
          const int64_t off = gr * 3 * d + h * HD + jh * 32 + x * 4;
          if (dqkv) st4(dqkv + off, v);
          if (dqkv_hi) store_split4(dqkv_hi + off, dqkv_lo + off, v);
        }
      }
    }
    // dK[c][..] = sum_r dS[r][c] q~[r] ; dV[c][..] = sum_r A[r][c] dO[r]     (lane owns column c = lane/2)
    {
      const int c = r;                 // reuse the (lane/2, lane%2) mapping with the row index read as a column
      float4 k4[8], v4[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) { k4[x] = make_float4(0, 0, 0, 0); v4[x] = make_float4(0, 0, 0, 0); }
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        if (rr / L == g) {             // c's block == rr's block
          float wk = Ds[rr * PS + c], wv = As[rr * PS + c];
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            fma4(k4[x], wk, ld4(qs + rr * TS + jh * 32 + x * 4));
            fma4(v4[x], wv, ld4(gs + rr * TS + jh * 32 + x * 4));
          }
        }
      }
      if (gr < M) {
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const int64_t offk = gr * 3 * d + d + h * HD + jh * 32 + x * 4, offv = offk + d;
          if (dqkv) { st4(dqkv + offk, k4[x]); st4(dqkv + offv, v4[x]); }
          if (dqkv_hi) { store_split4(dqkv_hi + offk, dqkv_lo + offk, k4[x]); store_split4(dqkv_hi + offv, dqkv_lo + offv, v4[x]); }
        }
      }
    }
    // relative tables: lane (idx = lane/2 < L, half)   dE1[idx] += sum_rows dS[row][j = i-(L-1-idx)] q~[row]
    {
      const int idx = r;
      if (idx < L) {
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
          int ii = rr % L, gb = (rr / L) * L;
          int j1 = ii - (L - 1 - idx);          // causal table e1
          int j2 = ii + idx;                    // anticausal table e2 (idx >= 1)
          float w1 = (j1 >= 0) ? Ds[rr * PS + gb + j1] : 0.f;
          float w2 = (idx >= 1 && j2 < L) ? Ds[rr * PS + gb + j2] : 0.f;
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            float4 q4 = ld4(qs + rr * TS + jh * 32 + x * 4);
            fma4(acc1[x], w1, q4);
            fma4(acc2[x], w2, q4);
          }
        }
      }
    }
  }

  // reduce the table gradients over the CTA's warps, then one atomic per element per CTA
  __syncthreads();
  float* red = e2s + L * TS;           // reuse the per-warp staging area: [WARPS][2][L][64]
  if (r < L) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      st4(red + ((warp * 2 + 0) * L + r) * HD + jh * 32 + x * 4, acc1[x]);
      st4(red + ((warp * 2 + 1) * L + r) * HD + jh * 32 + x * 4, acc2[x]);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * L * HD; e += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += red[w * 2 * L * HD + e];
    int which = e / (L * HD), rem = e % (L * HD);
    float* dst = (which == 0 ? de1 : de2) + (size_t)h * L * HD + rem;
    atomicAdd(dst, s);
  }
}

template <int L>
int launch_fwd(const float* qkv, const float* e1, const float* e2, float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
               float* probs, int64_t M, int H, int keep, cudaStream_t st) {
  constexpr int WARPS = 8;
  int64_t ntiles = vq_cdiv(M, TR);
  size_t smem = (size_t)(2 * L * TS + WARPS * 3 * TR * TS) * sizeof(float);
  static bool done = false;
  if (!done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(attn_fwd_kernel<L, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
  }
  int64_t gx = std::min<int64_t>(vq_cdiv(ntiles, WARPS), std::max(1, vq_sm_count() * 2 / H) * 4);
  attn_fwd_kernel<L, WARPS><<<dim3((unsigned)gx, H), WARPS * 32, smem, st>>>(qkv, e1, e2, out, out_hi, out_lo, probs, M, H, ntiles, keep);
  VQ_CUDA_LAUNCH_CHECK("attn_fwd_kernel");
  return VQ_OK;
}

template <int L>
int launch_bwd(const float* qkv, const float* e1, const float* e2, const float* probs, const float* dout, float* dqkv,
               __nv_bfloat16* dqkv_hi, __nv_bfloat16* dqkv_lo, float* de1, float* de2, int64_t M, int H, float q_scale, int keep,
               cudaStream_t st) {
  constexpr int WARPS = 4;
  int64_t ntiles = vq_cdiv(M, TR);
  size_t smem = (size_t)(2 * L * TS + WARPS * (4 * TR * TS + 2 * TR * 17)) * sizeof(float);
  static_assert(WARPS * (4 * TR * TS + 2 * TR * 17) >= WARPS * 2 * 16 * HD, "reduction scratch must fit");
  static bool done = false;
  if (!done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(attn_bwd_kernel<L, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
  }
  int64_t gx = std::min<int64_t>(vq_cdiv(ntiles, WARPS), std::max(1, vq_sm_count() * 2 / H) * 2);
  attn_bwd_kernel<L, WARPS><<<dim3((unsigned)gx, H), WARPS * 32, smem, st>>>(qkv, e1, e2, probs, dout, dqkv, dqkv_hi, dqkv_lo, de1, de2, M, H, ntiles, q_scale, keep);
  VQ_CUDA_LAUNCH_CHECK("attn_bwd_kernel");
  return VQ_OK;
}

}  // namespace

// qkv: (P*L, 3*H*64) with the q third already multiplied by hd^-0.5 (done in the in-proj GEMM epilogue);
// out (fp32) and/or (out_hi, out_lo) (bf16 split, operand of the out-proj GEMM): (P*L/keep, H*64) — only the query rows
// with token % keep == 0 are computed and stored, compacted (keep = 1: all rows; keep = downscale factor in the last
// layer of a stack = exact dead-row elimination); probs: (P, H, L, L) saved for backward (kept rows only; may be NULL).
VQ_API int vqcpc_attn_fwd(const float* qkv, const float* e1, const float* e2, float* out, void* out_hi, void* out_lo,
                          float* probs, int64_t P, int L, int H, int head_dim, int keep, void* stream) {
  VQ_CHECK_ARG(head_dim == HD, "attn_fwd: head_dim must be 64 (got %d)", head_dim);
  VQ_CHECK_ARG(H >= 1 && H <= 64, "attn_fwd: bad head count %d", H);
  VQ_CHECK_ARG(keep >= 1 && L % keep == 0, "attn_fwd: keep=%d must divide L=%d", keep, L);
  VQ_CHECK_ARG((out || out_hi) && ((out_hi == nullptr) == (out_lo == nullptr)), "attn_fwd: need out and/or (out_hi, out_lo)");
  if (P == 0) return VQ_OK;
  cudaStream_t st = vq_stream(stream);
  int64_t M = P * L;
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
  switch (L) {
    case 16: return launch_fwd<16>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, st);
    case 8: return launch_fwd<8>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, st);
    case 4: return launch_fwd<4>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, st);
    case 2: return launch_fwd<2>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, st);
    case 1: return launch_fwd<1>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, st);
    default: VQ_CHECK_ARG(false, "attn_fwd: block length L must be 1, 2, 4, 8 or 16 (got %d)", L);
  }
  return VQ_OK;
}

// d_out: (P*L/keep, H*64) gradient of the kept query rows (see vqcpc_attn_fwd).  d_qkv (fp32, P*L x 3*H*64) and/or its
// bf16 split (d_qkv_hi, d_qkv_lo) are overwritten (q third = gradient wrt the UNscaled projection: q_scale * dq~);
// d_e1 / d_e2 (H*L, 64) are ACCUMULATED.
VQ_API int vqcpc_attn_bwd(const float* qkv, const float* e1, const float* e2, const float* probs, const float* d_out,
                          float* d_qkv, void* d_qkv_hi, void* d_qkv_lo, float* d_e1, float* d_e2, int64_t P, int L, int H,
                          int head_dim, float q_scale, int keep, void* stream) {
  VQ_CHECK_ARG(head_dim == HD, "attn_bwd: head_dim must be 64 (got %d)", head_dim);
  VQ_CHECK_ARG(H >= 1 && H <= 64, "attn_bwd: bad head count %d", H);
  VQ_CHECK_ARG(keep >= 1 && L % keep == 0, "attn_bwd: keep=%d must divide L=%d", keep, L);
  VQ_CHECK_ARG((d_qkv || d_qkv_hi) && ((d_qkv_hi == nullptr) == (d_qkv_lo == nullptr)), "attn_bwd: need d_qkv and/or its split");
  if (P == 0) return VQ_OK;
  cudaStream_t st = vq_stream(stream);
  int64_t M = P * L;
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(d_qkv_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(d_qkv_lo);
  switch (L) {
    case 16: return launch_bwd<16>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, st);
    case 8: return launch_bwd<8>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, st);
    case 4: return launch_bwd<4>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, st);
    case 2: return launch_bwd<2>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, st);
    case 1: return launch_bwd<1>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, st);
    default: VQ_CHECK_ARG(false, "attn_bwd: block length L must be 1, 2, 4, 8 or 16 (got %d)", L);
  }
  return VQ_OK;
}
