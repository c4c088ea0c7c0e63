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
// Two implementations of the same math:
//   * attn_fwd_mma_kernel / attn_bwd_mma_kernel (default): every product on the tensor cores, mma.sync.m16n8k8 TF32 with
//     3xTF32 error compensation (~2^-20 per product), q/k/v/dO tiles staged in shared memory by cp.async with the next
//     tile prefetched behind the current one.  Forward: 78 % of the HBM roofline; backward: 55 % (instruction bound).
//   * attn_fwd_kernel / attn_bwd_kernel: plain fp32 FMAs on the CUDA cores (shared-memory-bandwidth bound, 1.4-2.7x slower),
//     kept for the exact 'fp32' arithmetic mode whose parity tests hold gradients to the fp32 noise floor, and as the
//     reference of the A/B tests (tests/test_gpu_kernels.py::test_attention_*_both_kernels_agree).
// The attention FLOPs are small (0.8 MFLOP per block-layer against ~100 MFLOP of GEMMs); what matters is moving 8.7 KB
// (forward) / 14.8 KB (backward) per token through HBM once.
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

// fp32 x4 -> bf16 (hi, lo) x4 with x ~= hi + lo, two values per conversion instruction (cvt.rn.bf16x2.f32)
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);                     // .x (low half) = x, .y = y
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - __uint_as_float(hi << 16), y - __uint_as_float(hi & 0xffff0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, float4 v) {
  uint2 hv, lv;
  split_pair(v.x, v.y, hv.x, lv.x);
  split_pair(v.z, v.w, hv.y, lv.y);
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
                float* __restrict__ probs, int64_t M, int H, int64_t ntiles, int keep,
                const unsigned char* __restrict__ drop_keep, float drop_scale) {
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
      int c = jh * 8 + jj;
      if (kept && gr < M && c / L == g) {
        const int64_t pi = (((gr / L) * H + h) * L + i) * L + (c % L);
        if (probs != nullptr) probs[pi] = a;                 // the softmax output; dropout (F.dropout on the weights,
        if (drop_keep != nullptr) a = drop_keep[pi] ? a * drop_scale : 0.f;   // multihead_attention_custom.py:339) acts on PV only
      }
      float o = __shfl_xor_sync(0xffffffffu, a, 1);
      a_all[jj] = jh == 0 ? a : o;     // selects, not dynamic register indexing
      a_all[8 + jj] = jh == 0 ? o : a;
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
// Forward on the tensor cores: mma.sync.m16n8k8 TF32 with the 3xTF32 error compensation (x = big + small, both TF32;
// D += a_small b_big + a_big b_small + a_big b_big: ~2^-20 relative to |a||b| per product, 16x coarser than fp32 FMA;
// the exact-fp32 GEMM mode therefore keeps the CUDA-core kernel).  The CUDA-core kernel above is
// bound by shared-memory bandwidth (17 LDS.128 per 8 dot products of the fused q.(k + E) form, ncu: l1tex 84 %); here
// the skewed bias is separated into two small matrix products over the relative tables,
//     S = Q K^T,   R1 = Q E1^T,   R2 = Q E2^T   (16 x 16, 16 x L, 16 x L),   S[i][j] += (j <= i ? R1[i][L-1-(i-j)] : R2[i][j-i])
// fragments are read with conflict-free LDS.32 (row stride 68 floats), the Q fragment of a k-step is split once and
// reused by all three products, and P V is a fourth product after the softmax.  ~8x less shared-memory traffic per tile.
constexpr int PSM = 20;   // padded stride of the 16 x 16 scratch matrices
static_assert(8 * 132 <= TR * TS, "a pair-interleaved table fits in a [16][TS] slot");

// x = big + small with big = x rounded to TF32 (10 explicit mantissa bits, nearest, ties away: one integer add + mask on
// the bit pattern; inf stays inf) and small = x - big (exact in fp32).  The tensor core ignores the low 13 mantissa bits of
// a .tf32 operand, so small is used as is: |x - big - tf32(small)| <= 2^-21 |x|.  (cvt.rna.tf32.f32 is emulated with ~5
// integer instructions on sm_100; at 336 conversions per tile it was 60 % of the kernel's instructions.)
__device__ __forceinline__ void tf32_split(float x, uint32_t& big, uint32_t& small) {
  big = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  small = __float_as_uint(x - __uint_as_float(big));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += A B with A given as (big, small) fragments and B as fp32 values split here
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ab)[4], const uint32_t (&as)[4], float b0f, float b1f) {
  uint32_t b0b, b0s, b1b, b1s;
  tf32_split(b0f, b0b, b0s);
  tf32_split(b1f, b1b, b1s);
  mma_tf32(c, as, b0b, b1b);
  mma_tf32(c, ab, b0s, b1s);
  mma_tf32(c, ab, b0b, b1b);
}

// 16-byte asynchronous global -> shared copy (LDGSTS); !valid zero-fills the destination without reading
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_but() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }   // all but the N newest groups

// asynchronous versions of load_tile / load_tile_compact (same layout, rows >= M or dropped rows are zero)
__device__ __forceinline__ void load_tile_async(float* s, const float* __restrict__ g, int64_t row0, int64_t M, int64_t ld, int lane) {
  const int r0 = lane >> 4, c = (lane & 15) * 4;
  const float* p = g + (row0 + r0) * ld + c;       // one 64-bit multiply per tile, then pointer steps
  float* sp = s + r0 * TS + c;
  const int64_t left = M - row0 - r0;              // row r0 + 2 it exists while 2 it < left
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const bool ok = 2 * it < left;
    cp_async16(sp + it * 2 * TS, ok ? p : g, ok);
    p += 2 * ld;
  }
}
// (keep is a power of two - it divides L in {1, 2, 4, 8, 16} - so the kernels work with kmask = keep - 1 and
// kshift = log2(keep): no integer divisions in the tile loops)
__device__ __forceinline__ void load_tile_compact_async(float* s, const float* __restrict__ g, int64_t row0, int64_t M, int64_t ld,
                                                        int lane, int L, int kmask, int kshift) {
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 2 + (lane >> 4), c = (lane & 15) * 4;
    const int64_t gr = row0 + r;
    const int i = (int)(gr % L);
    const bool ok = gr < M && (i & kmask) == 0;
    cp_async16(s + r * TS + c, ok ? g + ((gr / L) * (L >> kshift) + (i >> kshift)) * ld + c : g, ok);
  }
}

// B fragments of one k-step for all eight n-tiles in the "features are output columns" layout: column g of n-tile n is the
// physical feature 8g + n, so rows ra / rb (the k indices t and t + 4 of the step) of a [16][TS] tile give all eight
// n-tiles with two LDS.128 each; bb/bs[n] = split of row ra, cb/cs[n] = split of row rb.
__device__ __forceinline__ void load_bfrag_n(const float* __restrict__ Y, int ra, int rb, int g, uint32_t (&bb)[8], uint32_t (&bs)[8],
                                             uint32_t (&cb)[8], uint32_t (&cs)[8]) {
  const float4 x0 = ld4(Y + ra * TS + 8 * g), x1 = ld4(Y + ra * TS + 8 * g + 4);
  const float4 y0 = ld4(Y + rb * TS + 8 * g), y1 = ld4(Y + rb * TS + 8 * g + 4);
  const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    tf32_split(xv[n], bb[n], bs[n]);
    tf32_split(yv[n], cb[n], cs[n]);
  }
}
// Relative tables as B operand, split to TF32 once per CTA and stored pair-interleaved: row pair p = kk * 4 + t holds the
// table rows (kk*8 + t, kk*8 + t + 4) - the k indices t and t + 4 of k-step kk - as [feature][2], row-pair stride TPS.
// One LDS.128 then returns (b0, b1) of two n-tiles in adjacent registers, exactly what the MMA wants (loading the two rows
// separately cost two register moves per MMA); TPS = 132 makes the quarter-warp phases bank-conflict free.
constexpr int TPS = 132;
__device__ __forceinline__ int tab_index(int r, int f) { return ((r >> 3) * 4 + (r & 3)) * TPS + 2 * f + ((r >> 2) & 1); }
__device__ __forceinline__ void load_bfrag_n_tab(const float* __restrict__ Tb, const float* __restrict__ Tl, int kk, int t, int g,
                                                 uint32_t (&bb)[8], uint32_t (&bs)[8], uint32_t (&cb)[8], uint32_t (&cs)[8]) {
  const float* pb = Tb + (kk * 4 + t) * TPS + 16 * g;
  const float* pl = Tl + (kk * 4 + t) * TPS + 16 * g;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 x = *reinterpret_cast<const uint4*>(pb + 4 * i);
    const uint4 u = *reinterpret_cast<const uint4*>(pl + 4 * i);
    bb[2 * i] = x.x; cb[2 * i] = x.y; bb[2 * i + 1] = x.z; cb[2 * i + 1] = x.w;
    bs[2 * i] = u.x; cs[2 * i] = u.y; bs[2 * i + 1] = u.z; cs[2 * i + 1] = u.w;
  }
}
__device__ __forceinline__ void mma_3x_pre(float (&c)[4], const uint32_t (&ab)[4], const uint32_t (&as)[4], uint32_t b0b,
                                           uint32_t b0s, uint32_t b1b, uint32_t b1s) {
  mma_tf32(c, as, b0b, b1b);
  mma_tf32(c, ab, b0s, b1s);
  mma_tf32(c, ab, b0b, b1b);
}

// Index permutations keep every shared-memory access a 128-bit one:
//   * where the 64 features are contracted (S, R1, R2), the k-step kk = 2j + s pairs the logical columns (t, t + 4) with
//     the physical features 16j + 4t + 2s, + 1 in BOTH operands, so one LDS.128 feeds two k-steps;
//   * in P V the two k-steps contract the keys (n*8 + 2t, n*8 + 2t + 1), which are exactly the accumulator entries of S this
//     lane already holds: the probabilities never leave the registers;
//   * the output columns of P V are permuted (column g of n-tile n = feature 8g + n): all eight B fragments of a k-step are
//     two LDS.128 and a lane ends up with 16 consecutive output features per row (64-byte global stores).
// The next tile's k and q are fetched with cp.async as soon as the scores are final (they overlap softmax, P V and the
// stores), v right after P V.
template <int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
attn_fwd_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ e1, const float* __restrict__ e2,
                    float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                    float* __restrict__ probs, int64_t M, int H, int64_t ntiles, int keep,
                    const unsigned char* __restrict__ drop_keep, float drop_scale) {
  constexpr int NTR = (L + 7) / 8;                 // n-tiles of the relative-table products
  extern __shared__ __align__(16) float smem[];
  float* e1s = smem;                               // [16][TS], rows >= L zero
  float* e2s = e1s + TR * TS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = e2s + TR * TS + warp * (3 * TR * TS);
  float* ks = qs + TR * TS;
  float* vs = ks + TR * TS;
  float* r1s = qs;                                 // after the three Q products the q tile is scratch for R1, R2 (16 x 16):
  float* r2s = qs + TR * PSM;                      //   q . e1 / q . e2 before the skew
  const int h = blockIdx.y;
  const int d = H * HD;
  const int kmask = keep - 1, kshift = __ffs(keep) - 1;
  const int64_t tile0 = (int64_t)blockIdx.x * WARPS + warp, tstride = (int64_t)gridDim.x * WARPS;
  // cp.async groups are committed in the order [q + k], [v] per tile, and each is awaited only where its tile is first
  // read ("all but the newest group"), so the v prefetch issued at the very end of a tile overlaps the next tile's
  // score products instead of being waited for at its top.
  if (tile0 < ntiles) {
    load_tile_async(qs, qkv + h * HD, tile0 * TR, M, 3 * d, lane);
    load_tile_async(ks, qkv + d + h * HD, tile0 * TR, M, 3 * d, lane);
  }
  cp_async_commit();
  if (tile0 < ntiles) load_tile_async(vs, qkv + 2 * d + h * HD, tile0 * TR, M, 3 * d, lane);
  cp_async_commit();
  for (int i = threadIdx.x; i < TR * 16; i += blockDim.x) {
    int r = i / 16, c = (i % 16) * 4;
    float4 a = make_float4(0, 0, 0, 0), b = a;
    if (r < L) {
      a = ld4(e1 + ((size_t)h * L + r) * HD + c);
      b = ld4(e2 + ((size_t)h * L + r) * HD + c);
    }
    st4(e1s + r * TS + c, a);
    st4(e2s + r * TS + c, b);
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;           // mma fragment coordinates
  for (int64_t tile = tile0; tile < ntiles; tile += tstride) {
    const int64_t row0 = tile * TR;
    const int64_t nrow0 = (tile + tstride) * TR;
    const bool has_next = tile + tstride < ntiles;
    cp_async_wait_but<1>();                         // q and k of this tile have landed (v may still be in flight)
    __syncwarp();
    float sc[2][4], ra[NTR][4], rb[NTR][4];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) sc[n][x] = 0.f;
#pragma unroll
    for (int n = 0; n < NTR; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) { ra[n][x] = 0.f; rb[n][x] = 0.f; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 q0 = ld4(qs + g * TS + 16 * j + 4 * t), q1 = ld4(qs + (g + 8) * TS + 16 * j + 4 * t);
      float4 kf[2], ef1[NTR], ef2[NTR];
#pragma unroll
      for (int n = 0; n < 2; ++n) kf[n] = ld4(ks + (n * 8 + g) * TS + 16 * j + 4 * t);
#pragma unroll
      for (int n = 0; n < NTR; ++n) {
        ef1[n] = ld4(e1s + (n * 8 + g) * TS + 16 * j + 4 * t);
        ef2[n] = ld4(e2s + (n * 8 + g) * TS + 16 * j + 4 * t);
      }
      uint32_t qb[4], qsm[4];
      // k-step 2j: physical features 16j + 4t + {0, 1}
      tf32_split(q0.x, qb[0], qsm[0]); tf32_split(q1.x, qb[1], qsm[1]); tf32_split(q0.y, qb[2], qsm[2]); tf32_split(q1.y, qb[3], qsm[3]);
#pragma unroll
      for (int n = 0; n < 2; ++n) mma_3x(sc[n], qb, qsm, kf[n].x, kf[n].y);
#pragma unroll
      for (int n = 0; n < NTR; ++n) {
        mma_3x(ra[n], qb, qsm, ef1[n].x, ef1[n].y);
        mma_3x(rb[n], qb, qsm, ef2[n].x, ef2[n].y);
      }
      // k-step 2j + 1: physical features 16j + 4t + {2, 3}
      tf32_split(q0.z, qb[0], qsm[0]); tf32_split(q1.z, qb[1], qsm[1]); tf32_split(q0.w, qb[2], qsm[2]); tf32_split(q1.w, qb[3], qsm[3]);
#pragma unroll
      for (int n = 0; n < 2; ++n) mma_3x(sc[n], qb, qsm, kf[n].z, kf[n].w);
#pragma unroll
      for (int n = 0; n < NTR; ++n) {
        mma_3x(ra[n], qb, qsm, ef1[n].z, ef1[n].w);
        mma_3x(rb[n], qb, qsm, ef2[n].z, ef2[n].w);
      }
    }
    __syncwarp();                                   // every lane is done with q and k: prefetch the next tile's k
    if (has_next) load_tile_async(ks, qkv + d + h * HD, nrow0, M, 3 * d, lane);
    // ---- R1 / R2 to scratch (C layout: rows g, g + 8; cols n*8 + 2t, +1)
#pragma unroll
    for (int n = 0; n < NTR; ++n) {
      *reinterpret_cast<float2*>(r1s + g * PSM + n * 8 + 2 * t) = make_float2(ra[n][0], ra[n][1]);
      *reinterpret_cast<float2*>(r1s + (g + 8) * PSM + n * 8 + 2 * t) = make_float2(ra[n][2], ra[n][3]);
      *reinterpret_cast<float2*>(r2s + g * PSM + n * 8 + 2 * t) = make_float2(rb[n][0], rb[n][1]);
      *reinterpret_cast<float2*>(r2s + (g + 8) * PSM + n * 8 + 2 * t) = make_float2(rb[n][2], rb[n][3]);
    }
    __syncwarp();
    // ---- bias + block mask + softmax; this lane holds rows g and g + 8, columns n*8 + 2t + {0, 1}
    float pv[2][4];                                 // [n][c0 c1 (row g), c2 c3 (row g + 8)]
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int r = g + (x >> 1) * 8, c = n * 8 + 2 * t + (x & 1);
        const int i = r % L, j = c % L;
        float v = -INFINITY;
        if (r / L == c / L) v = sc[n][x] + ((j <= i) ? r1s[r * PSM + (L - 1 - (i - j))] : r2s[r * PSM + (j - i)]);
        pv[n][x] = v;
        mx[x >> 1] = fmaxf(mx[x >> 1], v);
      }
    __syncwarp();                                   // the scratch (= q tile) is free: prefetch the next tile's q
    if (has_next) load_tile_async(qs, qkv + h * HD, nrow0, M, 3 * d, lane);
    cp_async_commit();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 1));
      mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 2));
    }
    float sm[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        pv[n][x] = expf(pv[n][x] - mx[x >> 1]);     // exp(-inf) = 0 outside the block
        sm[x >> 1] += pv[n][x];
      }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      sm[rr] += __shfl_xor_sync(0xffffffffu, sm[rr], 1);
      sm[rr] += __shfl_xor_sync(0xffffffffu, sm[rr], 2);
      sm[rr] = 1.f / sm[rr];
    }
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int r = g + (x >> 1) * 8, c = n * 8 + 2 * t + (x & 1);
        float a = pv[n][x] * sm[x >> 1];
        const int64_t gr = row0 + r;
        const int i = r % L;
        if (gr < M && (i & kmask) == 0 && r / L == c / L) {
          const int64_t pi = (((gr / L) * H + h) * L + i) * L + (c % L);
          if (probs != nullptr) probs[pi] = a;
          if (drop_keep != nullptr) a = drop_keep[pi] ? a * drop_scale : 0.f;
        }
        pv[n][x] = a;
      }
    cp_async_wait_but<1>();                         // v of this tile has landed (the next q + k may still be in flight)
    __syncwarp();
    // ---- O = P V (16 x 64), keys contracted in the order this lane's S entries have them
    float oc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) oc[n][x] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t pb[4], psm_[4], bb[8], bs[8], cb[8], cs[8];
      tf32_split(pv[kk][0], pb[0], psm_[0]);        // (row g,     key kk*8 + 2t)
      tf32_split(pv[kk][2], pb[1], psm_[1]);        // (row g + 8, key kk*8 + 2t)
      tf32_split(pv[kk][1], pb[2], psm_[2]);        // (row g,     key kk*8 + 2t + 1)
      tf32_split(pv[kk][3], pb[3], psm_[3]);        // (row g + 8, key kk*8 + 2t + 1)
      load_bfrag_n(vs, kk * 8 + 2 * t, kk * 8 + 2 * t + 1, g, bb, bs, cb, cs);
#pragma unroll
      for (int n = 0; n < 8; ++n) mma_3x_pre(oc[n], pb, psm_, bb[n], bs[n], cb[n], cs[n]);
    }
    __syncwarp();                                   // v is free: prefetch the next tile's
    if (has_next) load_tile_async(vs, qkv + 2 * d + h * HD, nrow0, M, 3 * d, lane);
    cp_async_commit();
    // ---- store rows g and g + 8 (kept query rows only): features 16t .. 16t + 15 (16t + n = oc[n][0|2], 16t + 8 + n = oc[n][1|3])
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = g + rr * 8;
      const int64_t gr = row0 + r;
      const int i = r % L;
      if (gr < M && (i & kmask) == 0) {
        const int64_t off = ((gr / L) * (L >> kshift) + (i >> kshift)) * d + h * HD + 16 * t;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int x = rr * 2 + (q >> 1), n0 = (q & 1) * 4;
          const float4 v = make_float4(oc[n0][x], oc[n0 + 1][x], oc[n0 + 2][x], oc[n0 + 3][x]);
          if (out) st4(out + off + 4 * q, v);
          if (out_hi) store_split4(out_hi + off + 4 * q, out_lo + off + 4 * q, v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward: reads the saved probabilities A (nothing is recomputed).
//   dA = dO V^T ; dS = A * (dA - rowsum(A*dA)) ; dQ~ = dS K + T1 e1 + T2 e2 ; dK = dS^T Q~ ; dV = A^T dO ;
//   dE1 = T1^T Q~ ; dE2 = T2^T Q~ ; dq = q_scale * dq~
// where T1 / T2 are dS "skewed" onto the relative index: T1[i][L-1-(i-j)] = dS[i][j] (j <= i), T2[i][j-i] = dS[i][j]
// (j > i).  With that, every product is a small dense matmul over the 16-row tile, evaluated with register-blocked
// FMAs: 16x16 results use a (2 rows x 4 cols) tile per lane, 16x64 results a (4 rows x 8 cols) tile per lane with the
// reduction index vectorised by 4 -> ~10 FMA per shared-memory load instead of ~4.
constexpr int PS = 20;   // padded stride of the 16x16 matrices (rows 16-byte aligned)

// acc[4][8] += sum_k Msm[row(rg,ri)][k] * Bsm[k][dg*8 .. +8]   for k in [0, 16): "rows x reduction" coefficient matrix
__device__ __forceinline__ void mm_rows(float (&acc)[4][8], const float* __restrict__ Msm, const float* __restrict__ Bsm,
                                        int rg, int dg) {
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) {
    float4 m[4];
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) m[ri] = ld4(Msm + (rg * 4 + ri) * PS + k4 * 4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float4 b0 = ld4(Bsm + (k4 * 4 + kk) * TS + dg * 8);
      float4 b1 = ld4(Bsm + (k4 * 4 + kk) * TS + dg * 8 + 4);
#pragma unroll
      for (int ri = 0; ri < 4; ++ri) {
        float w = kk == 0 ? m[ri].x : (kk == 1 ? m[ri].y : (kk == 2 ? m[ri].z : m[ri].w));
        acc[ri][0] = fmaf(w, b0.x, acc[ri][0]); acc[ri][1] = fmaf(w, b0.y, acc[ri][1]);
        acc[ri][2] = fmaf(w, b0.z, acc[ri][2]); acc[ri][3] = fmaf(w, b0.w, acc[ri][3]);
        acc[ri][4] = fmaf(w, b1.x, acc[ri][4]); acc[ri][5] = fmaf(w, b1.y, acc[ri][5]);
        acc[ri][6] = fmaf(w, b1.z, acc[ri][6]); acc[ri][7] = fmaf(w, b1.w, acc[ri][7]);
      }
    }
  }
}

// acc[4][8] += sum_r Msm[r][cg*4 + ci] * Bsm[r][dg*8 .. +8]   for r in [0, 16): transposed coefficient matrix
__device__ __forceinline__ void mm_cols(float (&acc)[4][8], const float* __restrict__ Msm, const float* __restrict__ Bsm,
                                        int cg, int dg) {
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    float4 m = ld4(Msm + r * PS + cg * 4);
    float4 b0 = ld4(Bsm + r * TS + dg * 8);
    float4 b1 = ld4(Bsm + r * TS + dg * 8 + 4);
    const float w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      acc[ci][0] = fmaf(w[ci], b0.x, acc[ci][0]); acc[ci][1] = fmaf(w[ci], b0.y, acc[ci][1]);
      acc[ci][2] = fmaf(w[ci], b0.z, acc[ci][2]); acc[ci][3] = fmaf(w[ci], b0.w, acc[ci][3]);
      acc[ci][4] = fmaf(w[ci], b1.x, acc[ci][4]); acc[ci][5] = fmaf(w[ci], b1.y, acc[ci][5]);
      acc[ci][6] = fmaf(w[ci], b1.z, acc[ci][6]); acc[ci][7] = fmaf(w[ci], b1.w, acc[ci][7]);
    }
  }
}

template <int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ e1, const float* __restrict__ e2,
                const float* __restrict__ probs, const float* __restrict__ dout, float* __restrict__ dqkv,
                __nv_bfloat16* __restrict__ dqkv_hi, __nv_bfloat16* __restrict__ dqkv_lo,
                float* __restrict__ de1, float* __restrict__ de2, int64_t M, int H, int64_t ntiles, float q_scale, int keep,
                const unsigned char* __restrict__ drop_keep, float drop_scale) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PER_WARP = 4 * TR * TS + 4 * TR * PS;
  float* e1s = smem;                     // [16][TS], rows >= L are zero
  float* e2s = e1s + TR * TS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = e2s + TR * TS + warp * PER_WARP;
  float* ks = qs + TR * TS;
  float* vs = ks + TR * TS;
  float* gs = vs + TR * TS;              // dO tile
  float* As = gs + TR * TS;              // probabilities [16][PS] (zero outside the block / for dropped rows)
  float* Ds = As + TR * PS;              // dS
  float* T1 = Ds + TR * PS;              // dS skewed onto the causal table index
  float* T2 = T1 + TR * PS;              // ... anticausal
  const int h = blockIdx.y;
  const int d = H * HD;
  for (int i = threadIdx.x; i < TR * 16; i += blockDim.x) {
    int r = i / 16, c = (i % 16) * 4;
    float4 a = make_float4(0, 0, 0, 0), b = a;
    if (r < L) {
      a = ld4(e1 + ((size_t)h * L + r) * HD + c);
      b = ld4(e2 + ((size_t)h * L + r) * HD + c);
    }
    st4(e1s + r * TS + c, a);
    st4(e2s + r * TS + c, b);
  }
  __syncthreads();

  // lane tilings
  const int r2 = (lane >> 2) * 2, c4 = (lane & 3) * 4;      // 16x16 products: rows r2, r2+1; cols c4..c4+3
  const int rg = lane >> 3, dg = lane & 7;                   // 16x64 products: rows rg*4..+3; cols dg*8..+7
  const int sr = lane >> 1, sh = lane & 1;                   // skew step: row sr, table indices sh*8..+7
  float acc_e1[4][8], acc_e2[4][8];                          // dE1 / dE2 rows rg*4..+3 (table index), cols dg*8..+7
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) { acc_e1[a][b] = 0.f; acc_e2[a][b] = 0.f; }

  for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile < ntiles; tile += (int64_t)gridDim.x * WARPS) {
    const int64_t row0 = tile * TR;
    __syncwarp();
    load_tile(qs, qkv + h * HD, row0, M, 3 * d, lane);
    load_tile(ks, qkv + d + h * HD, row0, M, 3 * d, lane);
    load_tile(vs, qkv + 2 * d + h * HD, row0, M, 3 * d, lane);
    if (keep == 1) load_tile(gs, dout + h * HD, row0, M, d, lane);
    else load_tile_compact(gs, dout + h * HD, row0, M, d, lane, L, keep);     // dO exists for the kept query rows only
    // probabilities of this lane's 2x4 patch (and, with dropout, the factor m = keep * scale of each: the forward used
    // A' = A * m in O = A' V, so dV = A'^T dO and dA = (dO V^T) * m, while the softmax backward uses A itself)
    float a[2][4], mk[2][4];
#pragma unroll
    for (int ri = 0; ri < 2; ++ri) {
      const int r = r2 + ri;
      const int64_t gr = row0 + r;
      const int i = r % L;
      const bool rowok = gr < M && (i % keep) == 0;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c = c4 + ci;
        float v = 0.f, m = 1.f;
        if (rowok && c / L == r / L) {
          const int64_t pi = (((gr / L) * H + h) * L + i) * L + (c % L);
          v = probs[pi];
          if (drop_keep != nullptr) m = drop_keep[pi] ? drop_scale : 0.f;
        }
        a[ri][ci] = v;
        mk[ri][ci] = m;
      }
      st4(As + r * PS + c4, make_float4(a[ri][0] * mk[ri][0], a[ri][1] * mk[ri][1], a[ri][2] * mk[ri][2], a[ri][3] * mk[ri][3]));
    }
    __syncwarp();

    // ---- dA = dO V^T on the 2x4 patch, then dS
    float da[2][4];
#pragma unroll
    for (int ri = 0; ri < 2; ++ri)
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) da[ri][ci] = 0.f;
#pragma unroll 4
    for (int d4 = 0; d4 < 16; ++d4) {
      float4 g0 = ld4(gs + r2 * TS + d4 * 4), g1 = ld4(gs + (r2 + 1) * TS + d4 * 4);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        float4 v4 = ld4(vs + (c4 + ci) * TS + d4 * 4);
        da[0][ci] += dot4(g0, v4);
        da[1][ci] += dot4(g1, v4);
      }
    }
#pragma unroll
    for (int ri = 0; ri < 2; ++ri) {
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) da[ri][ci] *= mk[ri][ci];
      float rd = a[ri][0] * da[ri][0] + a[ri][1] * da[ri][1] + a[ri][2] * da[ri][2] + a[ri][3] * da[ri][3];
      rd += __shfl_xor_sync(0xffffffffu, rd, 1);
      rd += __shfl_xor_sync(0xffffffffu, rd, 2);
      st4(Ds + (r2 + ri) * PS + c4, make_float4(a[ri][0] * (da[ri][0] - rd), a[ri][1] * (da[ri][1] - rd),
                                                a[ri][2] * (da[ri][2] - rd), a[ri][3] * (da[ri][3] - rd)));
    }
    __syncwarp();
    // ---- skew dS onto the relative-table index
    {
      const int i = sr % L, gb = (sr / L) * L;
      float t1[8], t2[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const int idx = sh * 8 + x;
        const int j1 = i - (L - 1 - idx);            // e1 row idx <-> key j = i - (L-1-idx) <= i
        const int j2 = i + idx;                      // e2 row idx (>= 1) <-> key j = i + idx > i
        t1[x] = (idx < L && j1 >= 0) ? Ds[sr * PS + gb + j1] : 0.f;
        t2[x] = (idx >= 1 && idx < L && j2 < L) ? Ds[sr * PS + gb + j2] : 0.f;
      }
      st4(T1 + sr * PS + sh * 8, make_float4(t1[0], t1[1], t1[2], t1[3]));
      st4(T1 + sr * PS + sh * 8 + 4, make_float4(t1[4], t1[5], t1[6], t1[7]));
      st4(T2 + sr * PS + sh * 8, make_float4(t2[0], t2[1], t2[2], t2[3]));
      st4(T2 + sr * PS + sh * 8 + 4, make_float4(t2[4], t2[5], t2[6], t2[7]));
    }
    __syncwarp();

    float acc[4][8];
    // ---- dQ~ = dS K + T1 e1 + T2 e2   (rows rg*4..+3)
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int y = 0; y < 8; ++y) acc[x][y] = 0.f;
    mm_rows(acc, Ds, ks, rg, dg);
    mm_rows(acc, T1, e1s, rg, dg);
    mm_rows(acc, T2, e2s, rg, dg);
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
      const int64_t gr = row0 + rg * 4 + ri;
      if (gr < M) {
        const int64_t off = gr * 3 * d + h * HD + dg * 8;
        float4 o0 = make_float4(acc[ri][0] * q_scale, acc[ri][1] * q_scale, acc[ri][2] * q_scale, acc[ri][3] * q_scale);
        float4 o1 = make_float4(acc[ri][4] * q_scale, acc[ri][5] * q_scale, acc[ri][6] * q_scale, acc[ri][7] * q_scale);
        if (dqkv) { st4(dqkv + off, o0); st4(dqkv + off + 4, o1); }
        if (dqkv_hi) { store_split4(dqkv_hi + off, dqkv_lo + off, o0); store_split4(dqkv_hi + off + 4, dqkv_lo + off + 4, o1); }
      }
    }
    // ---- dK = dS^T Q~ ; dV = A^T dO   (rows = key index rg*4..+3)
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 8; ++y) acc[x][y] = 0.f;
      if (pass == 0) mm_cols(acc, Ds, qs, rg, dg);
      else mm_cols(acc, As, gs, rg, dg);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int64_t gr = row0 + rg * 4 + ci;
        if (gr < M) {
          const int64_t off = gr * 3 * d + (pass + 1) * d + h * HD + dg * 8;
          float4 o0 = make_float4(acc[ci][0], acc[ci][1], acc[ci][2], acc[ci][3]);
          float4 o1 = make_float4(acc[ci][4], acc[ci][5], acc[ci][6], acc[ci][7]);
          if (dqkv) { st4(dqkv + off, o0); st4(dqkv + off + 4, o1); }
          if (dqkv_hi) { store_split4(dqkv_hi + off, dqkv_lo + off, o0); store_split4(dqkv_hi + off + 4, dqkv_lo + off + 4, o1); }
        }
      }
    }
    // ---- relative tables: dE1 += T1^T Q~ ; dE2 += T2^T Q~   (rows = table index rg*4..+3; only < L are non-zero)
    if (rg * 4 < L) {
      mm_cols(acc_e1, T1, qs, rg, dg);
      mm_cols(acc_e2, T2, qs, rg, dg);
    }
  }

  // reduce the table gradients over the CTA's warps, then one atomic per element per CTA
  __syncthreads();
  float* red = e2s + TR * TS;           // reuse the per-warp staging area: [WARPS][2][16][64]
#pragma unroll
  for (int ci = 0; ci < 4; ++ci) {
    const int idx = rg * 4 + ci;
    st4(red + ((warp * 2 + 0) * TR + idx) * HD + dg * 8, make_float4(acc_e1[ci][0], acc_e1[ci][1], acc_e1[ci][2], acc_e1[ci][3]));
    st4(red + ((warp * 2 + 0) * TR + idx) * HD + dg * 8 + 4, make_float4(acc_e1[ci][4], acc_e1[ci][5], acc_e1[ci][6], acc_e1[ci][7]));
    st4(red + ((warp * 2 + 1) * TR + idx) * HD + dg * 8, make_float4(acc_e2[ci][0], acc_e2[ci][1], acc_e2[ci][2], acc_e2[ci][3]));
    st4(red + ((warp * 2 + 1) * TR + idx) * HD + dg * 8 + 4, make_float4(acc_e2[ci][4], acc_e2[ci][5], acc_e2[ci][6], acc_e2[ci][7]));
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * L * HD; e += blockDim.x) {
    const int which = e / (L * HD), rem = e % (L * HD);       // rem = idx * 64 + col, idx < L
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += red[(w * 2 + which) * TR * HD + rem];
    float* dst = (which == 0 ? de1 : de2) + (size_t)h * L * HD + rem;
    atomicAdd(dst, s);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Backward on the tensor cores (mma.sync.m16n8k8 3xTF32, see the forward): the same seven products as above, as 384 MMAs
// per (tile, head) instead of ~4100 FMAs + ~400 LDS.128 per lane.
//   (1) dA = dO V^T            features are the contraction index
//   (2) dS = A * (dA - rowsum(A * dA)) on the accumulator fragment; A', dS, T1, T2 (16 x 16) go to shared memory
//   (3) dQ~ = dS K + T1 e1 + T2 e2          (4) dK = dS^T Q~      (5) dV = A'^T dO      (6) dE1 += T1^T Q~, dE2 += T2^T Q~
// Index permutations keep every shared-memory access a 128-bit one:
//   * where the 64 features are contracted (1), the k-step kk = 2j + s pairs logical columns (t, t + 4) with the physical
//     features 16j + 4t + 2s, + 1 in BOTH operands, so one LDS.128 feeds two k-steps;
//   * where the features are the output columns (3)-(6), column g of n-tile n is the physical feature 8g + n, so the B
//     fragments of all eight n-tiles are two LDS.128 and a lane ends up with 16 consecutive output features per row
//     (64-byte global stores).
// (4) and (6) share the split Q~ fragments.  The relative tables' gradients stay in registers for the whole kernel.
// Measured and rejected (34 816 blocks x 16 tokens, on the 2.65 ms version): a warp-PAIR per tile sharing the staging buffers (128
// registers, 16 warps / SM instead of 8: 2.78 ms) and fully unrolled k-loops (2.75 ms); awaiting each cp.async group only
// where its tile is first read instead of at the top of the tile changes nothing either - the kernel is bound by its
// ~2900 dependent-issue-limited instructions per tile (issue slots 27 % busy, no memory or pipe above 60 %), not by load
// latency or occupancy.
// A fragment (split) of a row-major [16][PSM] matrix, k-step kk
__device__ __forceinline__ void load_afrag(const float* __restrict__ X, int kk, int g, int t, uint32_t (&ab)[4], uint32_t (&as)[4]) {
  tf32_split(X[g * PSM + kk * 8 + t], ab[0], as[0]);
  tf32_split(X[(g + 8) * PSM + kk * 8 + t], ab[1], as[1]);
  tf32_split(X[g * PSM + kk * 8 + t + 4], ab[2], as[2]);
  tf32_split(X[(g + 8) * PSM + kk * 8 + t + 4], ab[3], as[3]);
}
// A fragment of the TRANSPOSE of a row-major [16][PSM] matrix: A[m][k] = X[k][m]
__device__ __forceinline__ void load_afrag_t(const float* __restrict__ X, int kk, int g, int t, uint32_t (&ab)[4], uint32_t (&as)[4]) {
  tf32_split(X[(kk * 8 + t) * PSM + g], ab[0], as[0]);
  tf32_split(X[(kk * 8 + t) * PSM + g + 8], ab[1], as[1]);
  tf32_split(X[(kk * 8 + t + 4) * PSM + g], ab[2], as[2]);
  tf32_split(X[(kk * 8 + t + 4) * PSM + g + 8], ab[3], as[3]);
}

// store a 16 x 64 accumulator in the permuted layout: this lane holds, for rows g and g + 8, the 16 consecutive features
// 16t .. 16t + 15 (feature 16t + n = acc[n][0 | 2], feature 16t + 8 + n = acc[n][1 | 3])
__device__ __forceinline__ void store_acc_n(const float (&acc)[8][4], float scale, float* __restrict__ dst, __nv_bfloat16* __restrict__ dst_hi,
                                            __nv_bfloat16* __restrict__ dst_lo, int64_t row0, int64_t M, int64_t ld, int64_t col0, int g, int t) {
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int64_t gr = row0 + g + rr * 8;
    if (gr < M) {
      const int64_t off = gr * ld + col0 + 16 * t;
#pragma unroll
      for (int q = 0; q < 4; ++q) {                // q = 0, 1: features 16t + 4q ..; q = 2, 3: features 16t + 8 + 4(q-2) ..
        const int x = rr * 2 + (q >> 1), n0 = (q & 1) * 4;
        const float4 v = make_float4(acc[n0][x] * scale, acc[n0 + 1][x] * scale, acc[n0 + 2][x] * scale, acc[n0 + 3][x] * scale);
        if (dst) st4(dst + off + 4 * q, v);
        if (dst_hi) store_split4(dst_hi + off + 4 * q, dst_lo + off + 4 * q, v);
      }
    }
  }
}

template <int L, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
attn_bwd_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ e1, const float* __restrict__ e2,
                    const float* __restrict__ probs, const float* __restrict__ dout, float* __restrict__ dqkv,
                    __nv_bfloat16* __restrict__ dqkv_hi, __nv_bfloat16* __restrict__ dqkv_lo,
                    float* __restrict__ de1, float* __restrict__ de2, int64_t M, int H, int64_t ntiles, float q_scale, int keep,
                    const unsigned char* __restrict__ drop_keep, float drop_scale) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PER_WARP = 4 * TR * TS + 4 * TR * PSM;
  float* e1s = smem;                     // rows >= L are zero: TF32 "big" parts of the tables (pair-interleaved, 8 x TPS) ...
  float* e2s = e1s + TR * TS;
  float* e1l = e2s + TR * TS;            // ... and their "small" parts (split once per CTA, not once per tile)
  float* e2l = e1l + TR * TS;
  float* stage = e2l + TR * TS;          // per-warp staging areas
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = stage + warp * PER_WARP;
  float* ks = qs + TR * TS;
  float* vs = ks + TR * TS;
  float* gs = vs + TR * TS;              // dO tile
  float* As = gs + TR * TS;              // A' = probabilities * dropout factor, [16][PSM], zero outside the block
  float* Ds = As + TR * PSM;             // dS
  float* T1 = Ds + TR * PSM;             // dS skewed onto the causal table index
  float* T2 = T1 + TR * PSM;             // ... anticausal
  const int h = blockIdx.y;
  const int d = H * HD;
  const int kmask = keep - 1, kshift = __ffs(keep) - 1;
  for (int i = threadIdx.x; i < TR * 16; i += blockDim.x) {
    int r = i / 16, c = (i % 16) * 4;
    float4 a = make_float4(0, 0, 0, 0), b = a;
    if (r < L) {
      a = ld4(e1 + ((size_t)h * L + r) * HD + c);
      b = ld4(e2 + ((size_t)h * L + r) * HD + c);
    }
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
    uint32_t ab_[4], as_[4], bb_[4], bs_[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      tf32_split(av[x], ab_[x], as_[x]);
      tf32_split(bv[x], bb_[x], bs_[x]);
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {                  // pair-interleaved layout, see load_bfrag_n_tab
      const int o = tab_index(r, c + x);
      e1s[o] = __uint_as_float(ab_[x]);
      e1l[o] = __uint_as_float(as_[x]);
      e2s[o] = __uint_as_float(bb_[x]);
      e2l[o] = __uint_as_float(bs_[x]);
    }
  }
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;                     // mma fragment coordinates
  const int sr = lane >> 1, sh = lane & 1;                   // skew step: row sr, table indices sh*8..+7
  float acc_e1[8][4], acc_e2[8][4];                          // dE1 / dE2: rows (table index) g, g + 8; permuted features
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int x = 0; x < 4; ++x) { acc_e1[n][x] = 0.f; acc_e2[n][x] = 0.f; }

  const int64_t tile0 = (int64_t)blockIdx.x * WARPS + warp, tstride = (int64_t)gridDim.x * WARPS;
  auto fetch_g = [&](int64_t r0) {                          // dO exists for the kept query rows only
    if (keep == 1) load_tile_async(gs, dout + h * HD, r0, M, d, lane);
    else load_tile_compact_async(gs, dout + h * HD, r0, M, d, lane, L, kmask, kshift);
  };
  // The next tile's operands are fetched with cp.async into the same buffers as soon as the current tile's last reader is
  // done: v and dO after (1) + (5), k after (3), q after (4) + (6).  The groups are committed in that order - [v + dO], [k],
  // [q] - and each is awaited only where its tile is first read ("all but the two newest groups"), so k and q, which are
  // released late, have the whole front of the next tile to arrive.
  if (tile0 < ntiles) {
    load_tile_async(vs, qkv + 2 * d + h * HD, tile0 * TR, M, 3 * d, lane);
    fetch_g(tile0 * TR);
  }
  cp_async_commit();
  if (tile0 < ntiles) load_tile_async(ks, qkv + d + h * HD, tile0 * TR, M, 3 * d, lane);
  cp_async_commit();
  if (tile0 < ntiles) load_tile_async(qs, qkv + h * HD, tile0 * TR, M, 3 * d, lane);
  cp_async_commit();
  for (int64_t tile = tile0; tile < ntiles; tile += tstride) {
    const int64_t row0 = tile * TR;
    const int64_t nrow0 = (tile + tstride) * TR;
    const bool has_next = tile + tstride < ntiles;
    // probabilities of this lane's accumulator entries (rows g, g + 8; key columns n*8 + 2t, + 1) and, with dropout, the
    // factor m = keep * scale of each: the forward used A' = A * m in O = A' V, so dV = A'^T dO and dA = (dO V^T) * m,
    // while the softmax backward uses A itself
    float a[2][4], mk[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int r = g + (x >> 1) * 8, c = n * 8 + 2 * t + (x & 1);
        const int64_t gr = row0 + r;
        const int i = r % L;
        float v = 0.f, m = 1.f;
        if (gr < M && (i & kmask) == 0 && c / L == r / L) {
          const int64_t pi = (((gr / L) * H + h) * L + i) * L + (c % L);
          v = probs[pi];
          if (drop_keep != nullptr) m = drop_keep[pi] ? drop_scale : 0.f;
        }
        a[n][x] = v;
        mk[n][x] = m;
      }
    cp_async_wait_but<2>();                         // v and dO of this tile have landed
    __syncwarp();

    // ---- (1) dA = dO V^T
    float da[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) da[n][x] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 x0 = ld4(gs + g * TS + 16 * j + 4 * t), x1 = ld4(gs + (g + 8) * TS + 16 * j + 4 * t);
      const float4 v0 = ld4(vs + g * TS + 16 * j + 4 * t), v1 = ld4(vs + (g + 8) * TS + 16 * j + 4 * t);
      uint32_t ab[4], as[4];
      // k-step 2j: physical features 16j + 4t + {0, 1}
      tf32_split(x0.x, ab[0], as[0]); tf32_split(x1.x, ab[1], as[1]); tf32_split(x0.y, ab[2], as[2]); tf32_split(x1.y, ab[3], as[3]);
      mma_3x(da[0], ab, as, v0.x, v0.y);
      mma_3x(da[1], ab, as, v1.x, v1.y);
      // k-step 2j + 1: physical features 16j + 4t + {2, 3}
      tf32_split(x0.z, ab[0], as[0]); tf32_split(x1.z, ab[1], as[1]); tf32_split(x0.w, ab[2], as[2]); tf32_split(x1.w, ab[3], as[3]);
      mma_3x(da[0], ab, as, v0.z, v0.w);
      mma_3x(da[1], ab, as, v1.z, v1.w);
    }
    // ---- (2) dS = A * (dA * m - rowsum(A * dA * m)); A' and dS to shared memory
    float rd[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        da[n][x] *= mk[n][x];
        rd[x >> 1] = fmaf(a[n][x], da[n][x], rd[x >> 1]);
      }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      rd[rr] += __shfl_xor_sync(0xffffffffu, rd[rr], 1);
      rd[rr] += __shfl_xor_sync(0xffffffffu, rd[rr], 2);
    }
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int o = (g + rr * 8) * PSM + n * 8 + 2 * t;
        *reinterpret_cast<float2*>(Ds + o) = make_float2(a[n][rr * 2] * (da[n][rr * 2] - rd[rr]), a[n][rr * 2 + 1] * (da[n][rr * 2 + 1] - rd[rr]));
        *reinterpret_cast<float2*>(As + o) = make_float2(a[n][rr * 2] * mk[n][rr * 2], a[n][rr * 2 + 1] * mk[n][rr * 2 + 1]);
      }
    __syncwarp();
    // ---- skew dS onto the relative-table index
    {
      const int i = sr % L, gb = (sr / L) * L;
      float t1[8], t2[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const int idx = sh * 8 + x;
        const int j1 = i - (L - 1 - idx);            // e1 row idx <-> key j = i - (L-1-idx) <= i
        const int j2 = i + idx;                      // e2 row idx (>= 1) <-> key j = i + idx > i
        t1[x] = (idx < L && j1 >= 0) ? Ds[sr * PSM + gb + j1] : 0.f;
        t2[x] = (idx >= 1 && idx < L && j2 < L) ? Ds[sr * PSM + gb + j2] : 0.f;
      }
      st4(T1 + sr * PSM + sh * 8, make_float4(t1[0], t1[1], t1[2], t1[3]));
      st4(T1 + sr * PSM + sh * 8 + 4, make_float4(t1[4], t1[5], t1[6], t1[7]));
      st4(T2 + sr * PSM + sh * 8, make_float4(t2[0], t2[1], t2[2], t2[3]));
      st4(T2 + sr * PSM + sh * 8 + 4, make_float4(t2[4], t2[5], t2[6], t2[7]));
    }
    __syncwarp();

    float acc[8][4];
    uint32_t bb[8], bs[8], cb[8], cs[8], ab[4], as[4];
    // ---- (5) dV = A'^T dO
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) acc[n][x] = 0.f;
#pragma unroll 1                                  // (rolled: 3 % faster than unrolled, the loop body is instruction-cache bound)
    for (int kk = 0; kk < 2; ++kk) {
      load_bfrag_n(gs, kk * 8 + t, kk * 8 + t + 4, g, bb, bs, cb, cs);
      load_afrag_t(As, kk, g, t, ab, as);
#pragma unroll
      for (int n = 0; n < 8; ++n) mma_3x_pre(acc[n], ab, as, bb[n], bs[n], cb[n], cs[n]);
    }
    store_acc_n(acc, 1.f, dqkv, dqkv_hi, dqkv_lo, row0, M, 3 * d, 2 * d + h * HD, g, t);
    __syncwarp();                                   // v and dO are free
    if (has_next) {
      load_tile_async(vs, qkv + 2 * d + h * HD, nrow0, M, 3 * d, lane);
      fetch_g(nrow0);
    }
    cp_async_commit();
    cp_async_wait_but<2>();                         // k of this tile has landed
    __syncwarp();
    // ---- (3) dQ~ = dS K + T1 e1 + T2 e2
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) acc[n][x] = 0.f;
#pragma unroll 1                                  // (rolled: 3 % faster than unrolled, the loop body is instruction-cache bound)
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int which = 0; which < 3; ++which) {
        if (which > 0 && kk * 8 >= L) continue;      // table rows >= L are zero
        load_afrag(which == 0 ? Ds : (which == 1 ? T1 : T2), kk, g, t, ab, as);
        if (which == 0) load_bfrag_n(ks, kk * 8 + t, kk * 8 + t + 4, g, bb, bs, cb, cs);
        else load_bfrag_n_tab(which == 1 ? e1s : e2s, which == 1 ? e1l : e2l, kk, t, g, bb, bs, cb, cs);
#pragma unroll
        for (int n = 0; n < 8; ++n) mma_3x_pre(acc[n], ab, as, bb[n], bs[n], cb[n], cs[n]);
      }
    }
    store_acc_n(acc, q_scale, dqkv, dqkv_hi, dqkv_lo, row0, M, 3 * d, h * HD, g, t);
    __syncwarp();                                   // k is free
    if (has_next) load_tile_async(ks, qkv + d + h * HD, nrow0, M, 3 * d, lane);
    cp_async_commit();
    cp_async_wait_but<2>();                         // q of this tile has landed
    __syncwarp();
    // ---- (4) dK = dS^T Q~ and (6) dE1 += T1^T Q~, dE2 += T2^T Q~ on the same Q~ fragments
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int x = 0; x < 4; ++x) acc[n][x] = 0.f;
#pragma unroll 1                                  // (rolled: 3 % faster than unrolled, the loop body is instruction-cache bound)
    for (int kk = 0; kk < 2; ++kk) {
      load_bfrag_n(qs, kk * 8 + t, kk * 8 + t + 4, g, bb, bs, cb, cs);
      load_afrag_t(Ds, kk, g, t, ab, as);
#pragma unroll
      for (int n = 0; n < 8; ++n) mma_3x_pre(acc[n], ab, as, bb[n], bs[n], cb[n], cs[n]);
      if (L > 1) {
        load_afrag_t(T1, kk, g, t, ab, as);
#pragma unroll
        for (int n = 0; n < 8; ++n) mma_3x_pre(acc_e1[n], ab, as, bb[n], bs[n], cb[n], cs[n]);
        load_afrag_t(T2, kk, g, t, ab, as);
#pragma unroll
        for (int n = 0; n < 8; ++n) mma_3x_pre(acc_e2[n], ab, as, bb[n], bs[n], cb[n], cs[n]);
      }
    }
    store_acc_n(acc, 1.f, dqkv, dqkv_hi, dqkv_lo, row0, M, 3 * d, d + h * HD, g, t);
    __syncwarp();                                   // q is free
    if (has_next) load_tile_async(qs, qkv + h * HD, nrow0, M, 3 * d, lane);
    cp_async_commit();
  }

  // reduce the table gradients over the CTA's warps, then one atomic per element per CTA
  __syncthreads();
  float* red = stage;                   // reuse the per-warp staging areas: [WARPS][2][16][64]
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = g + rr * 8, x = rr * 2 + (q >> 1), n0 = (q & 1) * 4;
      st4(red + ((warp * 2 + 0) * TR + idx) * HD + 16 * t + 4 * q, make_float4(acc_e1[n0][x], acc_e1[n0 + 1][x], acc_e1[n0 + 2][x], acc_e1[n0 + 3][x]));
      st4(red + ((warp * 2 + 1) * TR + idx) * HD + 16 * t + 4 * q, make_float4(acc_e2[n0][x], acc_e2[n0 + 1][x], acc_e2[n0 + 2][x], acc_e2[n0 + 3][x]));
    }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * L * HD; e += blockDim.x) {
    const int which = e / (L * HD), rem = e % (L * HD);       // rem = idx * 64 + col, idx < L
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += red[(w * 2 + which) * TR * HD + rem];
    float* dst = (which == 0 ? de1 : de2) + (size_t)h * L * HD + rem;
    atomicAdd(dst, s);
  }
}

int g_attn_fwd_algo = 0;       // 0 = tensor-core (mma.sync 3xTF32) forward, 1 = CUDA-core forward (exact-fp32 mode / tuning hook)
int g_attn_bwd_algo = 0;       // same for the backward kernel

template <int L>
int launch_fwd(const float* qkv, const float* e1, const float* e2, float* out, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo,
               float* probs, int64_t M, int H, int keep, const unsigned char* drop_keep, float drop_scale, cudaStream_t st) {
  constexpr int WARPS = 8;
  int64_t ntiles = vq_cdiv(M, TR);
  int64_t gx = std::min<int64_t>(vq_cdiv(ntiles, WARPS), std::max(1, vq_sm_count() * 2 / H) * 4);
  if (g_attn_fwd_algo == 0) {
    size_t smem = (size_t)(2 * TR * TS + WARPS * 3 * TR * TS) * sizeof(float);
    static_assert(2 * TR * PSM <= TR * TS, "R1 and R2 must fit in the q tile");
    static bool done_mma = false;
    if (!done_mma) {
      VQ_CUDA_CALL(cudaFuncSetAttribute(attn_fwd_mma_kernel<L, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      done_mma = true;
    }
    attn_fwd_mma_kernel<L, WARPS><<<dim3((unsigned)gx, H), WARPS * 32, smem, st>>>(qkv, e1, e2, out, out_hi, out_lo, probs, M, H, ntiles, keep, drop_keep, drop_scale);
    VQ_CUDA_LAUNCH_CHECK("attn_fwd_mma_kernel");
    return VQ_OK;
  }
  size_t smem = (size_t)(2 * L * TS + WARPS * 3 * TR * TS) * sizeof(float);
  static bool done = false;
  if (!done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(attn_fwd_kernel<L, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
  }
  attn_fwd_kernel<L, WARPS><<<dim3((unsigned)gx, H), WARPS * 32, smem, st>>>(qkv, e1, e2, out, out_hi, out_lo, probs, M, H, ntiles, keep, drop_keep, drop_scale);
  VQ_CUDA_LAUNCH_CHECK("attn_fwd_kernel");
  return VQ_OK;
}
template <int L>
int launch_bwd(const float* qkv, const float* e1, const float* e2, const float* probs, const float* dout, float* dqkv,
               __nv_bfloat16* dqkv_hi, __nv_bfloat16* dqkv_lo, float* de1, float* de2, int64_t M, int H, float q_scale, int keep,
               const unsigned char* drop_keep, float drop_scale, cudaStream_t st) {
  constexpr int WARPS = 4;
  int64_t ntiles = vq_cdiv(M, TR);
  size_t smem = (size_t)(2 * TR * TS + WARPS * (4 * TR * TS + 4 * TR * PS)) * sizeof(float);
  static_assert(4 * TR * TS + 4 * TR * PS >= 2 * TR * HD, "reduction scratch must fit in the per-warp staging area");
  static bool done = false;
  if (!done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(attn_bwd_kernel<L, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
  }
  int64_t gx = std::min<int64_t>(vq_cdiv(ntiles, WARPS), std::max(1, vq_sm_count() * 2 / H) * 2);
  if (g_attn_bwd_algo == 0) {
    static_assert(PSM == PS, "both backward kernels use the same staging layout");
    const size_t smem_mma = smem + (size_t)(2 * TR * TS) * sizeof(float);      // + the small parts of the split tables
    static bool done_mma = false;
    if (!done_mma) {
      VQ_CUDA_CALL(cudaFuncSetAttribute(attn_bwd_mma_kernel<L, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma));
      done_mma = true;
    }
    attn_bwd_mma_kernel<L, WARPS><<<dim3((unsigned)gx, H), WARPS * 32, smem_mma, st>>>(qkv, e1, e2, probs, dout, dqkv, dqkv_hi, dqkv_lo, de1, de2, M, H, ntiles, q_scale, keep, drop_keep, drop_scale);
    VQ_CUDA_LAUNCH_CHECK("attn_bwd_mma_kernel");
    return VQ_OK;
  }
  attn_bwd_kernel<L, WARPS><<<dim3((unsigned)gx, H), WARPS * 32, smem, st>>>(qkv, e1, e2, probs, dout, dqkv, dqkv_hi, dqkv_lo, de1, de2, M, H, ntiles, q_scale, keep, drop_keep, drop_scale);
  VQ_CUDA_LAUNCH_CHECK("attn_bwd_kernel");
  return VQ_OK;
}

}  // namespace

// qkv: (P*L, 3*H*64) with the q third already multiplied by hd^-0.5 (done in the in-proj GEMM epilogue);
// out (fp32) and/or (out_hi, out_lo) (bf16 split, operand of the out-proj GEMM): (P*L/keep, H*64) — only the query rows
// with token % keep == 0 are computed and stored, compacted (keep = 1: all rows; keep = downscale factor in the last
// layer of a stack = exact dead-row elimination); probs: (P, H, L, L) saved for backward (kept rows only; may be NULL).
/* kernel choice, forward / backward: 0 = tensor cores (mma.sync 3xTF32, default), 1 = CUDA cores (exact fp32) */
VQ_API int vqcpc_attn_set_fwd_algo(int algo) {
  g_attn_fwd_algo = algo == 1 ? 1 : 0;
  return VQ_OK;
}
VQ_API int vqcpc_attn_set_bwd_algo(int algo) {
  g_attn_bwd_algo = algo == 1 ? 1 : 0;
  return VQ_OK;
}

// drop_keep (P, H, L, L) uint8 or NULL: dropout on the attention weights (F.dropout, multihead_attention_custom.py:339) with
// an explicit keep mask, drop_scale = 1 / (1 - p); `probs` always receives the softmax output itself.
VQ_API int vqcpc_attn_fwd_ex(const float* qkv, const float* e1, const float* e2, float* out, void* out_hi, void* out_lo,
                             float* probs, int64_t P, int L, int H, int head_dim, int keep, const unsigned char* drop_keep,
                             float drop_scale, void* stream) {
  VQ_CHECK_ARG(head_dim == HD, "attn_fwd: head_dim must be 64 (got %d)", head_dim);
  VQ_CHECK_ARG(H >= 1 && H <= 64, "attn_fwd: bad head count %d", H);
  VQ_CHECK_ARG(keep >= 1 && L % keep == 0 && (keep & (keep - 1)) == 0, "attn_fwd: keep=%d must be a power of two dividing L=%d", keep, L);
  VQ_CHECK_ARG((out || out_hi) && ((out_hi == nullptr) == (out_lo == nullptr)), "attn_fwd: need out and/or (out_hi, out_lo)");
  if (P == 0) return VQ_OK;
  cudaStream_t st = vq_stream(stream);
  int64_t M = P * L;
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
  switch (L) {
    case 16: return launch_fwd<16>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, drop_keep, drop_scale, st);
    case 8: return launch_fwd<8>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, drop_keep, drop_scale, st);
    case 4: return launch_fwd<4>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, drop_keep, drop_scale, st);
    case 2: return launch_fwd<2>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, drop_keep, drop_scale, st);
    case 1: return launch_fwd<1>(qkv, e1, e2, out, oh, ol, probs, M, H, keep, drop_keep, drop_scale, st);
    default: VQ_CHECK_ARG(false, "attn_fwd: block length L must be 1, 2, 4, 8 or 16 (got %d)", L);
  }
  return VQ_OK;
}

VQ_API int vqcpc_attn_fwd(const float* qkv, const float* e1, const float* e2, float* out, void* out_hi, void* out_lo,
                          float* probs, int64_t P, int L, int H, int head_dim, int keep, void* stream) {
  return vqcpc_attn_fwd_ex(qkv, e1, e2, out, out_hi, out_lo, probs, P, L, H, head_dim, keep, nullptr, 1.f, stream);
}

// d_out: (P*L/keep, H*64) gradient of the kept query rows (see vqcpc_attn_fwd).  d_qkv (fp32, P*L x 3*H*64) and/or its
// bf16 split (d_qkv_hi, d_qkv_lo) are overwritten (q third = gradient wrt the UNscaled projection: q_scale * dq~);
// d_e1 / d_e2 (H*L, 64) are ACCUMULATED.
VQ_API int vqcpc_attn_bwd_ex(const float* qkv, const float* e1, const float* e2, const float* probs, const float* d_out,
                             float* d_qkv, void* d_qkv_hi, void* d_qkv_lo, float* d_e1, float* d_e2, int64_t P, int L, int H,
                             int head_dim, float q_scale, int keep, const unsigned char* drop_keep, float drop_scale,
                             void* stream) {
  VQ_CHECK_ARG(head_dim == HD, "attn_bwd: head_dim must be 64 (got %d)", head_dim);
  VQ_CHECK_ARG(H >= 1 && H <= 64, "attn_bwd: bad head count %d", H);
  VQ_CHECK_ARG(keep >= 1 && L % keep == 0 && (keep & (keep - 1)) == 0, "attn_bwd: keep=%d must be a power of two dividing L=%d", keep, L);
  VQ_CHECK_ARG((d_qkv || d_qkv_hi) && ((d_qkv_hi == nullptr) == (d_qkv_lo == nullptr)), "attn_bwd: need d_qkv and/or its split");
  if (P == 0) return VQ_OK;
  cudaStream_t st = vq_stream(stream);
  int64_t M = P * L;
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(d_qkv_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(d_qkv_lo);
  switch (L) {
    case 16: return launch_bwd<16>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, drop_keep, drop_scale, st);
    case 8: return launch_bwd<8>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, drop_keep, drop_scale, st);
    case 4: return launch_bwd<4>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, drop_keep, drop_scale, st);
    case 2: return launch_bwd<2>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, drop_keep, drop_scale, st);
    case 1: return launch_bwd<1>(qkv, e1, e2, probs, d_out, d_qkv, oh, ol, d_e1, d_e2, M, H, q_scale, keep, drop_keep, drop_scale, st);
    default: VQ_CHECK_ARG(false, "attn_bwd: block length L must be 1, 2, 4, 8 or 16 (got %d)", L);
  }
  return VQ_OK;
}

VQ_API int vqcpc_attn_bwd(const float* qkv, const float* e1, const float* e2, const float* probs, const float* d_out,
                          float* d_qkv, void* d_qkv_hi, void* d_qkv_lo, float* d_e1, float* d_e2, int64_t P, int L, int H,
                          int head_dim, float q_scale, int keep, void* stream) {
  return vqcpc_attn_bwd_ex(qkv, e1, e2, probs, d_out, d_qkv, d_qkv_hi, d_qkv_lo, d_e1, d_e2, P, L, H, head_dim, q_scale, keep,
                           nullptr, 1.f, stream);
}
