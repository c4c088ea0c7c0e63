// Full-sequence relative self-attention (L up to 1024 tokens per sequence) for the student / teacher / auxiliary-decoder
// transformers of the distillation trainer: the reference's TransformerEncoderLayerCustom over the WHOLE
// chorale window (teachers/teacher_relative.py:33-47, auxiliary_decoders/auxiliary_decoder_relative.py:58-74), i.e.
// multihead_attention_custom.py:274-345 with the skewed bias of subsampled_relative_attention.py:30-122 in closed form:
//
//     s[i, j] = q~_i . k_j + q~_i . r(i, j),   r(i, j) = e1[h, L-1-(i-j)]  if j <= i,   e2[h, j-i]  if j > i
//     a = softmax_j(s),  o_i = sum_j a[i, j] v_j
//
// Layout as in attention.cu: qkv (P*L, 3*H*64) with the q third pre-scaled, e1 / e2 (H*L, 64), probabilities (P, H, L, L).
// The per-block kernels of attention.cu keep a whole block in registers (L <= 16); here a CTA owns R query rows (forward,
// dQ) or R key columns (dK, dV) of one (sequence, head) and streams the other side from L1 / L2, the L x L score matrix
// goes through global memory (the reference materialises and returns it as well).  fp32 CUDA-core arithmetic: these
// models run at batch 8 - 64 windows of 384 tokens, where attention is < 1 GFLOP per layer; the GEMMs of the layer use the
// tcgen05 path of gemm_tc*.cu.  Deterministic: no atomics (dE1 / dE2 are gathered along the diagonals).
#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int R = 8;            // query rows (or key columns) per CTA
constexpr int THREADS = 256;
constexpr int KT = 32;           // keys per shared-memory tile
constexpr int PADW = HD + 4;     // padded row (floats): 16-byte reads of 8 consecutive rows hit all 32 banks once

__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - __uint_as_float(hi << 16), y - __uint_as_float(hi & 0xffff0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void store2(float* o, __nv_bfloat16* oh, __nv_bfloat16* ol, size_t idx, float x, float y) {
  if (o) *reinterpret_cast<float2*>(o + idx) = make_float2(x, y);
  if (oh) {
    uint32_t hi, lo;
    split_pair(x, y, hi, lo);
    *reinterpret_cast<uint32_t*>(oh + idx) = hi;
    *reinterpret_cast<uint32_t*>(ol + idx) = lo;
  }
}
// relative-position row paired with (query i, key j)
__device__ __forceinline__ const float* rel_row(const float* e1h, const float* e2h, int L, int i, int j) {
  return (j <= i) ? e1h + (size_t)(L - 1 - (i - j)) * HD : e2h + (size_t)(j - i) * HD;
}

// grid (ceil(L / R), H, P); shared: q rows [R][64] + scores [R][L]
__global__ void __launch_bounds__(THREADS)
attn_long_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ e1, const float* __restrict__ e2,
                     float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                     float* __restrict__ probs, int L, int H, int mask_mode) {
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;                    // [R][64]
  float* sc = sm + R * HD;           // [R][L]
  const int i0 = blockIdx.x * R, h = blockIdx.y;
  const int64_t p = blockIdx.z;
  const int ld = 3 * H * HD;
  const float* base = qkv + p * L * (int64_t)ld;
  const float* e1h = e1 + (size_t)h * L * HD;
  const float* e2h = e2 + (size_t)h * L * HD;
  const int rows = min(R, L - i0);
  // mask_mode 1 (causal: keys <= query, decoder.py:295-298) / 2 (anticausal: keys >= query, :300-306): only the key tiles that
  // hold a visible key are visited; inside them hidden scores are -inf like the reference's additive mask
  const int jlo = mask_mode == 2 ? (i0 / KT) * KT : 0, jhi = mask_mode == 1 ? min(L, i0 + rows) : L;
  for (int t = threadIdx.x; t < R * HD; t += THREADS) {
    const int i = t / HD, d = t % HD;
    qs[t] = (i < rows) ? base[(int64_t)(i0 + i) * ld + h * HD + d] : 0.f;
  }
  __syncthreads();
  // scores, 32 keys at a time: the key tile and the two slabs of relative rows it pairs with (39 consecutive rows of e1 for
  // j <= i, of e2 for j > i: row = base + (jj - ii), jj = key in tile, ii = query in CTA) are staged in shared memory with
  // coalesced loads; thread -> (query ii, key jj) then reads padded rows without bank conflicts
  {
    float* ks = sc + (size_t)R * L;                 // [KT][PADW]
    float* s1 = ks + KT * PADW;                     // [KT + R - 1][PADW]  e1 slab
    float* s2 = s1 + (KT + R - 1) * PADW;           // [KT + R - 1][PADW]  e2 slab
    const int ii = threadIdx.x >> 5, jj = threadIdx.x & 31;
    for (int j0 = jlo; j0 < jhi; j0 += KT) {
      const bool need1 = e1 != nullptr && j0 <= i0 + R - 1, need2 = e2 != nullptr && j0 + KT - 1 > i0;   // NULL tables: no bias
      const int b1 = (L - 1 - i0 + j0) - (R - 1), b2 = (j0 - i0) - (R - 1);      // table row of slab row 0
      __syncthreads();
      for (int t = threadIdx.x; t < KT * (HD / 4); t += THREADS) {
        const int r = t / (HD / 4), c = t % (HD / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 + r < L) v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(j0 + r) * ld + (H + h) * HD) + c);
        *reinterpret_cast<float4*>(ks + r * PADW + c * 4) = v;
      }
      for (int t = threadIdx.x; t < (KT + R - 1) * (HD / 4); t += THREADS) {
        const int r = t / (HD / 4), c = t % (HD / 4);
        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1;
        if (need1 && b1 + r >= 0 && b1 + r < L) v1 = __ldg(reinterpret_cast<const float4*>(e1h + (size_t)(b1 + r) * HD) + c);
        if (need2 && b2 + r >= 1 && b2 + r < L) v2 = __ldg(reinterpret_cast<const float4*>(e2h + (size_t)(b2 + r) * HD) + c);
        *reinterpret_cast<float4*>(s1 + r * PADW + c * 4) = v1;
        *reinterpret_cast<float4*>(s2 + r * PADW + c * 4) = v2;
      }
      __syncthreads();
      const int j = j0 + jj;
      if (ii < rows && j < L) {
        const float4* kj = reinterpret_cast<const float4*>(ks + jj * PADW);
        const float4* rr = reinterpret_cast<const float4*>(((j <= i0 + ii) ? s1 : s2) + (jj - ii + R - 1) * PADW);
        const float4* qi = reinterpret_cast<const float4*>(qs + ii * HD);
        float a0 = 0.f, a1 = 0.f;          // q.k and q.r accumulated separately, like the reference (bmm + einsum)
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          const float4 q4 = qi[c], k4 = kj[c], r4 = rr[c];
          a0 += q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
          a1 += q4.x * r4.x + q4.y * r4.y + q4.z * r4.z + q4.w * r4.w;
        }
        const bool hidden = (mask_mode == 1 && j > i0 + ii) || (mask_mode == 2 && j < i0 + ii);
        sc[ii * L + j] = hidden ? -INFINITY : a0 + a1;
      }
    }
  }
  __syncthreads();
  // softmax: one warp per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < rows) {
    float* s = sc + warp * L;
    float m = -INFINITY;
    for (int j = jlo + lane; j < jhi; j += 32) m = fmaxf(m, s[j]);
    m = warp_max(m);
    float z = 0.f;
    for (int j = jlo + lane; j < jhi; j += 32) {
      const float ev = __expf(s[j] - m);
      s[j] = ev;
      z += ev;
    }
    z = warp_sum(z);
    const float inv = 1.f / z;
    float* pg = probs ? probs + ((p * H + h) * (int64_t)L + (i0 + warp)) * L : nullptr;
    for (int j = jlo + lane; j < jhi; j += 32) {
      const float a = s[j] * inv;
      s[j] = a;
      if (pg) pg[j] = a;
    }
    if (pg) {                                        // key tiles that were never visited
      for (int j = lane; j < jlo; j += 32) pg[j] = 0.f;
      for (int j = jhi + lane; j < L; j += 32) pg[j] = 0.f;
    }
  }
  __syncthreads();
  // o = a v: thread -> (row, feature pair)
  {
    const int i = threadIdx.x >> 5, d = (threadIdx.x & 31) * 2;
    if (i < rows) {
      const float* s = sc + i * L;
      const float* vb = base + (2 * H + h) * HD + d;
      float o0 = 0.f, o1 = 0.f;
#pragma unroll 4
      for (int j = jlo; j < jhi; ++j) {
        const float2 v2 = __ldg(reinterpret_cast<const float2*>(vb + (int64_t)j * ld));
        const float a = s[j];
        o0 += a * v2.x;
        o1 += a * v2.y;
      }
      store2(out, out_hi, out_lo, (size_t)(p * L + i0 + i) * (H * HD) + h * HD + d, o0, o1);
    }
  }
}

// Backward, query side: dP = dO v^T, dS = a (dP - sum_j dP a), dq~_i = sum_j dS[i, j] (k_j + r(i, j)); dS goes to `ds`.
__global__ void __launch_bounds__(THREADS)
attn_long_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ e1, const float* __restrict__ e2,
                       const float* __restrict__ probs, const float* __restrict__ d_out, float* __restrict__ ds,
                       float* __restrict__ d_qkv, __nv_bfloat16* __restrict__ dq_hi, __nv_bfloat16* __restrict__ dq_lo,
                       int L, int H, float q_scale) {
  extern __shared__ __align__(16) float sm[];
  float* dos = sm;                   // [R][64]
  float* sc = sm + R * HD;           // [R][L]  dP, then dS
  const int i0 = blockIdx.x * R, h = blockIdx.y;
  const int64_t p = blockIdx.z;
  const int ld = 3 * H * HD;
  const float* base = qkv + p * L * (int64_t)ld;
  const float* e1h = e1 + (size_t)h * L * HD;
  const float* e2h = e2 + (size_t)h * L * HD;
  const int rows = min(R, L - i0);
  for (int t = threadIdx.x; t < R * HD; t += THREADS) {
    const int i = t / HD, d = t % HD;
    dos[t] = (i < rows) ? d_out[(size_t)(p * L + i0 + i) * (H * HD) + h * HD + d] : 0.f;
  }
  __syncthreads();
  {  // dP = dO v^T, 32 keys at a time through a padded shared-memory tile of v (coalesced loads, conflict-free reads)
    float* vs = sc + (size_t)R * L;                 // [KT][PADW]
    const int ii = threadIdx.x >> 5, jj = threadIdx.x & 31;
    for (int j0 = 0; j0 < L; j0 += KT) {
      __syncthreads();
      for (int t = threadIdx.x; t < KT * (HD / 4); t += THREADS) {
        const int r = t / (HD / 4), c = t % (HD / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 + r < L) v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(j0 + r) * ld + (2 * H + h) * HD) + c);
        *reinterpret_cast<float4*>(vs + r * PADW + c * 4) = v;
      }
      __syncthreads();
      const int j = j0 + jj;
      if (ii < rows && j < L) {
        const float4* vj = reinterpret_cast<const float4*>(vs + jj * PADW);
        const float4* di = reinterpret_cast<const float4*>(dos + ii * HD);
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          const float4 d4 = di[c], v4 = vj[c];
          a += d4.x * v4.x + d4.y * v4.y + d4.z * v4.z + d4.w * v4.w;
        }
        sc[ii * L + j] = a;
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < rows) {
    float* s = sc + warp * L;
    const float* pg = probs + ((p * H + h) * (int64_t)L + (i0 + warp)) * L;
    float* dg = ds + ((p * H + h) * (int64_t)L + (i0 + warp)) * L;
    float dot = 0.f;
    for (int j = lane; j < L; j += 32) dot += s[j] * pg[j];
    dot = warp_sum(dot);
    for (int j = lane; j < L; j += 32) {
      const float v = pg[j] * (s[j] - dot);
      s[j] = v;
      dg[j] = v;
    }
  }
  __syncthreads();
  {
    const int i = threadIdx.x >> 5, d = (threadIdx.x & 31) * 2;
    if (i < rows) {
      const float* s = sc + i * L;
      const float* kb = base + (H + h) * HD + d;
      float g0 = 0.f, g1 = 0.f;
#pragma unroll 2
      for (int j = 0; j < L; ++j) {
        const float2 k2 = __ldg(reinterpret_cast<const float2*>(kb + (int64_t)j * ld));
        const float2 r2 = __ldg(reinterpret_cast<const float2*>(rel_row(e1h, e2h, L, i0 + i, j) + d));
        const float w = s[j];
        g0 += w * (k2.x + r2.x);
        g1 += w * (k2.y + r2.y);
      }
      store2(d_qkv, dq_hi, dq_lo, (size_t)(p * L + i0 + i) * ld + h * HD + d, g0 * q_scale, g1 * q_scale);
    }
  }
}

// Backward, key side: dv_j = sum_i a[i, j] dO_i, dk_j = sum_i dS[i, j] q~_i for R key columns per CTA.
__global__ void __launch_bounds__(THREADS)
attn_long_bwd_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ probs, const float* __restrict__ ds,
                        const float* __restrict__ d_out, float* __restrict__ d_qkv, __nv_bfloat16* __restrict__ dq_hi,
                        __nv_bfloat16* __restrict__ dq_lo, int L, int H) {
  extern __shared__ __align__(16) float sm[];
  float* pa = sm;                    // [L][R]  a[:, j0 : j0 + R]
  float* pd = sm + (size_t)L * R;    // [L][R]  dS[:, j0 : j0 + R]
  const int j0 = blockIdx.x * R, h = blockIdx.y;
  const int64_t p = blockIdx.z;
  const int ld = 3 * H * HD;
  const float* base = qkv + p * L * (int64_t)ld;
  const int cols = min(R, L - j0);
  const float* pg = probs + (p * H + h) * (int64_t)L * L;
  const float* dg = ds + (p * H + h) * (int64_t)L * L;
  for (int t = threadIdx.x; t < L * R; t += THREADS) {
    const int i = t / R, c = t % R;
    const bool ok = c < cols;
    pa[t] = ok ? pg[(int64_t)i * L + j0 + c] : 0.f;
    pd[t] = ok ? dg[(int64_t)i * L + j0 + c] : 0.f;
  }
  __syncthreads();
  const int c = threadIdx.x >> 5, d = (threadIdx.x & 31) * 2;
  if (c < cols) {
    const float* dob = d_out + (size_t)p * L * (H * HD) + h * HD + d;
    const float* qb = base + h * HD + d;
    float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
#pragma unroll 2
    for (int i = 0; i < L; ++i) {
      const float2 do2 = __ldg(reinterpret_cast<const float2*>(dob + (size_t)i * (H * HD)));
      const float2 q2 = __ldg(reinterpret_cast<const float2*>(qb + (int64_t)i * ld));
      const float a = pa[i * R + c], w = pd[i * R + c];
      v0 += a * do2.x; v1 += a * do2.y;
      k0 += w * q2.x; k1 += w * q2.y;
    }
    const size_t row = (size_t)(p * L + j0 + c) * ld;
    store2(d_qkv, dq_hi, dq_lo, row + (H + h) * HD + d, k0, k1);
    store2(d_qkv, dq_hi, dq_lo, row + (2 * H + h) * HD + d, v0, v1);
  }
}

// Backward, relative tables: row r of e1 pairs with the diagonal i - j = L-1-r, row r >= 1 of e2 with j - i = r (row 0 of e2
// is never read by the forward).  A CTA owns ER = 8 consecutive table rows: for a query i they pair with 8 CONSECUTIVE keys, so
// one 32-byte piece of the dS row and one q~ row feed 8 accumulators.  grid (ceil(L / 8), H, 2); 256 threads = 4 groups of 64
// features splitting the queries; the result is ACCUMULATED into d_e (the parameter's gradient buffer).
constexpr int ER = 8;
__global__ void __launch_bounds__(THREADS)
attn_long_bwd_e_kernel(const float* __restrict__ qkv, const float* __restrict__ ds, float* __restrict__ d_e1,
                       float* __restrict__ d_e2, int64_t P, int L, int H) {
  __shared__ float red[4][ER][HD];
  const int r0 = blockIdx.x * ER, h = blockIdx.y, which = blockIdx.z;
  const int ld = 3 * H * HD;
  const int g = threadIdx.x >> 6, d = threadIdx.x & 63;
  float acc[ER];
#pragma unroll
  for (int t = 0; t < ER; ++t) acc[t] = 0.f;
  // key paired with (query i, table row r0 + t): e1: j = i - (L-1) + r0 + t;  e2: j = i + r0 + t
  const int joff = which == 0 ? r0 - (L - 1) : r0;
  for (int64_t p = 0; p < P; ++p) {
    const float* dsp = ds + (p * H + h) * (int64_t)L * L;
    const float* qp = qkv + p * L * (int64_t)ld + h * HD + d;
    // queries for which at least one of the 8 keys exists
    const int i_lo = max(0, -joff - (ER - 1)), i_hi = min(L - 1, L - 1 - joff);
    for (int i = i_lo + g; i <= i_hi; i += 4) {
      const float q = __ldg(qp + (int64_t)i * ld);
      const float* row = dsp + (int64_t)i * L;
      const int jb = i + joff;
#pragma unroll
      for (int t = 0; t < ER; ++t) {
        const int j = jb + t;
        if (j >= 0 && j < L) acc[t] += __ldg(row + j) * q;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < ER; ++t) red[g][t][d] = acc[t];
  __syncthreads();
  if (g == 0) {
#pragma unroll
    for (int t = 0; t < ER; ++t) {
      const int r = r0 + t;
      if (r < L && !(which == 1 && r == 0)) {
        float* dst = (which == 0 ? d_e1 : d_e2) + ((size_t)h * L + r) * HD + d;
        *dst += red[0][t][d] + red[1][t][d] + red[2][t][d] + red[3][t][d];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ softmax cross-entropy
// One warp per row of logits (R x C, leading dimension ld): loss = lse(x) - x[t] (hard target) or lse(x) - sum softmax(y) x
// (distilled: target logits y, utils.py:131-159); d_logits = scale * (softmax(x) - target distribution).
__global__ void __launch_bounds__(128)
softmax_ce_kernel(const float* __restrict__ logits, int ld, const int64_t* __restrict__ target_idx,
                  const float* __restrict__ target_logits, int ldt, float* __restrict__ loss, float* __restrict__ d_logits,
                  int ldd, float scale, int64_t rows, int C) {
  const int64_t row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = logits + row * ld;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, x[c]);
  m = warp_max(m);
  float z = 0.f;
  for (int c = lane; c < C; c += 32) z += __expf(x[c] - m);
  z = warp_sum(z);
  const float lse = m + __logf(z);
  float l;
  if (target_idx) {
    const int t = (int)target_idx[row];
    l = lse - x[t];
    if (d_logits)
      for (int c = lane; c < C; c += 32) d_logits[row * ldd + c] = scale * (__expf(x[c] - lse) - (c == t ? 1.f : 0.f));
  } else {
    const float* y = target_logits + row * ldt;
    float my = -INFINITY;
    for (int c = lane; c < C; c += 32) my = fmaxf(my, y[c]);
    my = warp_max(my);
    float zy = 0.f;
    for (int c = lane; c < C; c += 32) zy += __expf(y[c] - my);
    zy = warp_sum(zy);
    const float lsey = my + __logf(zy);
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float pt = __expf(y[c] - lsey);
      dot += pt * (x[c] - lse);
      if (d_logits) d_logits[row * ldd + c] = scale * (__expf(x[c] - lse) - pt);
    }
    l = -warp_sum(dot);
  }
  if (lane == 0 && loss) loss[row] = l;
}

}  // namespace

static int attn_long_fwd_launch(const float* qkv, const float* e1, const float* e2, float* out, void* out_hi, void* out_lo,
                                float* probs, int64_t P, int L, int H, int head_dim, int mask_mode, void* stream) {
  VQ_CHECK_ARG(head_dim == HD, "attn_long_fwd: head_dim must be 64 (got %d)", head_dim);
  VQ_CHECK_ARG(L >= 1 && L <= 1024 && H >= 1 && H <= 64, "attn_long_fwd: bad shape L=%d H=%d", L, H);
  VQ_CHECK_ARG((out || out_hi) && ((out_hi == nullptr) == (out_lo == nullptr)), "attn_long_fwd: need out and/or (out_hi, out_lo)");
  if (P == 0) return VQ_OK;
  VQ_CHECK_ARG(P <= 65535, "attn_long_fwd: at most 65535 sequences per call (got %lld)", (long long)P);
  const size_t smem = (size_t)(R * HD + R * L + (KT + 2 * (KT + R - 1)) * PADW) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(attn_long_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((R * HD + R * 1024 + (KT + 2 * (KT + R - 1)) * PADW) * sizeof(float))));
    attr_done = true;
  }
  dim3 grid((unsigned)vq_cdiv(L, R), (unsigned)H, (unsigned)P);
  attn_long_fwd_kernel<<<grid, THREADS, smem, vq_stream(stream)>>>(qkv, e1, e2, out, reinterpret_cast<__nv_bfloat16*>(out_hi),
                                                                    reinterpret_cast<__nv_bfloat16*>(out_lo), probs, L, H, mask_mode);
  VQ_CUDA_LAUNCH_CHECK("attn_long_fwd_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_attn_long_fwd(const float* qkv, const float* e1, const float* e2, float* out, void* out_hi, void* out_lo,
                               float* probs, int64_t P, int L, int H, int head_dim, void* stream) {
  return attn_long_fwd_launch(qkv, e1, e2, out, out_hi, out_lo, probs, P, L, H, head_dim, 0, stream);
}

// the same with the decoder's attention masks (decoder.py:295-306): mask_mode 0 none, 1 causal (keys <= query), 2 anticausal
VQ_API int vqcpc_attn_long_fwd_masked(const float* qkv, const float* e1, const float* e2, float* out, float* probs, int64_t P, int L,
                                      int H, int head_dim, int mask_mode, void* stream) {
  VQ_CHECK_ARG(mask_mode >= 0 && mask_mode <= 2, "attn_long_fwd_masked: mask_mode 0 (none), 1 (causal), 2 (anticausal)");
  return attn_long_fwd_launch(qkv, e1, e2, out, nullptr, nullptr, probs, P, L, H, head_dim, mask_mode, stream);
}

// d_qkv (fp32) and / or its bf16 split are overwritten; d_e1 / d_e2 are accumulated; ds_scratch: (P, H, L, L) floats.
VQ_API int vqcpc_attn_long_bwd(const float* qkv, const float* e1, const float* e2, const float* probs, const float* d_out,
                               float* ds_scratch, float* d_qkv, void* d_qkv_hi, void* d_qkv_lo, float* d_e1, float* d_e2,
                               int64_t P, int L, int H, int head_dim, float q_scale, void* stream) {
  VQ_CHECK_ARG(head_dim == HD, "attn_long_bwd: head_dim must be 64 (got %d)", head_dim);
  VQ_CHECK_ARG(L >= 1 && L <= 1024 && H >= 1 && H <= 64, "attn_long_bwd: bad shape L=%d H=%d", L, H);
  VQ_CHECK_ARG((d_qkv || d_qkv_hi) && ((d_qkv_hi == nullptr) == (d_qkv_lo == nullptr)), "attn_long_bwd: need d_qkv and/or its split");
  VQ_CHECK_ARG(probs && ds_scratch && d_e1 && d_e2, "attn_long_bwd: null buffer");
  if (P == 0) return VQ_OK;
  VQ_CHECK_ARG(P <= 65535, "attn_long_bwd: at most 65535 sequences per call (got %lld)", (long long)P);
  cudaStream_t st = vq_stream(stream);
  __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(d_qkv_hi);
  __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(d_qkv_lo);
  dim3 grid((unsigned)vq_cdiv(L, R), (unsigned)H, (unsigned)P);
  attn_long_bwd_q_kernel<<<grid, THREADS, (size_t)(R * HD + R * L + KT * PADW) * sizeof(float), st>>>(qkv, e1, e2, probs, d_out, ds_scratch,
                                                                                          d_qkv, oh, ol, L, H, q_scale);
  VQ_CUDA_LAUNCH_CHECK("attn_long_bwd_q_kernel");
  const size_t smem_kv = (size_t)2 * L * R * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(attn_long_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 1024 * R * 4));
    attr_done = true;
  }
  attn_long_bwd_kv_kernel<<<grid, THREADS, smem_kv, st>>>(qkv, probs, ds_scratch, d_out, d_qkv, oh, ol, L, H);
  VQ_CUDA_LAUNCH_CHECK("attn_long_bwd_kv_kernel");
  attn_long_bwd_e_kernel<<<dim3((unsigned)vq_cdiv(L, ER), (unsigned)H, 2), THREADS, 0, st>>>(qkv, ds_scratch, d_e1, d_e2, P, L, H);
  VQ_CUDA_LAUNCH_CHECK("attn_long_bwd_e_kernel");
  return VQ_OK;
}

// rows x C logits; exactly one of target_idx (int64 class per row) / target_logits (rows x C, distilled) is given;
// loss (rows) and / or d_logits (rows x C, = scale * d loss / d logits) may be NULL.
VQ_API int vqcpc_softmax_ce(const float* logits, int ld, const int64_t* target_idx, const float* target_logits, int ldt,
                            float* loss, float* d_logits, int ldd, float scale, int64_t rows, int C, void* stream) {
  VQ_CHECK_ARG((target_idx != nullptr) != (target_logits != nullptr), "softmax_ce: give target_idx or target_logits");
  VQ_CHECK_ARG(C >= 1 && ld >= C && (!d_logits || ldd >= C) && (!target_logits || ldt >= C), "softmax_ce: bad leading dimension");
  if (rows == 0) return VQ_OK;
  softmax_ce_kernel<<<(unsigned)vq_cdiv(rows, 4), 128, 0, vq_stream(stream)>>>(logits, ld, target_idx, target_logits, ldt, loss,
                                                                               d_logits, ldd, scale, rows, C);
  VQ_CUDA_LAUNCH_CHECK("softmax_ce_kernel");
  return VQ_OK;
}
