// Incremental (key / value cached) autoregressive decoding for the decoder's generation loop (BASELINE
// config 5; reference: decoders/decoder.py:607-644 recomputes the full teacher-forced forward for every one of the
// num_events * num_channels tokens and samples on the host with numpy).
//
// Why a cache is exact here: the decoder's self-attention is causal (decoder.py:507) and the relative bias of BOTH attention
// flavours at entry (i, j) only involves query row i (VQCPCB/transformer/relative_maps.py asserts it for every (T, S) it builds),
// so row i of every decoder layer is a function of the code memory and of target rows <= i.  One generation step therefore
// needs ONE new row per layer: its q / k / v, attention over the cached keys 0..i, the cross-attention over the S memory rows,
// the feed-forward, and the logits of the channel being sampled.
//
// Every kernel reads the current position from DEVICE memory (`pos`), so the launch sequence of one event (num_channels token
// steps) is position independent and is replayed as one CUDA graph for all events of the window:
//     linear_small(linear_target of the feature row; the start vector at i == 0) -> per layer [ (ln_)linear_small(QKV) ->
//     attn_decode(self; appends k / v to the cache) -> linear_small(out-proj + residual) -> ln_linear_small(norm1, q) ->
//     attn_decode(cross) -> linear_small(out-proj + residual) -> ln_linear_small(norm2, FFN1, ReLU) -> linear_small(FFN2 +
//     residual) ] -> ln_linear_small(norm3, pre-softmax of the channel) -> decode_sample_embed (temperature, top-k, top-p,
//     inverse-CDF draw like numpy.random.choice; feature row of the next position; position += 1)
// = 3 + 8 * layers launches per token.  The post-LN LayerNorms run inside the consumer's staging loop (the producer wrote the
// pre-norm sum), so no row makes a separate normalisation round trip.
// The batch is 1 - 16 sequences, so the linears are weight-streaming fp32 products (exact fp32, L2 resident weights), not
// tensor-core tiles: the step is latency bound, which is what the graph is for.
#include "common.cuh"

namespace {

constexpr int HD = 64;

// Programmatic dependent launch: every kernel of the step lets its successor be scheduled at once (launch_dependents at entry)
// and only touches data a predecessor may have written after griddepcontrol.wait (= the predecessor grid has completed and its
// writes are visible).  What a kernel requests BEFORE the wait must be static for the whole generation: weight rows, biases,
// LayerNorm parameters - never the position, the activations or the caches.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Under programmatic launch a kernel is resident while its predecessors still write: anything they write (activations, caches,
// the position) is NOT read-only for the lifetime of this kernel, so it must never be fetched through the non-coherent path
// (which `const __restrict__` lets the compiler pick) nor be served from a stale L1 line -> L2-coherent loads (ld.global.cg).
__device__ __forceinline__ float4 ld_dyn4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ld_dyn(const float* p) { return __ldcg(p); }
__device__ __forceinline__ int ld_pos(const int* p) { return __ldcg(p); }

// ---------------------------------------------------------------------------------------------------------------------
// features of the row that enters the decoder at position i = *pos: the embedding of token i-1 with the channel / event
// positional embeddings of position i-1 (decoder.py:442-457, then shifted right by one :476-482).  Row 0 is the start symbol
// and is substituted after linear_target (linear_small's `row0`), so zeros are written here.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void decode_features_kernel(const int64_t* __restrict__ tokens, const int* __restrict__ pos,
                                       const float* __restrict__ tables, const int* __restrict__ table_off,
                                       const float* __restrict__ chan, const float* __restrict__ evt, float* __restrict__ feat,
                                       int T, int C, int up, int E, int pe, int pos_mode) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, i = ld_pos(pos), F = pos_mode ? E + pe : E + 2 * pe;
  float* f = feat + (size_t)b * F;
  if (i <= 0 || i >= T + 1) {
    for (int c = threadIdx.x; c < F; c += blockDim.x) f[c] = 0.f;
    return;
  }
  // pos_mode 0: [embedding | channel embedding | event-in-code embedding] (relative transformer, decoder.py:449-457);
  // pos_mode 1: [embedding | positional embedding of the absolute position] (chan = the (T, pe) table, decoder.py:458-466)
  const int t = i - 1, ch = t % C, ev = (t % up) / C, pidx = pos_mode ? t : ch;
  const int64_t tok = __ldcg(tokens + (size_t)b * T + t);
  const float* row = tables + ((size_t)table_off[ch] + tok) * E;
  for (int c = threadIdx.x; c < F; c += blockDim.x)
    f[c] = c < E ? row[c] : (c < E + pe ? chan[pidx * pe + (c - E)] : evt[ev * pe + (c - E - pe)]);
}

// ---------------------------------------------------------------------------------------------------------------------
// out[b, n] = epilogue( sum_k x^[b, k] W[n, k] + bias[n] ),  b < M <= 16.  One warp per output column: the W row lives in
// registers (KV float4 per lane), the M input rows are read straight through L1 (they are a few KB and every warp reads all of
// them), RB rows in flight.  No shared memory, no block barrier: the kernel is one dependent chain of
//   W / LayerNorm parameters (independent of the producer) -> rows -> warp reductions -> store,
// which is what bounds a launch at these sizes.
// LN: x^ = nn.LayerNorm(K)(x) computed per warp in registers (post-LN blocks, transformer_custom.py:283-289: the producer wrote
// the pre-norm sum; biased variance, eps inside the square root, two passes like layernorm_fwd_kernel); warp 0 of CTA 0 also
// stores the normalised rows (ln_out), the residual input of a later linear.
// Epilogue order = the layer code's: bias, scale of the first `scale_cols` columns (the q third of the in-projection,
// multihead_attention_custom.py:247), ReLU, residual.  row0 != NULL and *pos == 0: the whole row is replaced by row0 (the
// start-of-sequence vector, decoder.py:476-482).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

template <int KV, int RB, bool LN>
__global__ void __launch_bounds__(128)
linear_small_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
                    const float* __restrict__ residual, int ldr, float* __restrict__ out, int ldo, int M, int N, int K, int relu,
                    float scale, int scale_cols, const float* __restrict__ row0, const int* __restrict__ pos,
                    const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta, float ln_eps,
                    float* __restrict__ ln_out) {
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + warp;
  if (n >= N) return;
  const int K4 = K >> 2;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 w[KV], g[LN ? KV : 1], be[LN ? KV : 1];
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    const int c = i * 32 + lane;
    w[i] = c < K4 ? __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K) + c) : zero4;
    if (LN) {
      g[i] = c < K4 ? __ldg(reinterpret_cast<const float4*>(ln_gamma) + c) : zero4;
      be[i] = c < K4 ? __ldg(reinterpret_cast<const float4*>(ln_beta) + c) : zero4;
    }
  }
  const float bn = bias ? __ldg(bias + n) : 0.f;
  pdl_wait();                                      // above: static parameters only
  const bool use_row0 = row0 != nullptr && ld_pos(pos) == 0;
  const bool writer = LN && ln_out != nullptr && blockIdx.x == 0 && warp == 0;
  const float invK = 1.f / (float)K;
  for (int b0 = 0; b0 < M; b0 += RB) {
    float4 xr[RB][KV];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        const int c = i * 32 + lane;
        xr[r][i] = (b0 + r < M && c < K4) ? ld_dyn4(x + (size_t)(b0 + r) * ldx + 4 * c) : zero4;
      }
    float acc[RB];
    if (LN) {
      float mu[RB], rs[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < KV; ++i) t += (xr[r][i].x + xr[r][i].y) + (xr[r][i].z + xr[r][i].w);
        mu[r] = t;
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) mu[r] = warp_sum(mu[r]) * invK;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < KV; ++i) {
          if (i * 32 + lane < K4) {
            const float a0 = xr[r][i].x - mu[r], a1 = xr[r][i].y - mu[r], a2 = xr[r][i].z - mu[r], a3 = xr[r][i].w - mu[r];
            t += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
          }
        }
        rs[r] = t;
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) rs[r] = 1.0f / sqrtf(warp_sum(rs[r]) * invK + ln_eps);
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int i = 0; i < KV; ++i) {
          float4 o;
          o.x = (xr[r][i].x - mu[r]) * rs[r] * g[i].x + be[i].x;
          o.y = (xr[r][i].y - mu[r]) * rs[r] * g[i].y + be[i].y;
          o.z = (xr[r][i].z - mu[r]) * rs[r] * g[i].z + be[i].z;
          o.w = (xr[r][i].w - mu[r]) * rs[r] * g[i].w + be[i].w;
          xr[r][i] = o;
          const int c = i * 32 + lane;
          if (writer && b0 + r < M && c < K4) *(reinterpret_cast<float4*>(ln_out + (size_t)(b0 + r) * K) + c) = o;
        }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < KV; ++i) t += dot4(w[i], xr[r][i]);
      acc[r] = t;
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = warp_sum(acc[r]);
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int b = b0 + r;
      if (lane == r && b < M) {
        float v = acc[r] + bn;
        if (n < scale_cols) v *= scale;
        if (relu) v = fmaxf(v, 0.f);
        if (residual) v += ld_dyn(residual + (size_t)b * ldr + n);
        if (use_row0) v = row0[n];
        out[(size_t)b * ldo + n] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same product for 3 <= M <= 32 rows: the M input rows are staged in shared memory by M bulk (TMA) copies that one thread
// issues at kernel entry - ONE round of memory latency for the whole input, overlapped with every warp's W-row load - instead
// of one round per group of rows.  LN: warp w normalises rows w, w + 4, .. in place once per CTA.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ls_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if (it > (1u << 26)) { printf("vqcpc linear_small: mbarrier wait timed out (block %d)\n", blockIdx.x); __trap(); }
  }
}

template <int KV, int MT, bool LN>
__global__ void __launch_bounds__(128)
linear_staged_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
                     const float* __restrict__ residual, int ldr, float* __restrict__ out, int ldo, int M, int N, int K, int relu,
                     float scale, int scale_cols, const float* __restrict__ row0, const int* __restrict__ pos,
                     const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta, float ln_eps,
                     float* __restrict__ ln_out) {
  extern __shared__ __align__(128) unsigned char ls_smem[];   // [0, 8) mbarrier, [16, 16 + M * K * 4) rows
  float* xs = reinterpret_cast<float*>(ls_smem + 16);
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(ls_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int n = blockIdx.x * 4 + warp;
  const bool active = n < N;
  const int K4 = K >> 2;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 w[KV];
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    const int c = i * 32 + lane;
    w[i] = (active && c < K4) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K) + c) : zero4;
  }
  float4 g[LN ? KV : 1], be[LN ? KV : 1];                          // requested before the wait, like the W row
  if (LN) {
#pragma unroll
    for (int i = 0; i < KV; ++i) {
      const int c = i * 32 + lane;
      g[i] = c < K4 ? __ldg(reinterpret_cast<const float4*>(ln_gamma) + c) : zero4;
      be[i] = c < K4 ? __ldg(reinterpret_cast<const float4*>(ln_beta) + c) : zero4;
    }
  }
  const float bn = (active && bias) ? __ldg(bias + n) : 0.f;
  pdl_wait();                                      // above: static parameters only; the rows are the predecessor's output
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(M * K * 4)) : "memory");
    for (int b = 0; b < M; ++b)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"((uint32_t)__cvta_generic_to_shared(xs + (size_t)b * K)), "l"(x + (size_t)b * ldx), "r"((uint32_t)(K * 4)), "r"(bar)
                   : "memory");
  }
  const float resv = (active && residual && lane < M) ? ld_dyn(residual + (size_t)lane * ldr + n) : 0.f;
  const bool use_row0 = row0 != nullptr && ld_pos(pos) == 0;
  __syncthreads();                                 // the barrier initialisation is visible to the waiting threads
  ls_mbar_wait(bar, 0);
  if (LN) {
    // warp w normalises rows w, w + 4, w + 8, ... - all of them at once (independent reduction chains interleave)
    constexpr int NR = (MT + 3) / 4;
    const float invK = 1.f / (float)K;
    float4 v[NR][KV];
    float mu[NR], rs[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int bb = warp + 4 * r;
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < KV; ++i) {
        const int c = i * 32 + lane;
        v[r][i] = (bb < M && c < K4) ? reinterpret_cast<const float4*>(xs + (size_t)bb * K)[c] : zero4;
        t += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
      }
      mu[r] = t;
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) mu[r] = warp_sum(mu[r]) * invK;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < KV; ++i)
        if (i * 32 + lane < K4) {
          const float a0 = v[r][i].x - mu[r], a1 = v[r][i].y - mu[r], a2 = v[r][i].z - mu[r], a3 = v[r][i].w - mu[r];
          t += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
        }
      rs[r] = t;
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) rs[r] = 1.0f / sqrtf(warp_sum(rs[r]) * invK + ln_eps);
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int bb = warp + 4 * r;
      if (bb < M) {
#pragma unroll
        for (int i = 0; i < KV; ++i) {
          const int c = i * 32 + lane;
          if (c < K4) {
            float4 o;
            o.x = (v[r][i].x - mu[r]) * rs[r] * g[i].x + be[i].x;
            o.y = (v[r][i].y - mu[r]) * rs[r] * g[i].y + be[i].y;
            o.z = (v[r][i].z - mu[r]) * rs[r] * g[i].z + be[i].z;
            o.w = (v[r][i].w - mu[r]) * rs[r] * g[i].w + be[i].w;
            reinterpret_cast<float4*>(xs + (size_t)bb * K)[c] = o;
            if (ln_out && blockIdx.x == 0) *(reinterpret_cast<float4*>(ln_out + (size_t)bb * K) + c) = o;
          }
        }
      }
    }
    __syncthreads();
  }
  if (!active) return;
  float acc[MT];
#pragma unroll
  for (int b = 0; b < MT; ++b) acc[b] = 0.f;
#pragma unroll
  for (int i = 0; i < KV; ++i) {
    const int c = i * 32 + lane;
    if (c < K4) {
#pragma unroll
      for (int b = 0; b < MT; ++b)
        if (b < M) acc[b] += dot4(w[i], reinterpret_cast<const float4*>(xs + (size_t)b * K)[c]);
    }
  }
#pragma unroll
  for (int b = 0; b < MT; ++b) acc[b] = warp_sum(acc[b]);
#pragma unroll
  for (int b = 0; b < MT; ++b) {
    if (lane == b && b < M) {
      float v = acc[b] + bn;
      if (n < scale_cols) v *= scale;
      if (relu) v = fmaxf(v, 0.f);
      if (residual) v += resv;
      if (use_row0) v = row0[n];
      out[(size_t)b * ldo + n] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// attention of ONE query row per (sequence, head) at position i = *pos over n keys (causal: n = i + 1, else n = Lk), with the
// reference's relative bias gathered through row i of the index maps (see attn_map.cu).  k_new / v_new != NULL: the new key and
// value are appended to the cache at row i (decoder self-attention).
// The keys are split over the NS CTAs of a thread-block cluster (at batch 1 there are only H (sequence, head) pairs for 148
// SMs): CTA r scores keys [r * chunk, (r + 1) * chunk), keeps its own maximum / exp-sum / weighted value sum, and CTA 0 merges
// the NS partial results from the peers' shared memory (DSMEM) - the usual split-softmax identity.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AD_THREADS = 128;
constexpr int AD_MAX_SPLIT = 8;

template <bool PREFETCH_V>   // true: latency-optimised (few CTAs in flight, 255 registers); false: 3 CTAs per SM for full grids
__global__ void __launch_bounds__(AD_THREADS, PREFETCH_V ? 1 : 3)
attn_decode_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k_new, const float* __restrict__ v_new, int ldn,
                   float* kcache, float* vcache, int ldc, int Lmax, const float* __restrict__ e1, const float* __restrict__ e2,
                   const int* __restrict__ map1, const int* __restrict__ map2, const float* __restrict__ mask, int Lk,
                   const int* __restrict__ pos, int causal, float* __restrict__ out, int ldo, int chunk_max) {
  extern __shared__ __align__(16) float sm[];      // [64] q, [8][64] partial outputs, [8] reductions, [72] result, [8][72] gather, [chunk] scores
  float* qs = sm;
  float* part = qs + HD;                           // [8][64]
  float* red = part + 8 * HD;
  float* result = red + 8;                         // [0..63] sum_j e^(s_j - m) v_j, [64] m, [65] z
  float* gather = result + 72;                     // [8][72]: the results of all CTAs of the cluster (used in CTA 0)
  float* sc = gather + AD_MAX_SPLIT * 72;
  uint32_t crank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  pdl_launch_dependents();
  if (csize > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  pdl_wait();
  const int h = blockIdx.x, tid = threadIdx.x, i = ld_pos(pos);
  const int64_t p = blockIdx.y;
  const int n = causal ? min(i + 1, Lk) : Lk;
  int chunk = (((n + (int)csize - 1) / (int)csize) + 7) & ~7;
  chunk = min(chunk, chunk_max);
  const int j_lo = min(n, (int)crank * chunk), j_hi = min(n, j_lo + chunk), cnt = j_hi - j_lo;
  float* kc = kcache + (size_t)p * Lmax * ldc + h * HD;
  float* vc = vcache + (size_t)p * Lmax * ldc + h * HD;
  const float* kn = k_new ? k_new + (size_t)p * ldn + h * HD : nullptr;
  const float* vn = v_new ? v_new + (size_t)p * ldn + h * HD : nullptr;
  // Everything that only depends on the position is requested BEFORE the query is staged: the map / mask entries and the key row
  // of this thread's first key, and the value rows of the first accumulation pass.  With one key per thread (chunk <= 128) the
  // kernel is then three rounds of memory latency: position -> {q, maps, keys, values} -> relative-table rows.
  const float* e1h = e1 ? e1 + (size_t)h * Lk * HD : nullptr;
  const float* e2h = e2 ? e2 + (size_t)h * Lk * HD : nullptr;
  const int* m1row = map1 ? map1 + (size_t)i * Lk : nullptr;
  const int* m2row = map2 ? map2 + (size_t)i * Lk : nullptr;
  const float* mrow = mask ? mask + (size_t)i * Lk : nullptr;
  int j = j_lo + tid;
  int m1 = -1, m2 = -1;
  float add = 0.f;
  float4 kreg[HD / 4];
  if (j < j_hi) {
    m1 = m1row ? m1row[j] : -1;
    m2 = m2row ? m2row[j] : -1;
    add = mrow ? mrow[j] : 0.f;
    const float4* kr = reinterpret_cast<const float4*>((kn && j == i) ? kn : kc + (size_t)j * ldc);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) kreg[c] = __ldcg(kr + c);
  }
  const int sub = tid & 15, grp = tid >> 4;
  float4 vv[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int jj = t * 8 + grp, ja = j_lo + jj;
    const float* vrow = (vn && ja == i) ? vn : vc + (size_t)ja * ldc;
    if (PREFETCH_V) vv[t] = jj < cnt ? ld_dyn4(vrow + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tid < HD) {
    qs[tid] = ld_dyn(q + (size_t)p * ldq + h * HD + tid);
    if (kn && crank == 0) {                        // the appended row is read from k_new / v_new, never from the cache
      kc[(size_t)i * ldc + tid] = ld_dyn(kn + tid);
      vc[(size_t)i * ldc + tid] = ld_dyn(vn + tid);
    }
  }
  __syncthreads();
  // scores: ONE KEY PER THREAD - its 16 + 16 + 16 float4 loads (key row, e1 row, e2 row) are independent of each other (the rows
  // are whole 128-byte lines, nothing is over-fetched)
  const float4* q4 = reinterpret_cast<const float4*>(qs);
  while (j < j_hi) {
    const float4* r1 = m1 >= 0 ? reinterpret_cast<const float4*>(e1h + (size_t)(m1 % Lk) * HD) : nullptr;
    const float4* r2 = m2 >= 0 ? reinterpret_cast<const float4*>(e2h + (size_t)(m2 % Lk) * HD) : nullptr;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) s0 += dot4(q4[c], kreg[c]);
    if (r1) {
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) s1 += dot4(q4[c], __ldg(r1 + c));
    } else if (m1 == -2) add += -100.f;
    if (r2) {
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) s2 += dot4(q4[c], __ldg(r2 + c));
    } else if (m2 == -2) add += -100.f;
    sc[j - j_lo] = (s0 + s1 + s2) + add;
    j += AD_THREADS;
    if (j < j_hi) {
      m1 = m1row ? m1row[j] : -1;
      m2 = m2row ? m2row[j] : -1;
      add = mrow ? mrow[j] : 0.f;
      const float4* kr = reinterpret_cast<const float4*>((kn && j == i) ? kn : kc + (size_t)j * ldc);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) kreg[c] = __ldcg(kr + c);
    }
  }
  __syncthreads();
  // local softmax statistics
  const int warp = tid >> 5, lane = tid & 31;
  float mx = -INFINITY;
  for (int j = tid; j < cnt; j += AD_THREADS) mx = fmaxf(mx, sc[j]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float z = 0.f;
  for (int j = tid; j < cnt; j += AD_THREADS) {
    const float ev = mx == -INFINITY ? 0.f : __expf(sc[j] - mx);
    sc[j] = ev;
    z += ev;
  }
  z = warp_sum(z);
  if (lane == 0) red[4 + warp] = z;
  __syncthreads();
  // o[d] = sum_j e_j v[j, d]: 16 lanes x float4 per value row, 8 rows per pass, 8 passes in flight (the first 64 rows were
  // requested at kernel entry)
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = 0; j0 < cnt; j0 += 64) {
    float pw[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int jj = j0 + t * 8 + grp, ja = j_lo + jj;
      const bool ok = jj < cnt;
      if (!PREFETCH_V || j0 > 0) {
        const float* vrow = (vn && ja == i) ? vn : vc + (size_t)ja * ldc;
        vv[t] = ok ? ld_dyn4(vrow + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      pw[t] = ok ? sc[jj] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      o.x += pw[t] * vv[t].x;
      o.y += pw[t] * vv[t].y;
      o.z += pw[t] * vv[t].z;
      o.w += pw[t] * vv[t].w;
    }
  }
  reinterpret_cast<float4*>(part)[grp * 16 + sub] = o;          // part[8][64]
  __syncthreads();
  if (tid < HD) {
    float t = 0.f;
#pragma unroll
    for (int gph = 0; gph < 8; ++gph) t += part[gph * HD + tid];
    result[tid] = t;
  }
  if (tid == 0) {
    result[64] = mx;
    result[65] = red[4] + red[5] + red[6] + red[7];
  }
  if (csize == 1) {
    __syncthreads();
    if (tid < HD) out[(size_t)p * ldo + h * HD + tid] = result[tid] / result[65];
    return;
  }
  // every CTA pushes its 66 numbers into CTA 0's shared memory (remote stores, nobody waits on a remote load), one cluster
  // barrier, CTA 0 merges locally.  The arrive at kernel entry + this wait guarantee CTA 0 is running before it is written to.
  __syncthreads();
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tid < 66) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((uint32_t)__cvta_generic_to_shared(gather + crank * 72 + tid)), "r"(0));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(result[tid]) : "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (crank == 0 && tid < HD) {
    float M = -INFINITY;
    for (uint32_t r = 0; r < csize; ++r) M = fmaxf(M, gather[r * 72 + 64]);
    float Z = 0.f, O = 0.f;
    for (uint32_t r = 0; r < csize; ++r) {
      const float mr = gather[r * 72 + 64];
      if (mr == -INFINITY) continue;
      const float w = __expf(mr - M);
      Z += gather[r * 72 + 65] * w;
      O += gather[r * 72 + tid] * w;
    }
    out[(size_t)p * ldo + h * HD + tid] = O / Z;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// one token per sequence from a row of logits: logits / temperature, top-k, top-p (VQCPCB/utils.py:100-128), softmax, then the
// draw of numpy.random.choice(V, p=p) (decoder.py:636-644): cdf = cumsum(double(p)); cdf /= cdf[-1]; index of the first
// cdf > u, with u the uniform the caller supplies for this (position, sequence) - the same stream numpy would have consumed.
// greedy: argmax of the logits (lowest index on ties).  The token is written to tokens[b, *pos].
// ---------------------------------------------------------------------------------------------------------------------
// whole block; returns the drawn token in every thread (broadcast through shared memory)
__device__ int sample_row(const float* __restrict__ lg, int V, float temperature, int top_k, float top_p, double uniform, int greedy,
                          float* smf) {
  float* x = smf;                                  // [V] logits, [V] sorted, int [V] order (sorted position -> index), double [V] cdf
  float* srt = smf + V;
  int* order = reinterpret_cast<int*>(smf + 2 * V);
  double* cdf = reinterpret_cast<double*>(smf + 3 * V + (V & 1));
  __shared__ int s_result;
  const int tid = threadIdx.x;
  for (int c = tid; c < V; c += 256) x[c] = ld_dyn(lg + c) / temperature;
  __syncthreads();
  // rank = position in the descending order (ties: lower index first)
  for (int c = tid; c < V; c += 256) {
    const float v = x[c];
    int r = 0;
    for (int j = 0; j < V; ++j) {
      const float w = x[j];
      r += (w > v) || (w == v && j < c);
    }
    srt[r] = v;
    order[r] = c;
  }
  __syncthreads();
  // the exponentials and their sums are evaluated by the whole block (fixed-shape tree sums: deterministic); only the
  // order-dependent scans (top-p cumulative sum, cdf) run on one thread, over precomputed values
  __shared__ float s_red[8];
  const int warp = tid >> 5, lane = tid & 31;
  if (!greedy) {
    const int k = min(top_k, V);
    if (k > 0) {                                             // utils.py:112-115: logits < k-th largest removed (ties stay)
      const float kth = srt[k - 1];
      __syncthreads();
      for (int r = k + tid; r < V; r += 256)
        if (srt[r] < kth) srt[r] = -INFINITY;
      __syncthreads();
    }
    const float mx = srt[0];
    float* ex = reinterpret_cast<float*>(cdf);               // [V] scratch (the cdf is built later, by thread 0 only)
    if (top_p > 0.f) {                                       // utils.py:117-127
      float zp = 0.f;
      for (int r = tid; r < V; r += 256) {
        const float e = expf(srt[r] - mx);
        ex[r] = e;
        zp += e;
      }
      zp = warp_sum(zp);
      if (lane == 0) s_red[warp] = zp;
      __syncthreads();
      float zt = 0.f;
      for (int wi = 0; wi < 8; ++wi) zt += s_red[wi];
      for (int r = tid; r < V; r += 256) ex[r] = ex[r] / zt;
      __syncthreads();
      if (tid == 0) {
        float cum = 0.f;
        bool prev_above = false;
        for (int r = 0; r < V; ++r) {
          cum += ex[r];
          const bool above = cum > top_p;
          if (prev_above) srt[r] = -INFINITY;                // shifted right by one: the first token above the threshold stays
          prev_above = above;
        }
      }
      __syncthreads();
    }
    // softmax of the filtered logits in ORIGINAL index order
    for (int r = tid; r < V; r += 256) x[order[r]] = srt[r];
    __syncthreads();
    float zp = 0.f;
    for (int c = tid; c < V; c += 256) {
      const float e = expf(x[c] - mx);
      x[c] = e;
      zp += e;
    }
    zp = warp_sum(zp);
    if (lane == 0) s_red[warp] = zp;
    __syncthreads();
    float zt = 0.f;
    for (int wi = 0; wi < 8; ++wi) zt += s_red[wi];
    for (int c = tid; c < V; c += 256) x[c] = x[c] / zt;     // float32 probabilities, as torch.softmax hands them to numpy
    __syncthreads();
    if (tid == 0) {
      double run = 0.0;                                      // double cdf like numpy
      for (int c = 0; c < V; ++c) {
        run += (double)x[c];
        cdf[c] = run;
      }
      s_result = V - 1;
    }
    __syncthreads();
    const double uu = uniform * cdf[V - 1];                  // cdf /= cdf[-1]  <=>  compare with u * cdf[-1]
    for (int c = tid; c < V; c += 256)                       // first index with cdf > u (the cdf is non-decreasing: one winner)
      if (cdf[c] > uu && (c == 0 || !(cdf[c - 1] > uu))) s_result = c;
  } else if (tid == 0) {
    s_result = order[0];
  }
  __syncthreads();
  return s_result;
}

__global__ void __launch_bounds__(256)
sample_step_kernel(const float* __restrict__ logits, int ld, int V, float temperature, int top_k, float top_p,
                   const double* __restrict__ u, int u_stride, int64_t* __restrict__ tokens, int T, const int* __restrict__ pos,
                   int greedy, int B) {
  extern __shared__ __align__(16) float smf[];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, i = ld_pos(pos);
  if (i < 0 || i >= T) return;
  const int tok = sample_row(logits + (size_t)b * ld, V, temperature, top_k, top_p, greedy ? 0.0 : u[(size_t)i * u_stride + b], greedy, smf);
  if (threadIdx.x == 0) tokens[(size_t)b * T + i] = tok;
}

// sample_step + the feature row of the NEXT position (decode_features of position *pos + 1: the token just drawn with the
// positional embeddings of position *pos) + the position increment by the last CTA to finish (every CTA reads *pos before it
// arrives on `done`, so the increment cannot overtake a reader) - one launch instead of three per token.
__global__ void __launch_bounds__(256)
sample_embed_advance_kernel(const float* __restrict__ logits, int ld, int V, float temperature, int top_k, float top_p,
                            const double* __restrict__ u, int u_stride, int64_t* __restrict__ tokens, int T, int* pos, int greedy,
                            const float* __restrict__ tables, const int* __restrict__ table_off, const float* __restrict__ chan,
                            const float* __restrict__ evt, float* __restrict__ feat, int C, int up, int E, int pe, int* done,
                            int pos_mode) {
  extern __shared__ __align__(16) float smf[];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, i = *reinterpret_cast<volatile int*>(pos);
  if (i >= 0 && i < T) {
    const int tok = sample_row(logits + (size_t)b * ld, V, temperature, top_k, top_p, greedy ? 0.0 : u[(size_t)i * u_stride + b], greedy, smf);
    if (threadIdx.x == 0) tokens[(size_t)b * T + i] = tok;
    const int ch = i % C, ev = (i % up) / C, F = pos_mode ? E + pe : E + 2 * pe, pidx = pos_mode ? i : ch;
    const float* row = tables + ((size_t)table_off[ch] + tok) * E;
    for (int c = threadIdx.x; c < F; c += blockDim.x)
      feat[(size_t)b * F + c] = c < E ? row[c] : (c < E + pe ? chan[pidx * pe + (c - E)] : evt[ev * pe + (c - E - pe)]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
      *done = 0;
      *pos = i + 1;
    }
  }
}

// out[b, :] = nn.LayerNorm(d)(y[b, :]) + add[b, *pos, :]: the norm1 of an ALIGNED decoder layer followed by the residual of its
// "cross-attention", which is an MLP of the aligned memory event and therefore a per-position constant of the whole generation
// (transformer_custom.py:448-472), precomputed once per window as add (B, T, d).  One CTA per row.
__global__ void __launch_bounds__(128)
decode_ln_add_kernel(const float* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     const float* __restrict__ add, int T, const int* __restrict__ pos, float* __restrict__ out, int d) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i = min(max(ld_pos(pos), 0), T - 1);
  const float* row = y + (size_t)b * d;
  const float* arow = add + ((size_t)b * T + i) * d;
  float v[8], a[8];                                // d <= 1024
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = k * 128 + tid;
    v[k] = c < d ? ld_dyn(row + c) : 0.f;
    a[k] = c < d ? ld_dyn(arow + c) : 0.f;
    t += v[k];
  }
  t = warp_sum(t);
  if (lane == 0) red[warp] = t;
  __syncthreads();
  const float mu = (red[0] + red[1] + red[2] + red[3]) / d;
  t = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k * 128 + tid < d) t += (v[k] - mu) * (v[k] - mu);
  t = warp_sum(t);
  if (lane == 0) red[4 + warp] = t;
  __syncthreads();
  const float rs = 1.0f / sqrtf((red[4] + red[5] + red[6] + red[7]) / d + eps);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = k * 128 + tid;
    if (c < d) out[(size_t)b * d + c] = (v[k] - mu) * rs * gamma[c] + beta[c] + a[k];
  }
}

__global__ void elu_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[t];
    out[t] = v > 0.f ? v : expm1f(v);              // nn.ELU(alpha = 1)
  }
}

__global__ void decode_advance_kernel(int* pos, int delta) {
  pdl_launch_dependents();
  pdl_wait();
  *pos += delta;
}

static int g_decode_pdl = 1;

// <<<grid, block, smem, stream>>> with the programmatic-stream-serialization attribute (and an optional cluster along z)
template <typename... KArgs, typename... Args>
cudaError_t launch_step(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_z, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned na = 0;
  if (cluster_z > 0) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = cluster_z;
    ++na;
  }
  if (g_decode_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace

VQ_API int vqcpc_decode_set_pdl(int on) {
  g_decode_pdl = on != 0;
  return VQ_OK;
}

VQ_API int vqcpc_decode_features(const int64_t* tokens, const int* pos, const float* tables, const int* table_off, const float* chan,
                                 const float* evt, float* feat, int B, int T, int C, int total_upscaling, int E, int pe, int pos_mode,
                                 void* stream) {
  VQ_CHECK_ARG(B >= 1 && T >= 1 && C >= 1 && total_upscaling % C == 0 && E >= 1 && pe >= 0, "decode_features: bad shape");
  VQ_CUDA_CALL(launch_step(decode_features_kernel, dim3(B), dim3(64), 0, vq_stream(stream), 0, tokens, pos, tables, table_off, chan, evt,
                           feat, T, C, total_upscaling, E, pe, pos_mode));
  return VQ_OK;
}

static int linear_small_launch(const float* x, int ldx, const float* W, const float* bias, const float* residual, int ldr, float* out,
                               int ldo, int M, int N, int K, int relu, float scale, int scale_cols, const float* row0, const int* pos,
                               const float* ln_gamma, const float* ln_beta, float ln_eps, float* ln_out, void* stream) {
  VQ_CHECK_ARG(M >= 1 && M <= 32, "linear_small: 1 <= M <= 32 rows per call (got %d)", M);
  VQ_CHECK_ARG(!(ln_gamma != nullptr && M > 16 && K > 512), "linear_small: LayerNorm on load with more than 16 rows needs K <= 512");
  VQ_CHECK_ARG(N >= 1 && K >= 4 && (K % 4) == 0 && (ldx % 4) == 0 && K <= 2048, "linear_small: bad shape N=%d K=%d ldx=%d (K <= 2048)", N, K, ldx);
  VQ_CHECK_ARG(row0 == nullptr || pos != nullptr, "linear_small: row0 needs the position pointer");
  VQ_CHECK_ARG((ln_gamma == nullptr) == (ln_beta == nullptr) && (ln_out == nullptr || ln_gamma != nullptr), "linear_small: LayerNorm arguments");
  VQ_CHECK_ARG(ln_gamma == nullptr || K <= 1024, "linear_small: LayerNorm on load needs K <= 1024");
  const unsigned grid = (unsigned)vq_cdiv(N, 4);
  cudaStream_t st = vq_stream(stream);
  const bool ln = ln_gamma != nullptr;
  if (M >= 3 && K <= 1024) {
    const size_t smem = 16 + (size_t)M * K * sizeof(float);
    static bool attr_done = false;
#define VQ_ST_ATTR(KVV, MTV, LNV) \
  VQ_CUDA_CALL(cudaFuncSetAttribute(linear_staged_kernel<KVV, MTV, LNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 + MTV * 128 * KVV * 4))
    if (!attr_done) {                              // the instantiations whose rows exceed the default 48 KB
      VQ_ST_ATTR(8, 16, true);
      VQ_ST_ATTR(8, 16, false);
      VQ_ST_ATTR(8, 32, false);
      VQ_ST_ATTR(4, 32, true);
      VQ_ST_ATTR(4, 32, false);
      attr_done = true;
    }
#undef VQ_ST_ATTR
#define VQ_ST(KVV, MTV, LNV)                                                                                                            \
  VQ_CUDA_CALL(launch_step(linear_staged_kernel<KVV, MTV, LNV>, dim3(grid), dim3(128), smem, st, 0, x, ldx, W, bias, residual, ldr, out, ldo, \
                           M, N, K, relu, scale, scale_cols, row0, pos, ln_gamma, ln_beta, ln_eps, ln_out))
#define VQ_ST_M(KVV, LNV)                    \
  do {                                       \
    if (M <= 4) VQ_ST(KVV, 4, LNV);          \
    else if (M <= 8) VQ_ST(KVV, 8, LNV);     \
    else if (M <= 16) VQ_ST(KVV, 16, LNV);   \
    else VQ_ST(KVV, 32, LNV);                \
  } while (0)
#define VQ_ST_K(KVV)                         \
  do {                                       \
    if (ln) VQ_ST_M(KVV, true);              \
    else VQ_ST_M(KVV, false);                \
  } while (0)
    if (K <= 128) VQ_ST_K(1);
    else if (K <= 512) VQ_ST_K(4);
    else if (ln) {                             // K <= 1024 with LayerNorm: at most 16 rows (checked above)
      if (M <= 4) VQ_ST(8, 4, true);
      else if (M <= 8) VQ_ST(8, 8, true);
      else VQ_ST(8, 16, true);
    } else VQ_ST_M(8, false);
#undef VQ_ST_K
#undef VQ_ST_M
#undef VQ_ST
    return VQ_OK;
  }
  const int rb = M == 1 ? 1 : (M == 2 || K > 512) ? 2 : 4;          // rows in flight per warp (register budget: K / 128 float4 each)
#define VQ_LS(KVV, RBV, LNV)                                                                                                          \
  VQ_CUDA_CALL(launch_step(linear_small_kernel<KVV, RBV, LNV>, dim3(grid), dim3(128), 0, st, 0, x, ldx, W, bias, residual, ldr, out, ldo, M, \
                           N, K, relu, scale, scale_cols, row0, pos, ln_gamma, ln_beta, ln_eps, ln_out))
#define VQ_LS_R(KVV, LNV)                    \
  do {                                       \
    if (rb == 1) VQ_LS(KVV, 1, LNV);         \
    else if (rb == 2) VQ_LS(KVV, 2, LNV);    \
    else VQ_LS(KVV, 4, LNV);                 \
  } while (0)
#define VQ_LS_K(KVV)                         \
  do {                                       \
    if (ln) VQ_LS_R(KVV, true);              \
    else VQ_LS_R(KVV, false);                \
  } while (0)
  if (K <= 128) VQ_LS_K(1);
  else if (K <= 512) VQ_LS_K(4);
  else if (K <= 1024) {
    if (ln) { if (rb == 1) VQ_LS(8, 1, true); else VQ_LS(8, 2, true); }
    else { if (rb == 1) VQ_LS(8, 1, false); else VQ_LS(8, 2, false); }
  } else {
    if (rb == 1) VQ_LS(16, 1, false);
    else VQ_LS(16, 2, false);
  }
#undef VQ_LS_R
#undef VQ_LS_K
#undef VQ_LS
  return VQ_OK;
}

VQ_API int vqcpc_linear_small(const float* x, int ldx, const float* W, const float* bias, const float* residual, int ldr, float* out,
                              int ldo, int M, int N, int K, int relu, float scale, int scale_cols, const float* row0, const int* pos,
                              void* stream) {
  return linear_small_launch(x, ldx, W, bias, residual, ldr, out, ldo, M, N, K, relu, scale, scale_cols, row0, pos, nullptr, nullptr, 0.f,
                             nullptr, stream);
}

VQ_API int vqcpc_ln_linear_small(const float* x, int ldx, const float* ln_gamma, const float* ln_beta, float ln_eps, float* ln_out,
                                 const float* W, const float* bias, const float* residual, int ldr, float* out, int ldo, int M, int N,
                                 int K, int relu, float scale, int scale_cols, void* stream) {
  VQ_CHECK_ARG(ln_gamma != nullptr && ln_beta != nullptr, "ln_linear_small: LayerNorm parameters missing");
  return linear_small_launch(x, ldx, W, bias, residual, ldr, out, ldo, M, N, K, relu, scale, scale_cols, nullptr, nullptr, ln_gamma, ln_beta,
                             ln_eps, ln_out, stream);
}

static int g_attn_decode_split = 0;   // 0 = automatic
VQ_API int vqcpc_attn_decode_set_split(int ns) {
  VQ_CHECK_ARG(ns == 0 || ns == 1 || ns == 2 || ns == 4 || ns == 8, "attn_decode_set_split: 0 (auto), 1, 2, 4 or 8");
  g_attn_decode_split = ns;
  return VQ_OK;
}

VQ_API int vqcpc_attn_decode(const float* q, int ldq, const float* k_new, const float* v_new, int ldn, float* kcache, float* vcache,
                             int ldc, int Lmax, const float* e1, const float* e2, const int* map1, const int* map2, const float* mask,
                             int Lk, const int* pos, int causal, float* out, int ldo, int64_t P, int H, int head_dim, void* stream) {
  VQ_CHECK_ARG(head_dim == HD, "attn_decode: head_dim must be 64 (got %d)", head_dim);
  VQ_CHECK_ARG(Lk >= 1 && Lk <= 8192 && Lmax >= Lk && H >= 1 && P >= 1 && P <= 65535, "attn_decode: bad shape Lk=%d Lmax=%d H=%d", Lk, Lmax, H);
  VQ_CHECK_ARG((map1 == nullptr) == (e1 == nullptr) && (map2 == nullptr) == (e2 == nullptr), "attn_decode: a map needs its table");
  VQ_CHECK_ARG((k_new == nullptr) == (v_new == nullptr) && (ldc % 4) == 0 && (k_new == nullptr || (ldn % 4) == 0), "attn_decode: bad cache arguments");
  VQ_CHECK_ARG(pos != nullptr, "attn_decode: position pointer missing");
  // split the keys over a cluster: at most one CTA per 128 keys (one key per thread = one round of memory latency per chunk), and
  // about 256 CTAs in flight for the whole launch - more only adds waves of latency-bound CTAs (measured on B200, 1024 keys, 8
  // heads: batch 1 best with 8 splits, batch 16 with 2)
  int ns_max = 1;
  while (ns_max < AD_MAX_SPLIT && Lk > ns_max * AD_THREADS) ns_max *= 2;
  int ns = g_attn_decode_split;
  if (ns == 0) {
    ns = 1;
    while (ns < ns_max && (int64_t)H * P * ns * 2 <= 256) ns *= 2;
  }
  ns = ns < ns_max ? ns : ns_max;
  const int chunk_max = (int)((vq_cdiv(Lk, ns) + 7) & ~7);
  const size_t smem = ((size_t)chunk_max + 9 * HD + 8 + 72 + AD_MAX_SPLIT * 72) * sizeof(float);
  const dim3 grid((unsigned)H, (unsigned)P, (unsigned)ns);
  if ((int64_t)H * P * ns <= 2 * vq_sm_count())
    VQ_CUDA_CALL(launch_step(attn_decode_kernel<true>, grid, dim3(AD_THREADS), smem, vq_stream(stream), (unsigned)ns, q, ldq, k_new, v_new,
                             ldn, kcache, vcache, ldc, Lmax, e1, e2, map1, map2, mask, Lk, pos, causal, out, ldo, chunk_max));
  else
    VQ_CUDA_CALL(launch_step(attn_decode_kernel<false>, grid, dim3(AD_THREADS), smem, vq_stream(stream), (unsigned)ns, q, ldq, k_new, v_new,
                             ldn, kcache, vcache, ldc, Lmax, e1, e2, map1, map2, mask, Lk, pos, causal, out, ldo, chunk_max));
  return VQ_OK;
}

VQ_API int vqcpc_sample_step(const float* logits, int ld, int V, float temperature, int top_k, float top_p, const double* u,
                             int u_stride, int64_t* tokens, int T, const int* pos, int greedy, int B, void* stream) {
  VQ_CHECK_ARG(V >= 1 && V <= 4096 && B >= 1 && temperature > 0.f, "sample_step: bad arguments V=%d B=%d", V, B);
  VQ_CHECK_ARG(greedy || u != nullptr, "sample_step: uniforms missing");
  VQ_CHECK_ARG(pos != nullptr, "sample_step: position pointer missing");
  const size_t smem = (size_t)(3 * V + (V & 1)) * sizeof(float) + (size_t)V * sizeof(double);
  static bool attr_done = false;
  if (!attr_done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(sample_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_done = true;
  }
  VQ_CUDA_CALL(launch_step(sample_step_kernel, dim3(B), dim3(256), smem, vq_stream(stream), 0, logits, ld, V, temperature, top_k, top_p, u,
                           u_stride, tokens, T, pos, greedy, B));
  return VQ_OK;
}

VQ_API int vqcpc_decode_sample_embed(const float* logits, int ld, int V, float temperature, int top_k, float top_p, const double* u,
                                     int u_stride, int64_t* tokens, int T, int* pos, int greedy, int B, const float* tables,
                                     const int* table_off, const float* chan, const float* evt, float* feat, int C,
                                     int total_upscaling, int E, int pe, int* done, int pos_mode, void* stream) {
  VQ_CHECK_ARG(V >= 1 && V <= 4096 && B >= 1 && temperature > 0.f, "decode_sample_embed: bad arguments V=%d B=%d", V, B);
  VQ_CHECK_ARG(greedy || u != nullptr, "decode_sample_embed: uniforms missing");
  VQ_CHECK_ARG(pos != nullptr && done != nullptr, "decode_sample_embed: position / counter pointer missing");
  VQ_CHECK_ARG(C >= 1 && total_upscaling % C == 0 && E >= 1 && pe >= 0, "decode_sample_embed: bad feature shape");
  const size_t smem = (size_t)(3 * V + (V & 1)) * sizeof(float) + (size_t)V * sizeof(double);
  static bool attr_done = false;
  if (!attr_done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(sample_embed_advance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_done = true;
  }
  VQ_CUDA_CALL(launch_step(sample_embed_advance_kernel, dim3(B), dim3(256), smem, vq_stream(stream), 0, logits, ld, V, temperature, top_k,
                           top_p, u, u_stride, tokens, T, pos, greedy, tables, table_off, chan, evt, feat, C, total_upscaling, E, pe,
                           done, pos_mode));
  return VQ_OK;
}

VQ_API int vqcpc_decode_ln_add(const float* y, const float* gamma, const float* beta, float eps, const float* add, int T, const int* pos,
                               float* out, int B, int d, void* stream) {
  VQ_CHECK_ARG(B >= 1 && d >= 1 && d <= 1024 && T >= 1 && pos != nullptr, "decode_ln_add: bad arguments B=%d d=%d", B, d);
  VQ_CUDA_CALL(launch_step(decode_ln_add_kernel, dim3(B), dim3(128), 0, vq_stream(stream), 0, y, gamma, beta, eps, add, T, pos, out, d));
  return VQ_OK;
}

// nn.ELU() of the aligned decoder layer's memory MLP (transformer_custom.py:428-432)
VQ_API int vqcpc_elu_fwd(const float* x, float* out, int64_t n, void* stream) {
  if (n == 0) return VQ_OK;
  const int64_t want = vq_cdiv(n, 256 * 4), cap = (int64_t)vq_sm_count() * 8;
  elu_kernel<<<(unsigned)(want > cap ? cap : want), 256, 0, vq_stream(stream)>>>(x, out, n);
  VQ_CUDA_LAUNCH_CHECK("elu_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_decode_advance(int* pos, int delta, void* stream) {
  VQ_CHECK_ARG(pos != nullptr, "decode_advance: position pointer missing");
  VQ_CUDA_CALL(launch_step(decode_advance_kernel, dim3(1), dim3(1), 0, vq_stream(stream), 0, pos, delta));
  return VQ_OK;
}
