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

// ---------------------------------------------------------------------------------------------------------------------
// features of the row that enters the decoder at position i = *pos: the embedding of token i-1 with the channel / event
// positional embeddings of position i-1 (decoder.py:442-457, then shifted right by one :476-482).  Row 0 is the start symbol
// and is substituted after linear_target (linear_small's `row0`), so zeros are written here.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void decode_features_kernel(const int64_t* __restrict__ tokens, const int* __restrict__ pos,
                                       const float* __restrict__ tables, const int* __restrict__ table_off,
                                       const float* __restrict__ chan, const float* __restrict__ evt, float* __restrict__ feat,
                                       int T, int C, int up, int E, int pe) {
  const int b = blockIdx.x, i = *pos, F = E + 2 * pe;
  float* f = feat + (size_t)b * F;
  if (i <= 0 || i >= T + 1) {
    for (int c = threadIdx.x; c < F; c += blockDim.x) f[c] = 0.f;
    return;
  }
  const int t = i - 1, ch = t % C, ev = (t % up) / C;
  const int64_t tok = tokens[(size_t)b * T + t];
  const float* row = tables + ((size_t)table_off[ch] + tok) * E;
  for (int c = threadIdx.x; c < F; c += blockDim.x)
    f[c] = c < E ? row[c] : (c < E + pe ? chan[ch * pe + (c - E)] : evt[ev * pe + (c - E - pe)]);
}

// ---------------------------------------------------------------------------------------------------------------------
// out[b, n] = epilogue( sum_k x[b, k] W[n, k] + bias[n] ),  b < M <= 16.  One warp per output column, x staged in shared
// memory, W rows streamed with 16-byte loads.  Epilogue order = the layer code's: bias, scale of the first `scale_cols` columns
// (the q third of the in-projection, multihead_attention_custom.py:247), ReLU, residual.  row0 != NULL and *pos == 0: the whole
// row is replaced by row0 (the start-of-sequence vector, decoder.py:476-482).
// ---------------------------------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(128)
linear_small_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
                    const float* __restrict__ residual, int ldr, float* __restrict__ out, int ldo, int M, int N, int K, int relu,
                    float scale, int scale_cols, const float* __restrict__ row0, const int* __restrict__ pos,
                    const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta, float ln_eps,
                    float* __restrict__ ln_out) {
  extern __shared__ __align__(16) float xs[];      // [MT][K]
  for (int t = threadIdx.x; t < MT * (K / 4); t += blockDim.x) {
    const int b = t / (K / 4), c = t % (K / 4);
    reinterpret_cast<float4*>(xs)[t] = b < M ? *reinterpret_cast<const float4*>(x + (size_t)b * ldx + c * 4) : make_float4(0, 0, 0, 0);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (ln_gamma) {
    // nn.LayerNorm over the K input columns applied while staging (transformer_custom.py:283-289 post-LN: the producer wrote
    // the pre-norm sum); biased variance, eps inside the square root, two passes like layernorm_fwd_kernel.  CTA 0 also
    // stores the normalised rows: they are the residual input of a later linear.
    for (int b = warp; b < M; b += 4) {
      float* row = xs + b * K;
      float sum = 0.f;
      for (int c = lane; c < K; c += 32) sum += row[c];
      const float mu = warp_sum(sum) / K;
      float q = 0.f;
      for (int c = lane; c < K; c += 32) {
        const float a = row[c] - mu;
        q += a * a;
      }
      const float rs = 1.0f / sqrtf(warp_sum(q) / K + ln_eps);
      for (int c = lane; c < K; c += 32) {
        const float o = (row[c] - mu) * rs * ln_gamma[c] + ln_beta[c];
        row[c] = o;
        if (ln_out && blockIdx.x == 0) ln_out[(size_t)b * K + c] = o;
      }
    }
    __syncthreads();
  }
  const int n = blockIdx.x * 4 + warp;
  if (n >= N) return;                              // (after the last block-wide barrier)
  const bool use_row0 = row0 != nullptr && *pos == 0;
  float acc[MT];
#pragma unroll
  for (int b = 0; b < MT; ++b) acc[b] = 0.f;
  const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)n * K);
#pragma unroll 4
  for (int c = lane; c < K / 4; c += 32) {
    const float4 w = __ldg(w4 + c);
#pragma unroll
    for (int b = 0; b < MT; ++b) {
      const float4 v = reinterpret_cast<const float4*>(xs + b * K)[c];
      acc[b] += w.x * v.x + w.y * v.y + w.z * v.z + w.w * v.w;
    }
  }
#pragma unroll
  for (int b = 0; b < MT; ++b) acc[b] = warp_sum(acc[b]);
  const float bn = bias ? bias[n] : 0.f;
#pragma unroll
  for (int b = 0; b < MT; ++b) {
    if (lane == b && b < M) {
      float v = acc[b] + bn;
      if (n < scale_cols) v *= scale;
      if (relu) v = fmaxf(v, 0.f);
      if (residual) v += residual[(size_t)b * ldr + n];
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

__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__global__ void __launch_bounds__(AD_THREADS)
attn_decode_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k_new, const float* __restrict__ v_new, int ldn,
                   float* kcache, float* vcache, int ldc, int Lmax, const float* __restrict__ e1, const float* __restrict__ e2,
                   const int* __restrict__ map1, const int* __restrict__ map2, const float* __restrict__ mask, int Lk,
                   const int* __restrict__ pos, int causal, float* __restrict__ out, int ldo, int chunk_max) {
  extern __shared__ __align__(16) float sm[];      // [64] q, [2][64] partial outputs, [8] reductions, [66] result, [chunk] scores
  float* qs = sm;
  float* part = qs + HD;
  float* red = part + 2 * HD;
  float* result = red + 8;                         // [0..63] sum_j e^(s_j - m) v_j, [64] m, [65] z   (read by CTA 0 of the cluster)
  float* sc = result + 72;
  uint32_t crank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const int h = blockIdx.x, tid = threadIdx.x, i = *pos;
  const int64_t p = blockIdx.y;
  const int n = causal ? min(i + 1, Lk) : Lk;
  int chunk = (((n + (int)csize - 1) / (int)csize) + 7) & ~7;
  chunk = min(chunk, chunk_max);
  const int j_lo = min(n, (int)crank * chunk), j_hi = min(n, j_lo + chunk), cnt = j_hi - j_lo;
  float* kc = kcache + (size_t)p * Lmax * ldc + h * HD;
  float* vc = vcache + (size_t)p * Lmax * ldc + h * HD;
  const float* kn = k_new ? k_new + (size_t)p * ldn + h * HD : nullptr;
  const float* vn = v_new ? v_new + (size_t)p * ldn + h * HD : nullptr;
  if (tid < HD) {
    qs[tid] = q[(size_t)p * ldq + h * HD + tid];
    if (kn && crank == 0) {                        // the appended row is read from k_new / v_new below, never from the cache
      kc[(size_t)i * ldc + tid] = kn[tid];
      vc[(size_t)i * ldc + tid] = vn[tid];
    }
  }
  __syncthreads();
  // scores: 16 lanes per key, one float4 each, 8 keys per pass, two passes in flight
  const int sub = tid & 15, grp = tid >> 4;
  const float4 q4 = reinterpret_cast<const float4*>(qs)[sub];
  const float* e1h = e1 ? e1 + (size_t)h * Lk * HD : nullptr;
  const float* e2h = e2 ? e2 + (size_t)h * Lk * HD : nullptr;
  const int* m1row = map1 ? map1 + (size_t)i * Lk : nullptr;
  const int* m2row = map2 ? map2 + (size_t)i * Lk : nullptr;
  const float* mrow = mask ? mask + (size_t)i * Lk : nullptr;
  for (int j0 = j_lo; j0 < j_hi; j0 += 16) {
    float s[2], add[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int j = j0 + w * 8 + grp;
      s[w] = 0.f;
      add[w] = 0.f;
      if (j < j_hi) {
        const float* krow = (kn && j == i) ? kn : kc + (size_t)j * ldc;
        s[w] = dot4(q4, *reinterpret_cast<const float4*>(krow + sub * 4));
        if (m1row) {
          const int m = m1row[j];
          if (m >= 0) s[w] += dot4(q4, __ldg(reinterpret_cast<const float4*>(e1h + (size_t)(m % Lk) * HD) + sub));
          else if (m == -2) add[w] += -100.f;
        }
        if (m2row) {
          const int m = m2row[j];
          if (m >= 0) s[w] += dot4(q4, __ldg(reinterpret_cast<const float4*>(e2h + (size_t)(m % Lk) * HD) + sub));
          else if (m == -2) add[w] += -100.f;
        }
        if (mrow) add[w] += mrow[j];
      }
    }
#pragma unroll
    for (int w = 0; w < 2; ++w) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s[w] += __shfl_xor_sync(0xffffffffu, s[w], o);
      const int j = j0 + w * 8 + grp;
      if (sub == 0 && j < j_hi) sc[j - j_lo] = s[w] + add[w];
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
  // o[d] = sum_j e_j v[j, d]: two key phases x 64 dims, two independent accumulators each
  const int d = tid & 63, ph = tid >> 6;
  float o0 = 0.f, o1 = 0.f;
  int j = ph;
  for (; j + 2 < cnt; j += 4) {
    const int ja = j_lo + j, jb = ja + 2;
    const float* va = (vn && ja == i) ? vn : vc + (size_t)ja * ldc;
    const float* vb = (vn && jb == i) ? vn : vc + (size_t)jb * ldc;
    o0 += sc[j] * va[d];
    o1 += sc[j + 2] * vb[d];
  }
  for (; j < cnt; j += 2) {
    const int ja = j_lo + j;
    const float* va = (vn && ja == i) ? vn : vc + (size_t)ja * ldc;
    o0 += sc[j] * va[d];
  }
  part[ph * HD + d] = o0 + o1;
  __syncthreads();
  if (tid < HD) result[tid] = part[tid] + part[HD + tid];
  if (tid == 0) {
    result[64] = mx;
    result[65] = red[4] + red[5] + red[6] + red[7];
  }
  if (csize == 1) {
    __syncthreads();
    if (tid < HD) out[(size_t)p * ldo + h * HD + tid] = result[tid] / result[65];
    return;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (crank == 0 && tid < HD) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(result);
    float mr[AD_MAX_SPLIT], M = -INFINITY;
    for (uint32_t r = 0; r < csize; ++r) {
      uint32_t ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base + 4u * 64u), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(mr[r]) : "r"(ra));
      M = fmaxf(M, mr[r]);
    }
    float Z = 0.f, O = 0.f;
    for (uint32_t r = 0; r < csize; ++r) {
      if (mr[r] == -INFINITY) continue;
      uint32_t ra, rz;
      float zr, orr;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rz) : "r"(base + 4u * 65u), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(zr) : "r"(rz));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(base + 4u * (uint32_t)tid), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(orr) : "r"(ra));
      const float w = __expf(mr[r] - M);
      Z += zr * w;
      O += orr * w;
    }
    out[(size_t)p * ldo + h * HD + tid] = O / Z;
  }
  // peers must not exit (and release their shared memory) before CTA 0 has read it
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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
  for (int c = tid; c < V; c += 256) x[c] = lg[c] / temperature;
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
  if (tid == 0) {
    int result;
    if (greedy) {
      result = order[0];
    } else {
      const int k = min(top_k, V);
      if (k > 0) {                                           // utils.py:112-115: logits < k-th largest removed (ties stay)
        const float kth = srt[k - 1];
        for (int r = k; r < V; ++r)
          if (srt[r] < kth) srt[r] = -INFINITY;
      }
      if (top_p > 0.f) {                                     // utils.py:117-127
        const float mx = srt[0];
        float z = 0.f;
        for (int r = 0; r < V; ++r) z += expf(srt[r] - mx);
        float cum = 0.f;
        bool prev_above = false;
        for (int r = 0; r < V; ++r) {
          cum += expf(srt[r] - mx) / z;
          const bool above = cum > top_p;
          if (prev_above) srt[r] = -INFINITY;                // shifted right by one: the first token above the threshold stays
          prev_above = above;
        }
      }
      // softmax of the filtered logits in ORIGINAL index order, double cdf like numpy
      const float mx = srt[0];
      for (int r = 0; r < V; ++r) x[order[r]] = srt[r];
      float z = 0.f;
      for (int c = 0; c < V; ++c) z += expf(x[c] - mx);
      double run = 0.0;
      for (int c = 0; c < V; ++c) {
        run += (double)(expf(x[c] - mx) / z);
        cdf[c] = run;
      }
      const double uu = uniform * run;                       // cdf /= cdf[-1]  <=>  compare with u * cdf[-1]
      result = V - 1;
      for (int c = 0; c < V; ++c)
        if (cdf[c] > uu) { result = c; break; }
    }
    s_result = result;
  }
  __syncthreads();
  return s_result;
}

__global__ void __launch_bounds__(256)
sample_step_kernel(const float* __restrict__ logits, int ld, int V, float temperature, int top_k, float top_p,
                   const double* __restrict__ u, int u_stride, int64_t* __restrict__ tokens, int T, const int* __restrict__ pos,
                   int greedy, int B) {
  extern __shared__ __align__(16) float smf[];
  const int b = blockIdx.x, i = *pos;
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
                            const float* __restrict__ evt, float* __restrict__ feat, int C, int up, int E, int pe, int* done) {
  extern __shared__ __align__(16) float smf[];
  const int b = blockIdx.x, i = *reinterpret_cast<volatile int*>(pos);
  if (i >= 0 && i < T) {
    const int tok = sample_row(logits + (size_t)b * ld, V, temperature, top_k, top_p, greedy ? 0.0 : u[(size_t)i * u_stride + b], greedy, smf);
    if (threadIdx.x == 0) tokens[(size_t)b * T + i] = tok;
    const int ch = i % C, ev = (i % up) / C, F = E + 2 * pe;
    const float* row = tables + ((size_t)table_off[ch] + tok) * E;
    for (int c = threadIdx.x; c < F; c += blockDim.x)
      feat[(size_t)b * F + c] = c < E ? row[c] : (c < E + pe ? chan[ch * pe + (c - E)] : evt[ev * pe + (c - E - pe)]);
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

__global__ void decode_advance_kernel(int* pos, int delta) { *pos += delta; }

}  // namespace

VQ_API int vqcpc_decode_features(const int64_t* tokens, const int* pos, const float* tables, const int* table_off, const float* chan,
                                 const float* evt, float* feat, int B, int T, int C, int total_upscaling, int E, int pe,
                                 void* stream) {
  VQ_CHECK_ARG(B >= 1 && T >= 1 && C >= 1 && total_upscaling % C == 0 && E >= 1 && pe >= 0, "decode_features: bad shape");
  decode_features_kernel<<<B, 64, 0, vq_stream(stream)>>>(tokens, pos, tables, table_off, chan, evt, feat, T, C, total_upscaling, E, pe);
  VQ_CUDA_LAUNCH_CHECK("decode_features_kernel");
  return VQ_OK;
}

static int linear_small_launch(const float* x, int ldx, const float* W, const float* bias, const float* residual, int ldr, float* out,
                               int ldo, int M, int N, int K, int relu, float scale, int scale_cols, const float* row0, const int* pos,
                               const float* ln_gamma, const float* ln_beta, float ln_eps, float* ln_out, void* stream) {
  VQ_CHECK_ARG(M >= 1 && M <= 16, "linear_small: 1 <= M <= 16 rows per call (got %d)", M);
  VQ_CHECK_ARG(N >= 1 && K >= 4 && (K % 4) == 0 && (ldx % 4) == 0 && K <= 8192, "linear_small: bad shape N=%d K=%d ldx=%d", N, K, ldx);
  VQ_CHECK_ARG(row0 == nullptr || pos != nullptr, "linear_small: row0 needs the position pointer");
  VQ_CHECK_ARG((ln_gamma == nullptr) == (ln_beta == nullptr) && (ln_out == nullptr || ln_gamma != nullptr), "linear_small: LayerNorm arguments");
  const int MT = M <= 1 ? 1 : M <= 2 ? 2 : M <= 4 ? 4 : M <= 8 ? 8 : 16;
  const size_t smem = (size_t)MT * K * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(linear_small_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VQ_CUDA_CALL(cudaFuncSetAttribute(linear_small_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VQ_CUDA_CALL(cudaFuncSetAttribute(linear_small_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VQ_CUDA_CALL(cudaFuncSetAttribute(linear_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VQ_CUDA_CALL(cudaFuncSetAttribute(linear_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  VQ_CHECK_ARG(smem <= 200 * 1024, "linear_small: M x K does not fit shared memory");
  const unsigned grid = (unsigned)vq_cdiv(N, 4);
  cudaStream_t st = vq_stream(stream);
#define VQ_LS(MTV)                                                                                                                 \
  linear_small_kernel<MTV><<<grid, 128, smem, st>>>(x, ldx, W, bias, residual, ldr, out, ldo, M, N, K, relu, scale, scale_cols, row0, \
                                                    pos, ln_gamma, ln_beta, ln_eps, ln_out)
  switch (MT) {
    case 1: VQ_LS(1); break;
    case 2: VQ_LS(2); break;
    case 4: VQ_LS(4); break;
    case 8: VQ_LS(8); break;
    default: VQ_LS(16); break;
  }
#undef VQ_LS
  VQ_CUDA_LAUNCH_CHECK("linear_small_kernel");
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
  // split the keys over a cluster so that about 2 CTAs per SM are in flight, at least 64 keys per CTA
  int ns = g_attn_decode_split;
  if (ns == 0) {
    ns = 1;
    while (ns < AD_MAX_SPLIT && (int64_t)H * P * ns < 2 * vq_sm_count() && Lk / (2 * ns) >= 32) ns *= 2;
  }
  const int chunk_max = (int)((vq_cdiv(Lk, ns) + 7) & ~7);
  const size_t smem = ((size_t)chunk_max + 3 * HD + 8 + 72) * sizeof(float);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)H, (unsigned)P, (unsigned)ns);
  cfg.blockDim = dim3(AD_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = vq_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)ns;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VQ_CUDA_CALL(cudaLaunchKernelEx(&cfg, attn_decode_kernel, q, ldq, k_new, v_new, ldn, kcache, vcache, ldc, Lmax, e1, e2, map1, map2,
                                  mask, Lk, pos, causal, out, ldo, chunk_max));
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
  sample_step_kernel<<<B, 256, smem, vq_stream(stream)>>>(logits, ld, V, temperature, top_k, top_p, u, u_stride, tokens, T, pos,
                                                          greedy, B);
  VQ_CUDA_LAUNCH_CHECK("sample_step_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_decode_sample_embed(const float* logits, int ld, int V, float temperature, int top_k, float top_p, const double* u,
                                     int u_stride, int64_t* tokens, int T, int* pos, int greedy, int B, const float* tables,
                                     const int* table_off, const float* chan, const float* evt, float* feat, int C,
                                     int total_upscaling, int E, int pe, int* done, void* stream) {
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
  sample_embed_advance_kernel<<<B, 256, smem, vq_stream(stream)>>>(logits, ld, V, temperature, top_k, top_p, u, u_stride, tokens, T, pos,
                                                                   greedy, tables, table_off, chan, evt, feat, C, total_upscaling, E,
                                                                   pe, done);
  VQ_CUDA_LAUNCH_CHECK("sample_embed_advance_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_decode_advance(int* pos, int delta, void* stream) {
  VQ_CHECK_ARG(pos != nullptr, "decode_advance: position pointer missing");
  decode_advance_kernel<<<1, 1, 0, vq_stream(stream)>>>(pos, delta);
  VQ_CUDA_LAUNCH_CHECK("decode_advance_kernel");
  return VQ_OK;
}
