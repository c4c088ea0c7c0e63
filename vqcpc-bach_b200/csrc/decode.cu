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
//     decode_features -> linear_small(linear_target, sos at i == 0) -> per layer [ linear_small(QKV) -> attn_decode(self, appends
//     k / v to the cache) -> linear_small(out-proj + residual) -> layernorm -> linear_small(q) -> attn_decode(cross) ->
//     linear_small(out-proj + residual) -> layernorm -> linear_small(FFN1, ReLU) -> linear_small(FFN2 + residual) -> layernorm ]
//     -> linear_small(pre-softmax of the channel) -> sample_step (temperature, top-k, top-p, inverse-CDF draw like
//     numpy.random.choice) -> decode_advance.
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
                    float scale, int scale_cols, const float* __restrict__ row0, const int* __restrict__ pos) {
  extern __shared__ __align__(16) float xs[];      // [MT][K]
  for (int t = threadIdx.x; t < MT * (K / 4); t += blockDim.x) {
    const int b = t / (K / 4), c = t % (K / 4);
    reinterpret_cast<float4*>(xs)[t] = b < M ? *reinterpret_cast<const float4*>(x + (size_t)b * ldx + c * 4) : make_float4(0, 0, 0, 0);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + warp;
  if (n >= N) return;
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
// value are first appended to the cache at row i (decoder self-attention).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_decode_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k_new, const float* __restrict__ v_new, int ldn,
                   float* kcache, float* vcache, int ldc, int Lmax, const float* __restrict__ e1, const float* __restrict__ e2,
                   const int* __restrict__ map1, const int* __restrict__ map2, const float* __restrict__ mask, int Lk,
                   const int* __restrict__ pos, int causal, float* __restrict__ out, int ldo) {
  extern __shared__ __align__(16) float sm[];      // [64] q, [2][64] partial outputs, [8] reductions, [Lk] scores
  float* qs = sm;
  float* part = qs + HD;
  float* red = part + 2 * HD;
  float* sc = red + 8;
  const int h = blockIdx.x, tid = threadIdx.x, i = *pos;
  const int64_t p = blockIdx.y;
  const int n = causal ? min(i + 1, Lk) : Lk;
  float* kc = kcache + (size_t)p * Lmax * ldc + h * HD;
  float* vc = vcache + (size_t)p * Lmax * ldc + h * HD;
  if (tid < HD) {
    qs[tid] = q[(size_t)p * ldq + h * HD + tid];
    if (k_new) {
      kc[(size_t)i * ldc + tid] = k_new[(size_t)p * ldn + h * HD + tid];
      vc[(size_t)i * ldc + tid] = v_new[(size_t)p * ldn + h * HD + tid];
    }
  }
  __syncthreads();
  // scores: 16 lanes per key, one float4 each
  const int sub = tid & 15, grp = tid >> 4;         // 8 keys per pass
  const float4 q4 = reinterpret_cast<const float4*>(qs)[sub];
  const float* e1h = e1 ? e1 + (size_t)h * Lk * HD : nullptr;
  const float* e2h = e2 ? e2 + (size_t)h * Lk * HD : nullptr;
  const int* m1row = map1 ? map1 + (size_t)i * Lk : nullptr;
  const int* m2row = map2 ? map2 + (size_t)i * Lk : nullptr;
  const float* mrow = mask ? mask + (size_t)i * Lk : nullptr;
  for (int j0 = 0; j0 < n; j0 += 8) {
    const int j = j0 + grp;
    float s = 0.f, add = 0.f;
    if (j < n) {
      const float4 kk = *reinterpret_cast<const float4*>(kc + (size_t)j * ldc + sub * 4);
      s = q4.x * kk.x + q4.y * kk.y + q4.z * kk.z + q4.w * kk.w;
      if (m1row) {
        const int m = m1row[j];
        if (m >= 0) {
          const float4 ee = __ldg(reinterpret_cast<const float4*>(e1h + (size_t)(m % Lk) * HD) + sub);
          s += q4.x * ee.x + q4.y * ee.y + q4.z * ee.z + q4.w * ee.w;
        } else if (m == -2) add += -100.f;
      }
      if (m2row) {
        const int m = m2row[j];
        if (m >= 0) {
          const float4 ee = __ldg(reinterpret_cast<const float4*>(e2h + (size_t)(m % Lk) * HD) + sub);
          s += q4.x * ee.x + q4.y * ee.y + q4.z * ee.z + q4.w * ee.w;
        } else if (m == -2) add += -100.f;
      }
      if (mrow) add += mrow[j];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (sub == 0 && j < n) sc[j] = s + add;
  }
  __syncthreads();
  // softmax over n scores
  const int warp = tid >> 5, lane = tid & 31;
  float mx = -INFINITY;
  for (int j = tid; j < n; j += 128) mx = fmaxf(mx, sc[j]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float z = 0.f;
  for (int j = tid; j < n; j += 128) {
    const float ev = __expf(sc[j] - mx);
    sc[j] = ev;
    z += ev;
  }
  z = warp_sum(z);
  if (lane == 0) red[4 + warp] = z;
  __syncthreads();
  const float inv = 1.f / (red[4] + red[5] + red[6] + red[7]);
  // o[d] = sum_j a[j] v[j, d]: two key phases x 64 dims
  const int d = tid & 63, ph = tid >> 6;
  float o0 = 0.f, o1 = 0.f;
  int j = ph;
  for (; j + 2 < n; j += 4) {
    o0 += sc[j] * vc[(size_t)j * ldc + d];
    o1 += sc[j + 2] * vc[(size_t)(j + 2) * ldc + d];
  }
  for (; j < n; j += 2) o0 += sc[j] * vc[(size_t)j * ldc + d];
  part[ph * HD + d] = o0 + o1;
  __syncthreads();
  if (tid < HD) out[(size_t)p * ldo + h * HD + tid] = (part[tid] + part[HD + tid]) * inv;
}

// ---------------------------------------------------------------------------------------------------------------------
// one token per sequence from a row of logits: logits / temperature, top-k, top-p (VQCPCB/utils.py:100-128), softmax, then the
// draw of numpy.random.choice(V, p=p) (decoder.py:636-644): cdf = cumsum(double(p)); cdf /= cdf[-1]; index of the first
// cdf > u, with u the uniform the caller supplies for this (position, sequence) - the same stream numpy would have consumed.
// greedy: argmax of the logits (lowest index on ties).  The token is written to tokens[b, *pos].
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sample_step_kernel(const float* __restrict__ logits, int ld, int V, float temperature, int top_k, float top_p,
                   const double* __restrict__ u, int u_stride, int64_t* __restrict__ tokens, int T, const int* __restrict__ pos,
                   int greedy, int B) {
  extern __shared__ __align__(16) float smf[];     // [V] logits, [V] sorted, then int [V] order (sorted position -> index)
  float* x = smf;
  float* srt = smf + V;
  int* order = reinterpret_cast<int*>(smf + 2 * V);
  double* cdf = reinterpret_cast<double*>(smf + 3 * V + (V & 1));
  const int b = blockIdx.x, tid = threadIdx.x, i = *pos;
  if (i < 0 || i >= T) return;
  const float* lg = logits + (size_t)b * ld;
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
      const double uu = u[(size_t)i * u_stride + b] * run;   // cdf /= cdf[-1]  <=>  compare with u * cdf[-1]
      result = V - 1;
      for (int c = 0; c < V; ++c)
        if (cdf[c] > uu) { result = c; break; }
    }
    tokens[(size_t)b * T + i] = result;
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

VQ_API int vqcpc_linear_small(const float* x, int ldx, const float* W, const float* bias, const float* residual, int ldr, float* out,
                              int ldo, int M, int N, int K, int relu, float scale, int scale_cols, const float* row0, const int* pos,
                              void* stream) {
  VQ_CHECK_ARG(M >= 1 && M <= 16, "linear_small: 1 <= M <= 16 rows per call (got %d)", M);
  VQ_CHECK_ARG(N >= 1 && K >= 4 && (K % 4) == 0 && (ldx % 4) == 0 && K <= 8192, "linear_small: bad shape N=%d K=%d ldx=%d", N, K, ldx);
  VQ_CHECK_ARG(row0 == nullptr || pos != nullptr, "linear_small: row0 needs the position pointer");
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
#define VQ_LS(MTV) \
  linear_small_kernel<MTV><<<grid, 128, smem, st>>>(x, ldx, W, bias, residual, ldr, out, ldo, M, N, K, relu, scale, scale_cols, row0, pos)
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

VQ_API int vqcpc_attn_decode(const float* q, int ldq, const float* k_new, const float* v_new, int ldn, float* kcache, float* vcache,
                             int ldc, int Lmax, const float* e1, const float* e2, const int* map1, const int* map2, const float* mask,
                             int Lk, const int* pos, int causal, float* out, int ldo, int64_t P, int H, int head_dim, void* stream) {
  VQ_CHECK_ARG(head_dim == HD, "attn_decode: head_dim must be 64 (got %d)", head_dim);
  VQ_CHECK_ARG(Lk >= 1 && Lk <= 8192 && Lmax >= Lk && H >= 1 && P >= 1 && P <= 65535, "attn_decode: bad shape Lk=%d Lmax=%d H=%d", Lk, Lmax, H);
  VQ_CHECK_ARG((map1 == nullptr) == (e1 == nullptr) && (map2 == nullptr) == (e2 == nullptr), "attn_decode: a map needs its table");
  VQ_CHECK_ARG((k_new == nullptr) == (v_new == nullptr) && (ldc % 4) == 0, "attn_decode: bad cache arguments");
  VQ_CHECK_ARG(pos != nullptr, "attn_decode: position pointer missing");
  const size_t smem = ((size_t)Lk + 3 * HD + 8) * sizeof(float);
  dim3 grid((unsigned)H, (unsigned)P);
  attn_decode_kernel<<<grid, 128, smem, vq_stream(stream)>>>(q, ldq, k_new, v_new, ldn, kcache, vcache, ldc, Lmax, e1, e2, map1, map2,
                                                             mask, Lk, pos, causal, out, ldo);
  VQ_CUDA_LAUNCH_CHECK("attn_decode_kernel");
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

VQ_API int vqcpc_decode_advance(int* pos, int delta, void* stream) {
  VQ_CHECK_ARG(pos != nullptr, "decode_advance: position pointer missing");
  decode_advance_kernel<<<1, 1, 0, vq_stream(stream)>>>(pos, delta);
  VQ_CUDA_LAUNCH_CHECK("decode_advance_kernel");
  return VQ_OK;
}
