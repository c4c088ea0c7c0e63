// General relative attention forward for the inference path of the decoder (decoders/decoder.py:431-543
// forward, :552-722 generate): encoder self-attention over the code sequence, causal decoder self-attention, and the
// target -> source cross-attention whose relative bias comes from SubsampledRelativeAttention with seq_len_tgt != seq_len_src
// (transformer/subsampled_relative_attention.py:30-122: a (T, S) product q.e^T is REINTERPRETED as (S, T) by `view`, padded
// and re-viewed, so entry (i, j) of the bias may be the product of ANOTHER query row t with table row s).
//
// All of that is covered by ONE kernel through two host-built integer maps (Lq x Lk each, one per table; VQCPCB/transformer/
// relative_maps.py replays the reference's pad / view / slice / mask sequence on flat indices):
//     m >= 0 : bias term = q~_t . e[h, s]  with  t = m / Lk, s = m % Lk      m == -1 : 0 (masked)      m == -2 : -100 (pad)
//     s[i, j] = q~_i . k_j + term(map1, e1) + term(map2, e2) + mask[i, j];   a = softmax_j(s);   o_i = sum_j a[i, j] v_j
// q (P*Lq rows) and k / v (P*Lk rows) may live in different matrices (cross-attention) with arbitrary column offsets.
// Forward only, fp32 CUDA cores: generation runs at batch 1 - 8; the layer's GEMMs use the tcgen05 path.
#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int R = 8;
constexpr int THREADS = 256;

__device__ __forceinline__ float dot64(const float* __restrict__ a, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll 4
  for (int c = 0; c < HD / 4; ++c) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + c), y = __ldg(reinterpret_cast<const float4*>(b) + c);
    s += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  return s;
}

__global__ void __launch_bounds__(THREADS)
attn_map_fwd_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, const float* __restrict__ v, int ldkv,
                    const float* __restrict__ e1, const float* __restrict__ e2, const int* __restrict__ map1,
                    const int* __restrict__ map2, const float* __restrict__ mask, float* __restrict__ out,
                    float* __restrict__ probs, int Lq, int Lk, int H) {
  extern __shared__ __align__(16) float sc[];        // [R][Lk]
  const int i0 = blockIdx.x * R, h = blockIdx.y;
  const int64_t p = blockIdx.z;
  const float* qb = q + p * Lq * (int64_t)ldq + h * HD;
  const float* kb = k + p * Lk * (int64_t)ldkv + h * HD;
  const float* vb = v + p * Lk * (int64_t)ldkv + h * HD;
  const float* e1h = e1 ? e1 + (size_t)h * Lk * HD : nullptr;
  const float* e2h = e2 ? e2 + (size_t)h * Lk * HD : nullptr;
  const int rows = min(R, Lq - i0);
  for (int t = threadIdx.x; t < rows * Lk; t += THREADS) {
    const int ii = t / Lk, j = t % Lk, i = i0 + ii;
    float s = dot64(qb + (int64_t)i * ldq, kb + (int64_t)j * ldkv);
    if (map1) {
      const int m = map1[(int64_t)i * Lk + j];
      s += m >= 0 ? dot64(qb + (int64_t)(m / Lk) * ldq, e1h + (size_t)(m % Lk) * HD) : (m == -2 ? -100.f : 0.f);
    }
    if (map2) {
      const int m = map2[(int64_t)i * Lk + j];
      s += m >= 0 ? dot64(qb + (int64_t)(m / Lk) * ldq, e2h + (size_t)(m % Lk) * HD) : (m == -2 ? -100.f : 0.f);
    }
    if (mask) s += mask[(int64_t)i * Lk + j];
    sc[ii * Lk + j] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < rows) {
    float* s = sc + warp * Lk;
    float m = -INFINITY;
    for (int j = lane; j < Lk; j += 32) m = fmaxf(m, s[j]);
    m = warp_max(m);
    float z = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float ev = __expf(s[j] - m);
      s[j] = ev;
      z += ev;
    }
    z = warp_sum(z);
    const float inv = 1.f / z;
    float* pg = probs ? probs + ((p * H + h) * (int64_t)Lq + (i0 + warp)) * Lk : nullptr;
    for (int j = lane; j < Lk; j += 32) {
      const float a = s[j] * inv;
      s[j] = a;
      if (pg) pg[j] = a;
    }
  }
  __syncthreads();
  const int ii = threadIdx.x >> 5, d = (threadIdx.x & 31) * 2;
  if (ii < rows) {
    const float* s = sc + ii * Lk;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll 4
    for (int j = 0; j < Lk; ++j) {
      const float2 v2 = __ldg(reinterpret_cast<const float2*>(vb + (int64_t)j * ldkv + d));
      o0 += s[j] * v2.x;
      o1 += s[j] * v2.y;
    }
    *reinterpret_cast<float2*>(out + (size_t)(p * Lq + i0 + ii) * (H * HD) + h * HD + d) = make_float2(o0, o1);
  }
}

}  // namespace

// q: (P*Lq, ldq) with head h at columns [h*64, +64) of the given pointer (pre-scaled by hd^-0.5); k, v likewise over P*Lk rows
// (ldkv); e1 / e2 (H*Lk, 64) or NULL; map1 / map2 (Lq, Lk) int32 or NULL; mask (Lq, Lk) additive or NULL; out (P*Lq, H*64);
// probs (P, H, Lq, Lk) or NULL.
VQ_API int vqcpc_attn_map_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* e1, const float* e2,
                              const int* map1, const int* map2, const float* mask, float* out, float* probs, int64_t P, int Lq,
                              int Lk, int H, int head_dim, void* stream) {
  VQ_CHECK_ARG(head_dim == HD, "attn_map_fwd: head_dim must be 64 (got %d)", head_dim);
  VQ_CHECK_ARG(Lq >= 1 && Lk >= 1 && Lk <= 4096 && H >= 1 && H <= 64, "attn_map_fwd: bad shape Lq=%d Lk=%d H=%d", Lq, Lk, H);
  VQ_CHECK_ARG((map1 == nullptr) == (e1 == nullptr) && (map2 == nullptr) == (e2 == nullptr), "attn_map_fwd: a map needs its table");
  VQ_CHECK_ARG((ldq % 4) == 0 && (ldkv % 4) == 0, "attn_map_fwd: leading dimensions must be multiples of 4 floats");
  if (P == 0) return VQ_OK;
  VQ_CHECK_ARG(P <= 65535, "attn_map_fwd: at most 65535 sequences per call");
  const size_t smem = (size_t)R * Lk * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    VQ_CUDA_CALL(cudaFuncSetAttribute(attn_map_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, R * 4096 * 4));
    attr_done = true;
  }
  dim3 grid((unsigned)vq_cdiv(Lq, R), (unsigned)H, (unsigned)P);
  attn_map_fwd_kernel<<<grid, THREADS, smem, vq_stream(stream)>>>(q, ldq, k, v, ldkv, e1, e2, map1, map2, mask, out, probs, Lq, Lk, H);
  VQ_CUDA_LAUNCH_CHECK("attn_map_fwd_kernel");
  return VQ_OK;
}
