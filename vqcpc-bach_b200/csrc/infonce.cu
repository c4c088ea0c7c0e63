// Fused InfoNCE scoring: Fks bilinear scores + log-sum-exp + accuracy bit, forward and backward.
//
// Reference: VQCPCB/vqcpc_helper.py:79-98 (FksModule: u[b,k,:] = W[:,:,k] c[b];  f = u . z) called for the positive
// (vqcpc_encoder_trainer.py:240) and for the N negatives (:245-263); nce_loss vqcpc_helper.py:5-29;
// score matrix pos > max(neg) vqcpc_encoder_trainer.py:269.
//
// Negatives are per sample (negative_samples[b] is only scored against c[b]), so scoring is a batched GEMV:
// per (b, k) one u vector against N+1 vectors of zd floats -> 0.5 FLOP/B, HBM-bound.  One warp per (b, k) with
// lane == z index (zd <= 32 per pass); W[:, :, k] is staged once per CTA (blockIdx.y = k).  The positive and every
// negative go through the SAME dot-product code so equal vectors give bit-equal scores (ties are not "correct").
//
// Layouts: c (B, cd); W (zd, cd, KR) [reference parameter layout]; y_pos (B, KR, zd); y_neg (B, N, KR, zd)
// (= the encoder output on negative_samples.view(B*N*KR, T, V), no permute needed);
// outputs u (B, KR, zd), pos (B, KR), neg (B, KR, N), lse (B, KR), correct (B, KR) int32.
#include "common.cuh"

namespace {

constexpr int MAXZ = 128;   // zd <= 128 (4 floats per lane)

template <int ZPL>   // z elements per lane: zd <= 32*ZPL
__global__ void __launch_bounds__(256)
infonce_fwd_kernel(const float* __restrict__ c, const float* __restrict__ W, const float* __restrict__ ypos,
                   const float* __restrict__ yneg, int B, int N, int KR, int zd, int cd, float* __restrict__ u_out,
                   float* __restrict__ pos_out, float* __restrict__ neg_out, float* __restrict__ lse_out,
                   float* __restrict__ nll_out, int* __restrict__ correct_out) {
  extern __shared__ float Wk[];                  // [zd][cd + 1]
  const int k = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cs = cd + 1;
  for (int i = threadIdx.x; i < zd * cd; i += blockDim.x) {
    int z = i / cd, cc = i % cd;
    Wk[z * cs + cc] = W[((size_t)z * cd + cc) * KR + k];
  }
  __syncthreads();
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int b = blockIdx.x * (blockDim.x >> 5) + warp; b < B; b += warps_total) {
    float u[ZPL];
#pragma unroll
    for (int t = 0; t < ZPL; ++t) {
      int z = lane + 32 * t;
      float a = 0.f;
      if (z < zd)
        for (int cc = 0; cc < cd; ++cc) a = fmaf(__ldg(c + (size_t)b * cd + cc), Wk[z * cs + cc], a);
      u[t] = a;
      if (z < zd) u_out[((size_t)b * KR + k) * zd + z] = a;
    }
    auto score = [&](const float* __restrict__ y) {
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < ZPL; ++t) {
        int z = lane + 32 * t;
        if (z < zd) a = fmaf(u[t], y[z], a);
      }
      return warp_sum(a);
    };
    float p = score(ypos + ((size_t)b * KR + k) * zd);
    // online log-sum-exp over [neg_0 .. neg_{N-1}, pos]; every lane carries the same running (max, sum).
    // Negatives are fetched 8 at a time (independent 128-byte loads in flight) before the dependent reduce chain.
    float mx = p, se = 1.f, nmax = -INFINITY;
    for (int n0 = 0; n0 < N; n0 += 8) {
      float yv[8][ZPL];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int t = 0; t < ZPL; ++t) {
          int z = lane + 32 * t;
          yv[j][t] = (n0 + j < N && z < zd) ? yneg[(((size_t)b * N + n0 + j) * KR + k) * zd + z] : 0.f;
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n0 + j < N) {
          float a = 0.f;
#pragma unroll
          for (int t = 0; t < ZPL; ++t) a = fmaf(u[t], yv[j][t], a);     // same fma order as score()
          float s = warp_sum(a);
          nmax = fmaxf(nmax, s);
          if (lane == 0) neg_out[((size_t)b * KR + k) * N + n0 + j] = s;
          float m2 = fmaxf(mx, s);
          se = se * expf(mx - m2) + expf(s - m2);
          mx = m2;
        }
      }
    }
    if (lane == 0) {
      pos_out[(size_t)b * KR + k] = p;
      const float l = mx + logf(se);
      lse_out[(size_t)b * KR + k] = l;
      nll_out[(size_t)b * KR + k] = l - p;         // -(pos - lse): summed by the caller (vqcpc_reduce_sum)
      correct_out[(size_t)b * KR + k] = (p > nmax) ? 1 : 0;
    }
  }
}

// backward.  scale = dL/dnce / B_total  (nce = -(1/B) sum_{b,k} (pos - lse)).
//   pi = softmax([neg, pos]); dpos = -(1 - pi_pos) * scale ; dneg_n = pi_n * scale
//   du = dpos * y_pos + sum_n dneg_n * y_neg[n] ; dy_pos = dpos * u ; dy_neg[n] = dneg_n * u
//   dc[b] += sum_k W[:, :, k]^T du[b, k] ;  dW[:, :, k] += sum_b du[b, k] (x) c[b]
template <int ZPL>
__global__ void __launch_bounds__(256)
infonce_bwd_kernel(const float* __restrict__ c, const float* __restrict__ W, const float* __restrict__ ypos,
                   const float* __restrict__ yneg, const float* __restrict__ u_in, const float* __restrict__ pos_in,
                   const float* __restrict__ neg_in, const float* __restrict__ lse_in, const float* __restrict__ g_dev,
                   float scale_host, int B, int N, int KR,
                   int zd, int cd, float* __restrict__ dc, float* __restrict__ dW, float* __restrict__ dypos,
                   float* __restrict__ dyneg) {
  extern __shared__ float sm[];
  const int cs = cd + 1;
  float* Wk = sm;                     // [zd][cs]
  float* dWk = sm + zd * cs;          // [zd][cs] CTA-private accumulator
  const int k = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = scale_host * (g_dev ? __ldg(g_dev) : 1.f);
  for (int i = threadIdx.x; i < zd * cd; i += blockDim.x) {
    int z = i / cd, cc = i % cd;
    Wk[z * cs + cc] = W[((size_t)z * cd + cc) * KR + k];
    dWk[z * cs + cc] = 0.f;
  }
  __syncthreads();
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int b = blockIdx.x * (blockDim.x >> 5) + warp; b < B; b += warps_total) {
    const size_t bk = (size_t)b * KR + k;
    const float lse = lse_in[bk];
    const float p = pos_in[bk];
    const float dpos = -(1.f - expf(p - lse)) * scale;
    float u[ZPL], du[ZPL];
#pragma unroll
    for (int t = 0; t < ZPL; ++t) {
      int z = lane + 32 * t;
      u[t] = (z < zd) ? u_in[bk * zd + z] : 0.f;
      du[t] = (z < zd) ? dpos * ypos[bk * zd + z] : 0.f;
      if (z < zd) dypos[bk * zd + z] = dpos * u[t];
    }
    for (int n0 = 0; n0 < N; n0 += 8) {
      float yv[8][ZPL], dn[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dn[j] = (n0 + j < N) ? expf(neg_in[bk * N + n0 + j] - lse) * scale : 0.f;
#pragma unroll
        for (int t = 0; t < ZPL; ++t) {
          int z = lane + 32 * t;
          yv[j][t] = (n0 + j < N && z < zd) ? yneg[(((size_t)b * N + n0 + j) * KR + k) * zd + z] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n0 + j < N) {
          size_t off = (((size_t)b * N + n0 + j) * KR + k) * zd;
#pragma unroll
          for (int t = 0; t < ZPL; ++t) {
            int z = lane + 32 * t;
            if (z < zd) {
              du[t] = fmaf(dn[j], yv[j][t], du[t]);
              dyneg[off + z] = dn[j] * u[t];
            }
          }
        }
      }
    }
    // dW_k[z][cc] += du[z] * c[b][cc]   (CTA-private smem, lane owns rows z)
    // dc[b][cc]   += sum_z Wk[z][cc] * du[z]
    for (int cc = 0; cc < cd; ++cc) {
      float cv = __ldg(c + (size_t)b * cd + cc);
      float part = 0.f;
#pragma unroll
      for (int t = 0; t < ZPL; ++t) {
        int z = lane + 32 * t;
        if (z < zd) {
          atomicAdd(&dWk[z * cs + cc], du[t] * cv);
          part = fmaf(Wk[z * cs + cc], du[t], part);
        }
      }
      part = warp_sum(part);
      if (lane == 0) atomicAdd(dc + (size_t)b * cd + cc, part);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < zd * cd; i += blockDim.x) {
    int z = i / cd, cc = i % cd;
    atomicAdd(dW + ((size_t)z * cd + cc) * KR + k, dWk[z * cs + cc]);
  }
}

}  // namespace

VQ_API int vqcpc_infonce_fwd(const float* c, const float* W, const float* y_pos, const float* y_neg, int B, int N, int KR,
                             int zd, int cd, float* u, float* pos, float* neg, float* lse, float* nll, int* correct, void* stream) {
  VQ_CHECK_ARG(zd >= 1 && zd <= MAXZ, "infonce_fwd: z_dim must be in 1..%d (got %d)", MAXZ, zd);
  VQ_CHECK_ARG(KR >= 1 && KR <= 65535 && N >= 1, "infonce_fwd: bad KR=%d / N=%d", KR, N);
  if (B == 0) return VQ_OK;
  size_t smem = (size_t)zd * (cd + 1) * sizeof(float);
  VQ_CHECK_ARG(smem <= 48 * 1024, "infonce_fwd: W[:, :, k] tile (%zu B) exceeds 48 KB", smem);
  int gx = (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(B, 8), (int64_t)vq_sm_count() * 8 / KR + 1));
  dim3 grid(gx, KR);
  cudaStream_t st = vq_stream(stream);
  int zpl = (zd + 31) / 32;
  if (zpl == 1) infonce_fwd_kernel<1><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, B, N, KR, zd, cd, u, pos, neg, lse, nll, correct);
  else if (zpl == 2) infonce_fwd_kernel<2><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, B, N, KR, zd, cd, u, pos, neg, lse, nll, correct);
  else infonce_fwd_kernel<4><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, B, N, KR, zd, cd, u, pos, neg, lse, nll, correct);
  VQ_CUDA_LAUNCH_CHECK("infonce_fwd_kernel");
  return VQ_OK;
}

// d_c (B, cd) and d_W (zd, cd, KR) are ACCUMULATED; d_ypos / d_yneg are overwritten.
// upstream gradient of sum(nll) = scale * g_dev[0] (g_dev: device scalar or NULL -> 1), so no host sync is needed.
VQ_API int vqcpc_infonce_bwd(const float* c, const float* W, const float* y_pos, const float* y_neg, const float* u,
                             const float* pos, const float* neg, const float* lse, const float* g_dev, float scale, int B, int N,
                             int KR, int zd,
                             int cd, float* d_c, float* d_W, float* d_ypos, float* d_yneg, void* stream) {
  VQ_CHECK_ARG(zd >= 1 && zd <= MAXZ, "infonce_bwd: z_dim must be in 1..%d (got %d)", MAXZ, zd);
  if (B == 0) return VQ_OK;
  size_t smem = (size_t)2 * zd * (cd + 1) * sizeof(float);
  VQ_CHECK_ARG(smem <= 48 * 1024, "infonce_bwd: W tile (%zu B) exceeds 48 KB", smem);
  int gx = (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(B, 8), (int64_t)vq_sm_count() * 4 / KR + 1));
  dim3 grid(gx, KR);
  cudaStream_t st = vq_stream(stream);
  int zpl = (zd + 31) / 32;
  if (zpl == 1) infonce_bwd_kernel<1><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, u, pos, neg, lse, g_dev, scale, B, N, KR, zd, cd, d_c, d_W, d_ypos, d_yneg);
  else if (zpl == 2) infonce_bwd_kernel<2><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, u, pos, neg, lse, g_dev, scale, B, N, KR, zd, cd, d_c, d_W, d_ypos, d_yneg);
  else infonce_bwd_kernel<4><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, u, pos, neg, lse, g_dev, scale, B, N, KR, zd, cd, d_c, d_W, d_ypos, d_yneg);
  VQ_CUDA_LAUNCH_CHECK("infonce_bwd_kernel");
  return VQ_OK;
}
