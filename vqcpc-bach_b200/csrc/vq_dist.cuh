// Bit-exact squared-distance evaluation shared by every nearest-code search kernel (quantizer.cu, vq_tc.cu).
//
// Reference: dist[n, k] = sum_d (x[n, d] - E[k, d])^2 by broadcast + torch.sum(dim=-1)  (quantizer/vector_quantizer.py:105-112).
// fp32, no FMA contraction, and the SAME summation order torch's CPU sum uses for a contiguous inner dimension
// (8 interleaved partial sums combined in order; 32 when D % 32 == 0), so distances - and therefore the argmin, ties
// included - are bit-identical to the reference for D <= 8 and D in {16, 24, 32, 64, 128} (checked against torch 2.11 CPU).
#pragma once
#include "common.cuh"

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

// loss of one vector against its code, exactly as vq_gather_kernel: q_latent + beta * e_latent (vector_quantizer.py:72-83);
// optionally writes the straight-through value z + (q - z) (:148)
template <int DS>
__device__ __forceinline__ float vq_loss_fixed(const float* xr, const float* __restrict__ q, float beta, int squared, float* zsg) {
  float acc[8];
#pragma unroll
  for (int l = 0; l < 8; ++l) acc[l] = 0.f;
#pragma unroll
  for (int dd = 0; dd < DS; ++dd) {
    float df = __fsub_rn(q[dd], xr[dd]);
    if (zsg) zsg[dd] = __fadd_rn(xr[dd], df);
    if (!squared) df = __fadd_rn(df, 1e-5f);
    acc[dd & 7] = __fadd_rn(acc[dd & 7], __fmul_rn(df, df));
  }
  float s = acc[0];
#pragma unroll
  for (int l = 1; l < 8; ++l) s = __fadd_rn(s, acc[l]);
  float ql = squared ? s : sqrtf(s);
  return __fadd_rn(ql, __fmul_rn(beta, ql));
}
