// Epilogue shared by the CUDA-core and the tcgen05 GEMMs.
#pragma once
#include "common.cuh"

enum : int {
  EPI_BIAS = 1,        // + bias[n]
  EPI_RELU = 2,        // max(x, 0)
  EPI_ACCUMULATE = 4,  // C += result (non-atomic; one writer per element)
  EPI_RESIDUAL = 8,    // + residual[m, n]
  EPI_RELU_MASK = 16,  // * (residual[m, n] > 0)   (ReLU backward fused into the dgrad GEMM; uses the residual pointer)
  EPI_ATOMIC = 32,     // internal: split-K partial -> atomicAdd into C
  EPI_SCALE_COLS = 64, // columns n < scale_ncols multiplied by alpha after the bias (q *= hd^-0.5)
  EPI_MASK_BF16 = 128, // the EPI_RELU_MASK source is a bf16 matrix (the hi half of a split activation), not fp32
};

struct EpilogueParams {
  const float* bias;
  const float* residual;
  int ldr;
  int flags;
  float alpha;
  int scale_ncols;
};

// first_split: bias / residual / column scaling are applied by the split that owns k = 0 only.
__device__ __forceinline__ float epilogue_apply(float v, int m, int n, const EpilogueParams& ep, bool first_split) {
  if (first_split && (ep.flags & EPI_BIAS)) v += __ldg(ep.bias + n);
  if ((ep.flags & EPI_SCALE_COLS) && n < ep.scale_ncols) v *= ep.alpha;
  if (first_split && (ep.flags & EPI_RESIDUAL)) v += __ldg(ep.residual + (size_t)m * ep.ldr + n);
  if (ep.flags & EPI_RELU) v = fmaxf(v, 0.f);
  if (ep.flags & EPI_RELU_MASK) {
    bool pos;
    if (ep.flags & EPI_MASK_BF16)
      pos = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ep.residual)[(size_t)m * ep.ldr + n]) > 0.f;
    else
      pos = __ldg(ep.residual + (size_t)m * ep.ldr + n) > 0.f;
    v = pos ? v : 0.f;
  }
  return v;
}
