// Optimizer tail on the flat parameter / gradient buffers: global-norm clip + Adam in two launches, no host sync.
//
// Reference: nn.utils.clip_grad_norm_(self.parameters(), 5) then torch.optim.Adam(lr).step()
// (VQCPCB/vqcpc_encoder_trainer.py:313-314, Adam defaults beta=(0.9, 0.999), eps=1e-8, no weight decay).
#include "common.cuh"

namespace {

constexpr int SUMSQ_BLOCKS = 1024;

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ partial) {
  __shared__ double red[8];
  double a = 0.0;
  int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = ld4(g + i * 4);
    a += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) a += (double)g[i] * g[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

// out[0] = grad_scale^2 * sum(partials)  (the squared global norm of the scaled gradient), fixed order
__global__ void sumsq_final_kernel(const double* __restrict__ partial, int nblocks, float grad_scale, float* __restrict__ out) {
  __shared__ double red[32];
  double a = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) a += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    out[0] = (float)(s * (double)grad_scale * (double)grad_scale);
  }
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 const float* __restrict__ gnorm_sq, float max_norm, float grad_scale, float lr, float b1, float b2,
                 float eps, float bc1, float bc2_sqrt) {
  float coef = grad_scale;
  if (max_norm > 0.f) {
    float total = sqrtf(gnorm_sq[0]);
    coef *= fminf(1.f, max_norm / (total + 1e-6f));          // clip_grad_norm_ semantics
  }
  const float step_size = lr / bc1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * coef;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

double* g_partials = nullptr;

}  // namespace

// out[0] = || grad_scale * g ||_2^2 over the flat buffer (deterministic two-stage fp64 reduction)
VQ_API int vqcpc_grad_sumsq(const float* g, int64_t n, float grad_scale, float* out, void* stream) {
  VQ_CHECK_ARG((reinterpret_cast<uintptr_t>(g) & 15) == 0, "grad_sumsq: flat gradient buffer must be 16-byte aligned");
  if (!g_partials) VQ_CUDA_CALL(cudaMalloc(&g_partials, SUMSQ_BLOCKS * sizeof(double)));
  int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(SUMSQ_BLOCKS, vq_cdiv(n / 4 + 1, 256)));
  cudaStream_t st = vq_stream(stream);
  sumsq_partial_kernel<<<blocks, 256, 0, st>>>(g, n, g_partials);
  VQ_CUDA_LAUNCH_CHECK("sumsq_partial_kernel");
  sumsq_final_kernel<<<1, 256, 0, st>>>(g_partials, blocks, grad_scale, out);
  VQ_CUDA_LAUNCH_CHECK("sumsq_final_kernel");
  return VQ_OK;
}

// p, m, v updated in place.  g is read as grad_scale * g (grad_scale = 1/world_size after a sum-allreduce), clipped by
// min(1, max_norm / (sqrt(gnorm_sq[0]) + 1e-6)) when max_norm > 0 (gnorm_sq is a DEVICE scalar from vqcpc_grad_sumsq).
// step is the 1-based Adam step count (bias corrections computed on the host in double).
VQ_API int vqcpc_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* gnorm_sq, float max_norm,
                           float grad_scale, float lr, float beta1, float beta2, float eps, int step, void* stream) {
  VQ_CHECK_ARG(step >= 1, "clip_adam: step must be >= 1");
  VQ_CHECK_ARG(max_norm <= 0.f || gnorm_sq != nullptr, "clip_adam: clipping needs the device norm scalar");
  if (n == 0) return VQ_OK;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(n, 256), (int64_t)vq_sm_count() * 8));
  clip_adam_kernel<<<blocks, 256, 0, vq_stream(stream)>>>(p, g, m, v, n, gnorm_sq, max_norm, grad_scale, lr, beta1, beta2, eps,
                                                         (float)bc1, (float)sqrt(bc2));
  VQ_CUDA_LAUNCH_CHECK("clip_adam_kernel");
  return VQ_OK;
}
