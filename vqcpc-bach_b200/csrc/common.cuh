// Shared helpers for the vqcpc_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#define VQ_API extern "C" __attribute__((visibility("default")))

// error codes returned through the C ABI
enum : int {
  VQ_OK = 0,
  VQ_ERR_INVALID = -1,   // bad argument / unsupported shape
  VQ_ERR_CUDA = -2,      // CUDA runtime error (see vqcpc_last_error)
  VQ_ERR_DRIVER = -3,    // driver entry point (TMA descriptor encode) unavailable
};

void vq_set_error(const char* fmt, ...);

#define VQ_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      vq_set_error(__VA_ARGS__);                \
      return VQ_ERR_INVALID;                    \
    }                                           \
  } while (0)

#define VQ_CUDA_LAUNCH_CHECK(name)                                                        \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      vq_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));               \
      return VQ_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define VQ_CUDA_CALL(expr)                                                                \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      vq_set_error("%s failed: %s", #expr, cudaGetErrorString(e__));                      \
      return VQ_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

static inline cudaStream_t vq_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t vq_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
int vq_sm_count();

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
