// Micro-benchmark: global-store throughput per SM for the access patterns an epilogue warp can use to write a
// 32-row x 256-column fp32 slab (row stride ld floats).  One CTA of 128 threads per SM, persistent over slabs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/store_bench tools/micro/store_bench.cu && build/store_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int PAT>
__global__ void __launch_bounds__(128, 1) k(float* out, int ld, long long slabs, int with_load, const float* res) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 v = make_float4(lane, warp, 1.f, 2.f);
  for (long long s = (long long)blockIdx.x * 4 + warp; s < slabs; s += (long long)gridDim.x * 4) {
    const long long mslab = s / (ld / 256), nslab = s % (ld / 256);
    float* base = out + mslab * 32 * ld + nslab * 256;
    const float* rb = res + mslab * 32 * ld + nslab * 256;
#pragma unroll 1
    for (int ch = 0; ch < 8; ++ch) {            // 32-column chunks
      float* cb = base + ch * 32;
      const float* rc = rb + ch * 32;
      if (PAT == 0) {                           // thread = row: 8 x 16 B per lane, 32 lines per instruction
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 w = v;
          if (with_load) { float4 r = __ldg(reinterpret_cast<const float4*>(rc + (size_t)lane * ld + i * 4)); w.x += r.x; w.y += r.y; w.z += r.z; w.w += r.w; }
          *reinterpret_cast<float4*>(cb + (size_t)lane * ld + i * 4) = w;
        }
      } else if (PAT == 1) {                    // quad: 8 rows x 64 B per instruction
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = (lane >> 2) + 8 * (i >> 1), col = (lane & 3) * 4 + 16 * (i & 1);
          float4 w = v;
          if (with_load) { float4 r = __ldg(reinterpret_cast<const float4*>(rc + (size_t)row * ld + col)); w.x += r.x; w.y += r.y; w.z += r.z; w.w += r.w; }
          *reinterpret_cast<float4*>(cb + (size_t)row * ld + col) = w;
        }
      } else if (PAT == 2) {                    // full lines: 4 rows x 128 B per instruction
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = (lane >> 3) + 4 * i, col = (lane & 7) * 4;
          float4 w = v;
          if (with_load) { float4 r = __ldg(reinterpret_cast<const float4*>(rc + (size_t)row * ld + col)); w.x += r.x; w.y += r.y; w.z += r.z; w.w += r.w; }
          *reinterpret_cast<float4*>(cb + (size_t)row * ld + col) = w;
        }
      }
    }
    if (PAT == 3) {                             // one row x 512 B per instruction (whole 1 KB row in two instructions)
#pragma unroll 1
      for (int row = 0; row < 32; ++row)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float4 w = v;
          const int col = lane * 4 + 128 * i;
          if (with_load) { float4 r = __ldg(reinterpret_cast<const float4*>(rb + (size_t)row * ld + col)); w.x += r.x; w.y += r.y; w.z += r.z; w.w += r.w; }
          *reinterpret_cast<float4*>(base + (size_t)row * ld + col) = w;
        }
    }
  }
}

int main() {
  const int ld = 1536;
  const long long M = 139264 * 2;
  float *out, *res;
  cudaMalloc(&out, (size_t)M * ld * 4);
  cudaMalloc(&res, (size_t)M * ld * 4);
  cudaMemset(res, 0, (size_t)M * ld * 4);
  const long long slabs = (M / 32) * (ld / 256);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int with_load = 0; with_load < 2; ++with_load)
    for (int pat = 0; pat < 4; ++pat) {
      float best = 1e30f;
      for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        if (pat == 0) k<0><<<sms, 128>>>(out, ld, slabs, with_load, res);
        if (pat == 1) k<1><<<sms, 128>>>(out, ld, slabs, with_load, res);
        if (pat == 2) k<2><<<sms, 128>>>(out, ld, slabs, with_load, res);
        if (pat == 3) k<3><<<sms, 128>>>(out, ld, slabs, with_load, res);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double bytes = (double)M * ld * 4;
      printf("pattern %d (%s) load=%d: %.3f ms  %.2f TB/s written  %.1f B/us/SM\n", pat,
             pat == 0 ? "thread=row, 32 lines/instr" : pat == 1 ? "quad, 8 rows x 64 B" : pat == 2 ? "4 rows x 128 B" : "1 row x 512 B",
             with_load, best, bytes / best / 1e9, bytes / best / 1e3 / sms);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
