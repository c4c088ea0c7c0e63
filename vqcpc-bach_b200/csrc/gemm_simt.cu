// Exact-fp32 GEMM on CUDA cores ("fp32" arithmetic mode) with the fused epilogues the encoder needs.
//
//   C[m,n] = epi( sum_k A(m,k) * B(k,n) )
//     a_kmajor = 1 : A(m,k) = A[m*lda + k]      (activations, row-major M x K)
//     a_kmajor = 0 : A(m,k) = A[k*lda + m]      (transposed view, weight-gradient GEMMs)
//     b_kmajor = 1 : B(k,n) = B[n*ldb + k]      (nn.Linear weight layout N x K)
//     b_kmajor = 0 : B(k,n) = B[k*ldb + n]
//   epi: (+bias[n]) -> (cols < scale_ncols) *= alpha -> (+residual[m,n]) -> relu -> (* (mask[m,n] > 0))
//        then store, accumulate (C +=) or atomic accumulate (split-K).
//
// This is the precision reference inside the product (fp32 FMA, k ascending per split) and the kernel used for
// tiny/odd shapes (K=3 upscaler input, N=3 output_linear); the large GEMMs run on tcgen05 (gemm_tc.cu).
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, EpilogueParams ep, int k_per_split) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int ty = tid / 16, tx = tid % 16;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile (BM x BK) into As[k][m]
    if (A_KMAJOR) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        int e = tid + it * 256;          // 0..2047
        int m = e / BK, k = e % BK;
        int gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < M && gk < kend) ? A[(size_t)gm * lda + gk] : 0.f;
      }
    } else {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        int e = tid + it * 256;
        int k = e / BM, m = e % BM;
        int gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < M && gk < kend) ? A[(size_t)gk * lda + gm] : 0.f;
      }
    }
    if (B_KMAJOR) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        int e = tid + it * 256;
        int n = e / BK, k = e % BK;
        int gn = n0 + n, gk = k0 + k;
        Bs[k][n] = (gn < N && gk < kend) ? B[(size_t)gn * ldb + gk] : 0.f;
      }
    } else {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        int e = tid + it * 256;
        int k = e / BN, n = e % BN;
        int gn = n0 + n, gk = k0 + k;
        Bs[k][n] = (gn < N && gk < kend) ? B[(size_t)gk * ldb + gn] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool first_split = (blockIdx.z == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float v = epilogue_apply(acc[i][j], gm, gn, ep, first_split);
      float* dst = C + (size_t)gm * ldc + gn;
      if (ep.flags & EPI_ATOMIC) atomicAdd(dst, v);
      else if (ep.flags & EPI_ACCUMULATE) *dst += v;
      else *dst = v;
    }
  }
}

}  // namespace

VQ_API int vqcpc_gemm_f32(int a_kmajor, int b_kmajor, int M, int N, int K, const float* A, int lda, const float* B,
                          int ldb, float* C, int ldc, const float* bias, const float* residual, int ldr, int flags,
                          float alpha, int scale_ncols, int ksplit, void* stream) {
  VQ_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm_f32: bad shape M=%d N=%d K=%d", M, N, K);
  VQ_CHECK_ARG(A && B && C, "gemm_f32: null operand");
  if (M == 0) return VQ_OK;
  EpilogueParams ep;
  ep.bias = bias; ep.residual = residual; ep.ldr = ldr; ep.flags = flags; ep.alpha = alpha; ep.scale_ncols = scale_ncols;
  VQ_CHECK_ARG(!(flags & (EPI_RESIDUAL | EPI_RELU_MASK)) || residual, "gemm_f32: residual/mask flag without pointer");
  VQ_CHECK_ARG(!(flags & EPI_BIAS) || bias, "gemm_f32: bias flag without pointer");
  if (ksplit < 1) ksplit = 1;
  int k_per_split = (int)vq_cdiv(vq_cdiv(K, ksplit), BK) * BK;
  ksplit = (int)vq_cdiv(K, k_per_split);
  if (ksplit > 1) {
    VQ_CHECK_ARG(!(flags & (EPI_RELU | EPI_RELU_MASK)), "gemm_f32: split-K cannot fuse a non-linear epilogue");
    ep.flags |= EPI_ATOMIC;
    if (!(flags & EPI_ACCUMULATE))   // partial sums are atomically added: start from zero unless accumulating
      VQ_CUDA_CALL(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, vq_stream(stream)));
  }
  dim3 grid((unsigned)vq_cdiv(M, BM), (unsigned)vq_cdiv(N, BN), (unsigned)ksplit);
  VQ_CHECK_ARG(grid.y <= 65535u && grid.z <= 65535u, "gemm_f32: N or ksplit too large (N=%d ksplit=%d)", N, ksplit);
  cudaStream_t st = vq_stream(stream);
  if (a_kmajor && b_kmajor) gemm_simt_kernel<true, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, k_per_split);
  else if (a_kmajor && !b_kmajor) gemm_simt_kernel<true, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, k_per_split);
  else if (!a_kmajor && b_kmajor) gemm_simt_kernel<false, true><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, k_per_split);
  else gemm_simt_kernel<false, false><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, ep, k_per_split);
  VQ_CUDA_LAUNCH_CHECK("gemm_simt_kernel");
  return VQ_OK;
}
