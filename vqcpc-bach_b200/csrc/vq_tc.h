// Internal interface of the tensor-core nearest-code search (vq_tc.cu), used by the C-ABI entry points in quantizer.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

// shapes the tcgen05 search handles (others stay on the direct-form kernel)
bool vq_tc_supported(int D, int Kc);
// same contract as vqcpc_vq_fused_fwd_rows (y covers the rows [y_row0, y_row0 + y_rows) only); dbg_scores (n, ceil(Kc/256)*256) and stats[4098] (two counters + 8 clock64 stamps
// per CTA) are optional test / profiling hooks
int vq_tc_fused_fwd(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta, int squared,
                    int64_t* idx, float* loss, float* z_sg, float* y, int64_t y_row0, int64_t y_rows, int* counts,
                    float* dbg_scores, int* stats, int variant, cudaStream_t st);
