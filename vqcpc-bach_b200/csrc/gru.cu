// GRU cell point-wise stages (torch.nn.GRU semantics: CModule, vqcpc_helper.py:54-76; LstmDownscaler,
// downscalers/lstm_downscaler.py:31-48).  The matrix products of the recurrence run on the library's own GEMMs
//   GI = X W_ih^T + b_ih          one GEMM over all (b, t) rows of a layer
//   GH = h_{t-1} W_hh^T + b_hh    one GEMM per time step
// and these kernels do the rest of step t, gate order (r, z, n):
//   r = sigmoid(GI_r + GH_r),  z = sigmoid(GI_z + GH_z),  n = tanh(GI_n + r * GH_n),  h_t = (1 - z) * n + z * h_{t-1}.
// Layouts: batch-first.  GI (B*T, 3H) rows b*T + t;  GH (B, 3H);  out (B, T, H) (h_t is written to out[:, t]; the next
// step's GEMM reads it there with row stride T*H);  gates (B*T, 3H) saves (r, z, n) for backward, ghn (B*T, H) saves GH_n.
// Both kernels can also emit the bf16 (hi, lo) split of what the next tensor-core GEMM consumes (h_t; dGH_t), so the
// recurrence needs no separate operand-split launches.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void gru_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ hprev,
                                    int64_t hprev_ld, float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                                    __nv_bfloat16* __restrict__ out_lo, float* __restrict__ gates,
                                    float* __restrict__ ghn, int B, int T, int H, int t) {
  const int64_t total = (int64_t)B * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), j = (int)(i % H);
    const int64_t row = (int64_t)b * T + t;
    const float* gir = gi + row * 3 * H;
    const float* ghr = gh + (int64_t)b * 3 * H;
    const float hp = hprev ? hprev[(int64_t)b * hprev_ld + j] : 0.f;
    const float r = sigmoidf_(gir[j] + ghr[j]);
    const float z = sigmoidf_(gir[H + j] + ghr[H + j]);
    const float gn = ghr[2 * H + j];
    const float n = tanhf(gir[2 * H + j] + r * gn);
    const float hn = (1.f - z) * n + z * hp;
    out[row * H + j] = hn;
    if (out_hi) {
      const __nv_bfloat16 hh = __float2bfloat16_rn(hn);
      out_hi[row * H + j] = hh;
      out_lo[row * H + j] = __float2bfloat16_rn(hn - __bfloat162float(hh));
    }
    if (gates) {
      gates[row * 3 * H + j] = r;
      gates[row * 3 * H + H + j] = z;
      gates[row * 3 * H + 2 * H + j] = n;
      ghn[row * H + j] = gn;
    }
  }
}

// backward of step t.  dh (B, H) holds the recurrent part of the gradient wrt h_t on entry; the upstream gradient
// dout[:, t] ((B, T, H), may be NULL) is added here.  dh is overwritten with the direct part of the gradient wrt h_{t-1}: dh * z (the GEMM dGH W_hh is added by the caller).
// dgi / dgh (B*T, 3H) rows b*T + t receive the gradients of GI_t / GH_t.
__global__ void gru_gate_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ ghn,
                                    const float* __restrict__ hprev, int64_t hprev_ld, float* __restrict__ dh,
                                    const float* __restrict__ dout, float* __restrict__ dgi, float* __restrict__ dgh,
                                    __nv_bfloat16* __restrict__ dgh_hi, __nv_bfloat16* __restrict__ dgh_lo,
                                    int B, int T, int H, int t) {
  const int64_t total = (int64_t)B * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), j = (int)(i % H);
    const int64_t row = (int64_t)b * T + t;
    const float r = gates[row * 3 * H + j], z = gates[row * 3 * H + H + j], n = gates[row * 3 * H + 2 * H + j];
    const float gn = ghn[row * H + j];
    const float hp = hprev ? hprev[(int64_t)b * hprev_ld + j] : 0.f;
    const float g = dh[i] + (dout ? dout[row * H + j] : 0.f);
    const float dn = g * (1.f - z) * (1.f - n * n);          // through tanh
    const float dz = g * (hp - n) * z * (1.f - z);           // through sigmoid
    const float dr = dn * gn * r * (1.f - r);
    dgi[row * 3 * H + j] = dr;
    dgi[row * 3 * H + H + j] = dz;
    dgi[row * 3 * H + 2 * H + j] = dn;
    const float dhn = dn * r;
    if (dgh) {
      dgh[row * 3 * H + j] = dr;
      dgh[row * 3 * H + H + j] = dz;
      dgh[row * 3 * H + 2 * H + j] = dhn;
    }
    if (dgh_hi) {
      const float v3[3] = {dr, dz, dhn};
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const __nv_bfloat16 hh = __float2bfloat16_rn(v3[q]);
        dgh_hi[row * 3 * H + q * H + j] = hh;
        dgh_lo[row * 3 * H + q * H + j] = __float2bfloat16_rn(v3[q] - __bfloat162float(hh));
      }
    }
    dh[i] = g * z;
  }
}

inline int gate_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(n, 256), (int64_t)vq_sm_count() * 8)); }

}  // namespace

// h_prev: (B, H) view with row stride hprev_ld (NULL = zero initial state).  gates / ghn may be NULL (inference).
VQ_API int vqcpc_gru_gate_fwd(const float* gi, const float* gh, const float* h_prev, int64_t hprev_ld, float* out, void* out_hi,
                              void* out_lo, float* gates, float* ghn, int B, int T, int H, int t, void* stream) {
  VQ_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr), "gru_gate_fwd: out_hi and out_lo go together");
  VQ_CHECK_ARG(t >= 0 && t < T && (gates == nullptr) == (ghn == nullptr), "gru_gate_fwd: bad step %d / %d or gates without ghn", t, T);
  if (B == 0) return VQ_OK;
  gru_gate_fwd_kernel<<<gate_grid((int64_t)B * H), 256, 0, vq_stream(stream)>>>(
      gi, gh, h_prev, hprev_ld, out, reinterpret_cast<__nv_bfloat16*>(out_hi), reinterpret_cast<__nv_bfloat16*>(out_lo), gates, ghn, B, T, H, t);
  VQ_CUDA_LAUNCH_CHECK("gru_gate_fwd_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_gru_gate_bwd(const float* gates, const float* ghn, const float* h_prev, int64_t hprev_ld, float* dh,
                              const float* dout, float* dgi, float* dgh, void* dgh_hi, void* dgh_lo, int B, int T, int H, int t,
                              void* stream) {
  VQ_CHECK_ARG(t >= 0 && t < T, "gru_gate_bwd: bad step %d / %d", t, T);
  VQ_CHECK_ARG((dgh || dgh_hi) && ((dgh_hi == nullptr) == (dgh_lo == nullptr)), "gru_gate_bwd: need d_gh and / or its split");
  if (B == 0) return VQ_OK;
  gru_gate_bwd_kernel<<<gate_grid((int64_t)B * H), 256, 0, vq_stream(stream)>>>(
      gates, ghn, h_prev, hprev_ld, dh, dout, dgi, dgh, reinterpret_cast<__nv_bfloat16*>(dgh_hi), reinterpret_cast<__nv_bfloat16*>(dgh_lo), B, T, H, t);
  VQ_CUDA_LAUNCH_CHECK("gru_gate_bwd_kernel");
  return VQ_OK;
}
