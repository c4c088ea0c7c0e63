// GRU cell point-wise stages (torch.nn.GRU semantics: CModule, vqcpc_helper.py:54-76; LstmDownscaler,
// downscalers/lstm_downscaler.py:31-48).  The matrix products of the recurrence run on the library's own GEMMs
//   GI = X W_ih^T + b_ih          one GEMM over all (b, t) rows of a layer
//   GH = h_{t-1} W_hh^T + b_hh    one GEMM per time step
// and these kernels do the rest of step t, gate order (r, z, n):
//   r = sigmoid(GI_r + GH_r),  z = sigmoid(GI_z + GH_z),  n = tanh(GI_n + r * GH_n),  h_t = (1 - z) * n + z * h_{t-1}.
// Layouts: batch-first.  GI (B*T, 3H) rows b*T + t;  GH (B, 3H);  out (B, T, H) (h_t is written to out[:, t]; the next
// step's GEMM reads it there with row stride T*H);  gates (B*T, 3H) saves (r, z, n) for backward, ghn (B*T, H) saves GH_n.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void gru_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ hprev,
                                    int64_t hprev_ld, float* __restrict__ out, float* __restrict__ gates,
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
    out[row * H + j] = (1.f - z) * n + z * hp;
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
    dgh[row * 3 * H + j] = dr;
    dgh[row * 3 * H + H + j] = dz;
    dgh[row * 3 * H + 2 * H + j] = dn * r;
    dh[i] = g * z;
  }
}

inline int gate_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(n, 256), (int64_t)vq_sm_count() * 8)); }

}  // namespace

// h_prev: (B, H) view with row stride hprev_ld (NULL = zero initial state).  gates / ghn may be NULL (inference).
VQ_API int vqcpc_gru_gate_fwd(const float* gi, const float* gh, const float* h_prev, int64_t hprev_ld, float* out, float* gates,
                              float* ghn, int B, int T, int H, int t, void* stream) {
  VQ_CHECK_ARG(t >= 0 && t < T && (gates == nullptr) == (ghn == nullptr), "gru_gate_fwd: bad step %d / %d or gates without ghn", t, T);
  if (B == 0) return VQ_OK;
  gru_gate_fwd_kernel<<<gate_grid((int64_t)B * H), 256, 0, vq_stream(stream)>>>(gi, gh, h_prev, hprev_ld, out, gates, ghn, B, T, H, t);
  VQ_CUDA_LAUNCH_CHECK("gru_gate_fwd_kernel");
  return VQ_OK;
}

VQ_API int vqcpc_gru_gate_bwd(const float* gates, const float* ghn, const float* h_prev, int64_t hprev_ld, float* dh,
                              const float* dout, float* dgi, float* dgh, int B, int T, int H, int t, void* stream) {
  VQ_CHECK_ARG(t >= 0 && t < T, "gru_gate_bwd: bad step %d / %d", t, T);
  if (B == 0) return VQ_OK;
  gru_gate_bwd_kernel<<<gate_grid((int64_t)B * H), 256, 0, vq_stream(stream)>>>(gates, ghn, h_prev, hprev_ld, dh, dout, dgi, dgh, B, T, H, t);
  VQ_CUDA_LAUNCH_CHECK("gru_gate_bwd_kernel");
  return VQ_OK;
}
