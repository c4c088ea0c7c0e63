// Batch assembly of the CPC data path from a chorale tensor dataset that is RESIDENT in HBM (SURVEY.md §8f row 4:
// datasets/chorale_dataset.py:542-595 `data_loaders` over the cached `tensor_dataset`, dataloaders/bach_cpc_dataloader.py:110-252).
// The reference slices host tensors per batch (DataLoader collate -> slicing -> transpose -> pinned copy); here a step only
// ships the sampled row indices and the token windows are gathered on the device, voice dimension last, int64:
//     out[r, e, c] = data[rows[r], c, t0 + e]                                             (windows: x_left, x_right, 'random' negatives)
//     out[b, m, k, e, c] = data[rows[b], c, block(m, k) * evb + e]                          ('same_sequence' negatives)
// HBM-bound integer work: 4 (or 8) bytes read + 8 bytes written per token.
#include "common.cuh"

namespace {

// One thread per (row, event): the V voices of an event are V strided reads (consecutive threads read consecutive ticks of the
// same voice row: coalesced) and ONE contiguous V * 8-byte store.
template <typename T, int VV>
__device__ __forceinline__ void store_event(const T* __restrict__ src, int ticks, int V, int64_t* __restrict__ dst) {
  if (VV == 4) {
    const longlong2 a = make_longlong2((long long)src[0], (long long)src[ticks]);
    const longlong2 b = make_longlong2((long long)src[2 * (int64_t)ticks], (long long)src[3 * (int64_t)ticks]);
    reinterpret_cast<longlong2*>(dst)[0] = a;
    reinterpret_cast<longlong2*>(dst)[1] = b;
  } else {
    for (int c = 0; c < V; ++c) dst[c] = (int64_t)src[c * (int64_t)ticks];
  }
}

template <typename T, int VV>
__global__ void __launch_bounds__(256)
cpc_window_gather_kernel(const T* __restrict__ data, int V, int ticks, const int64_t* __restrict__ rows, int64_t n, int t0,
                         int n_ev, int64_t* __restrict__ out) {
  const int64_t total = n * n_ev;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / n_ev;
    const int e = (int)(t - r * n_ev);
    store_event<T, VV>(data + rows[r] * V * (int64_t)ticks + t0 + e, ticks, V, out + t * V);
  }
}

// bach_cpc_dataloader.py:162-184 `_build_negatives_sameSeq(x_a, x_b)`: for every k the negatives of position k are all blocks of
// the first part followed by the blocks of the second part except block k.  The window holds part A at ticks [a0, a0 + KA * evb)
// and part B at [b0, b0 + KB * evb); M = KA + KB - 1; k runs over nk <= KB positions (the reference always loops over
// num_blocks_right, also for the backward negatives where part B is the left half).
template <typename T, int VV>
__global__ void __launch_bounds__(256)
cpc_same_seq_kernel(const T* __restrict__ data, int V, int ticks, const int64_t* __restrict__ rows, int64_t B, int a0, int KA,
                    int b0, int KB, int nk, int evb, int64_t* __restrict__ out) {
  const int M = KA + KB - 1;
  const int per_row = M * nk * evb;
  const int64_t total = B * per_row;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / per_row;
    int q = (int)(t - b * per_row);
    const int e = q % evb; q /= evb;
    const int k = q % nk;
    const int m = q / nk;
    int tick;
    if (m < KA) tick = a0 + m * evb + e;
    else {
      const int j = m - KA;
      tick = b0 + (j < k ? j : j + 1) * evb + e;
    }
    store_event<T, VV>(data + rows[b] * V * (int64_t)ticks + tick, ticks, V, out + t * V);
  }
}

unsigned gather_grid(int64_t total) {
  const int64_t want = vq_cdiv(total, 256 * 2);
  const int64_t cap = (int64_t)vq_sm_count() * 16;
  return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

// data: (n_rows, V, ticks) int32 (elem_bytes 4: the reference caches `.int()` tensors, chorale_dataset.py:166) or int64 (8);
// rows (n) int64 row ids; out (n, n_ev, V) int64.
VQ_API int vqcpc_cpc_window_gather(const void* data, int elem_bytes, int64_t n_rows, int V, int ticks, const int64_t* rows, int64_t n,
                                   int t0, int n_ev, int64_t* out, void* stream) {
  VQ_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "cpc_window_gather: elem_bytes must be 4 or 8");
  VQ_CHECK_ARG(V >= 1 && ticks >= 1 && t0 >= 0 && n_ev >= 1 && t0 + n_ev <= ticks && n_rows >= 1, "cpc_window_gather: window [%d, %d) outside %d ticks", t0, t0 + n_ev, ticks);
  if (n == 0) return VQ_OK;
  const int64_t total = n * n_ev;
  const unsigned grid = gather_grid(total);
  cudaStream_t st = vq_stream(stream);
#define VQ_WG(TT, VVV) cpc_window_gather_kernel<TT, VVV><<<grid, 256, 0, st>>>((const TT*)data, V, ticks, rows, n, t0, n_ev, out)
  if (elem_bytes == 4) { if (V == 4) VQ_WG(int, 4); else VQ_WG(int, 0); }
  else { if (V == 4) VQ_WG(int64_t, 4); else VQ_WG(int64_t, 0); }
#undef VQ_WG
  VQ_CUDA_LAUNCH_CHECK("cpc_window_gather_kernel");
  return VQ_OK;
}

// out (B, KA + KB - 1, nk, evb, V) int64.
VQ_API int vqcpc_cpc_same_seq_negatives(const void* data, int elem_bytes, int64_t n_rows, int V, int ticks, const int64_t* rows, int64_t B,
                                        int a0, int KA, int b0, int KB, int nk, int evb, int64_t* out, void* stream) {
  VQ_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "cpc_same_seq_negatives: elem_bytes must be 4 or 8");
  VQ_CHECK_ARG(V >= 1 && KA >= 1 && KB >= 1 && nk >= 1 && nk <= KB && evb >= 1 && a0 >= 0 && b0 >= 0 && a0 + KA * evb <= ticks && b0 + KB * evb <= ticks && n_rows >= 1,
               "cpc_same_seq_negatives: blocks outside %d ticks", ticks);
  if (B == 0) return VQ_OK;
  const int64_t total = B * (KA + KB - 1) * nk * (int64_t)evb;
  const unsigned grid = gather_grid(total);
  cudaStream_t st = vq_stream(stream);
#define VQ_SS(TT, VVV) cpc_same_seq_kernel<TT, VVV><<<grid, 256, 0, st>>>((const TT*)data, V, ticks, rows, B, a0, KA, b0, KB, nk, evb, out)
  if (elem_bytes == 4) { if (V == 4) VQ_SS(int, 4); else VQ_SS(int, 0); }
  else { if (V == 4) VQ_SS(int64_t, 4); else VQ_SS(int64_t, 0); }
#undef VQ_SS
  VQ_CUDA_LAUNCH_CHECK("cpc_same_seq_kernel");
  return VQ_OK;
}
