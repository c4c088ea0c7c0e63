/* vqcpc_b200 — C ABI of libvqcpc_b200.so (sm_100a).
 *
 * Drop-in boundary for the VQ-CPC encoder-training hot path of SonyCSLParis/vqcpc-bach.  The reference is pure
 * Python/PyTorch and has no FFI of its own; these entry points are what its Python classes bind (through ctypes,
 * see INTEGRATION.md) in place of the ATen calls cited next to each function (paths relative to the reference repo).
 *
 * Conventions
 *   - plain pointers and sizes, no torch types; every pointer is DEVICE memory unless marked "host";
 *   - the caller owns all buffers, kernels never allocate (one lazily allocated 8 KB scratch in vqcpc_grad_sumsq);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous and stream-ordered;
 *   - return value: 0 = ok, <0 = error (VQ_ERR_*); vqcpc_last_error() returns the message (thread local);
 *   - "ACCUMULATED" outputs are += targets (gradient buffers); all others are overwritten;
 *   - matrices are row-major fp32 unless stated; `ld*` are leading dimensions in elements.
 */
#ifndef VQCPC_B200_H
#define VQCPC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQ_OK 0
#define VQ_ERR_INVALID (-1)
#define VQ_ERR_CUDA (-2)
#define VQ_ERR_DRIVER (-3)

/* GEMM epilogue flags */
#define VQ_EPI_BIAS 1        /* + bias[n] */
#define VQ_EPI_RELU 2        /* max(x, 0) */
#define VQ_EPI_ACCUMULATE 4  /* C += result */
#define VQ_EPI_RESIDUAL 8    /* + residual[m, n] */
#define VQ_EPI_RELU_MASK 16  /* * (residual[m, n] > 0): ReLU backward fused into the data-gradient GEMM */
#define VQ_EPI_SCALE_COLS 64 /* columns n < scale_ncols multiplied by alpha after the bias */
#define VQ_EPI_MASK_BF16 128 /* with VQ_EPI_RELU_MASK: the mask source is a bf16 matrix (hi half of a split activation) */

/* ---- runtime ------------------------------------------------------------------------------------------------ */
const char* vqcpc_last_error(void);
int vqcpc_abi_version(void);
int vqcpc_device_info(int* sm_count /*host*/, int* cc_major /*host*/, int* cc_minor /*host*/);

/* ---- data processor: per-voice nn.Embedding lookup, voice = position % V -------------------------------------- */
/* VQCPCB/data_processor/bach_cpc_data_processor.py:42-68, data_processor.py:26-32.
 * tokens (n) int64; tables: host array of V device pointers, table v is (rows[v], E); out (n, E). */
int vqcpc_embed_fwd(const int64_t* tokens, int64_t n, int V, const float* const* tables /*host*/, const int* rows /*host*/,
                    int E, float* out, void* stream);
/* d_table[v][tok] += d_out — grads: host array of V device pointers, ACCUMULATED. */
int vqcpc_embed_bwd(const int64_t* tokens, int64_t n, int V, float* const* grads /*host*/, const int* rows /*host*/, int E,
                    const float* d_out, void* stream);

/* ---- downscaler pieces ------------------------------------------------------------------------------------- */
/* VQCPCB/downscalers/relative_transformer_downscaler.py:106-112: h[m, col0 + j] = chan[(m%S)%V][j] (j < pe) |
 * evt[(m%S)/V][j - pe]; chan (V, pe), evt (S/V, pe).  The backward ACCUMULATES d_chan / d_evt. */
int vqcpc_posemb_fill(float* h, int ld, int col0, const float* chan, const float* evt, int64_t M, int S, int V, int pe,
                      void* stream);
int vqcpc_posemb_bwd(const float* dh, int ld, int col0, float* d_chan, float* d_evt, int64_t M, int S, int V, int pe,
                     void* stream);
/* nn.LayerNorm(d), eps inside the sqrt (transformer/transformer_custom.py:261-262, :283, :289); mean/rstd (M) saved. */
/* out (fp32) and/or (out_hi, out_lo) = the result pre-split into bf16 for the next tensor-core GEMM (either may be NULL). */
int vqcpc_layernorm_fwd(const float* y, const float* gamma, const float* beta, float* out, void* out_hi, void* out_lo,
                        float* mean, float* rstd, int64_t M, int d, float eps, void* stream);
/* d_y overwritten (+ optional bf16 split d_y_hi/d_y_lo); d_gamma / d_beta ACCUMULATED. */
int vqcpc_layernorm_bwd(const float* d_out, const float* y, const float* mean, const float* rstd, const float* gamma,
                        float* d_y, void* d_y_hi, void* d_y_lo, float* d_gamma, float* d_beta, int64_t M, int d,
                        void* stream);
/* Same; d_y_colsum (d floats, may be NULL) additionally ACCUMULATES the column sums of d_y = the bias gradient of the
 * linear layer feeding the LN (out_proj.bias for norm1, linear2.bias for norm2; transformer_custom.py:283-289). */
int vqcpc_layernorm_bwd_ex(const float* d_out, const float* y, const float* mean, const float* rstd, const float* gamma,
                           float* d_y, void* d_y_hi, void* d_y_lo, float* d_gamma, float* d_beta, float* d_y_colsum,
                           int64_t M, int d, void* stream);
/* relative_transformer_downscaler.py:124 `output[::downscaling]` in blocks-major layout: x (P*L, d) -> out (P*L/f, d). */
int vqcpc_subsample_fwd(const float* x, float* out, int64_t P, int L, int factor, int d, void* stream);
int vqcpc_subsample_bwd(const float* d_out, float* d_x, int64_t P, int L, int factor, int d, void* stream);
/* bias gradients: out[n] += sum_m x[m, n]  (ACCUMULATED). */
int vqcpc_colsum(const float* x, int ld, float* out, int64_t M, int N, void* stream);
/* same on a bf16 (hi, lo) split matrix: out[n] += sum_m (hi + lo)[m, n]. */
int vqcpc_colsum_split(const void* x_hi, const void* x_lo, int ld, float* out, int64_t M, int N, void* stream);
/* nn.SELU of the MLP upscaler (upscalers/mlp_upscaler.py:21-26). */
int vqcpc_selu_fwd(const float* x, float* out, int64_t n, void* stream);
int vqcpc_selu_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream);
/* nn.Dropout (mlp_upscaler.py:23) with an explicit keep mask (one byte per element, drawn by the caller so that
 * forward, backward and tests share it): out = keep ? x * scale : 0, scale = 1 / (1 - p). */
int vqcpc_mask_scale(const float* x, const unsigned char* keep, float scale, float* out, int64_t n, void* stream);
/* The dropout sites of the encoder layer (transformer_custom.py:278-290: dropout1 / dropout2 before the residual add,
 * `dropout` after the activation) and their backward: out = (keep ? x * scale : 0) (+ residual, may be NULL), written as
 * fp32 (`out`) and / or as the bf16 split (out_hi, out_lo) the next GEMM consumes; n % 4 == 0. */
int vqcpc_mask_scale_add(const float* x, const unsigned char* keep, float scale, const float* residual, float* out,
                         void* out_hi, void* out_lo, int64_t n, void* stream);
/* out[0] = scale * sum(x[0..n)) in a fixed order (deterministic). */
int vqcpc_reduce_sum(const float* x, int64_t n, float* out, float scale, void* stream);
/* Code-usage histogram: counts[idx[i]] += 1 for 0 <= idx[i] < bins (int32, ACCUMULATED).  Replaces the torch.unique
 * monitoring of VQCPCB/vqcpc_encoder_trainer.py:325-331 and feeds the per-code counts of the data-parallel all-reduce. */
int vqcpc_code_hist(const int64_t* idx, int64_t n, int bins, int* counts, void* stream);

/* ---- GEMMs (F.linear / matmul call sites: multihead_attention_custom.py:171,346; transformer_custom.py:285;
 *      relative_transformer_downscaler.py:104,132; mlp_upscaler.py:21-26; vqcpc_helper.py:69 and their backward) -------
 * C[M,N] = epi(sum_k A(m,k) B(k,n));  a_kmajor: A(m,k)=A[m*lda+k] else A[k*lda+m];  b_kmajor: B(k,n)=B[n*ldb+k] else
 * B[k*ldb+n].  epi: +bias -> scale cols -> +residual -> relu -> relu-mask; ksplit > 1 = split-K via fp32 atomics. */
int vqcpc_gemm_f32(int a_kmajor, int b_kmajor, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                   float* C, int ldc, const float* bias, const float* residual, int ldr, int flags, float alpha,
                   int scale_ncols, int ksplit, void* stream);
/* fp32 (rows, cols; ldx) -> bf16 hi, lo (ldo) with x ~= hi + lo: operand preparation for vqcpc_gemm_bf16x3. */
int vqcpc_split_bf16(const float* x, int ldx, void* hi, void* lo, int ldo, int64_t rows, int cols, void* stream);
/* tcgen05/TMEM/TMA tensor-core GEMM on pre-split bf16 operands; passes = 3 (hi.hi + hi.lo + lo.hi, parity mode) or
 * 1 (hi.hi, fast mode); fp32 accumulation and fp32 C; same layout/epilogue semantics as vqcpc_gemm_f32. */
int vqcpc_gemm_bf16x3(int a_kmajor, int b_kmajor, int M, int N, int K, const void* A_hi, const void* A_lo, int lda,
                      const void* B_hi, const void* B_lo, int ldb, float* C, int ldc, const float* bias,
                      const float* residual, int ldr, int flags, float alpha, int scale_ncols, int passes, int ksplit,
                      void* stream);

/* extended form: also (or only, C == NULL) writes the result pre-split into bf16 (out_hi, out_lo; ldo) for the next GEMM;
 * `residual` is fp32, or bf16 when VQ_EPI_MASK_BF16 accompanies VQ_EPI_RELU_MASK. */
int vqcpc_gemm_bf16x3_ex(int a_kmajor, int b_kmajor, int M, int N, int K, const void* A_hi, const void* A_lo, int lda,
                         const void* B_hi, const void* B_lo, int ldb, float* C, int ldc, void* out_hi, void* out_lo, int ldo,
                         const float* bias, const void* residual, int ldr, int flags, float alpha, int scale_ncols,
                         int passes, int ksplit, void* stream);

/* tuning / test hook: force the tcgen05 tile configuration: 0 = 128x128x64, 1 = 128x256x32, 2 / 3 = 128x256x32 with the
 * B tile TMA-multicast over a cluster of 2 / 4 CTAs, 4 = CTA pairs (cta_group::2, 256x256 tiles), 5 = 4 with eight epilogue
 * warps per CTA (experimental), -1 = automatic per shape. */
int vqcpc_gemm_set_tile_config(int cfg);

/* which epilogue variant of the tcgen05 GEMM the library was built with (-DVQ_EPI_KIND): 0 = quad (mma-fragment TMEM loads +
 * one shuffle exchange, 8 cache lines per warp access; default), 1 = staged through shared memory, 2 = thread-per-row.
 * Results are bit-identical; the variants exist for A/B measurements (tools/gemm_check.py --libs). */
int vqcpc_gemm_epilogue_kind(void);

/* profiling hook: device buffer of 12 uint64 that the CTA-pair GEMM kernel adds phase timers to (NULL = off); layout in
 * csrc/gemm_tc.cu */
int vqcpc_gemm_set_debug(void* buf);

/* profiling hook: timing experiments inside the GEMM epilogue (0 = off; non-zero values skip work and give INVALID results) */
int vqcpc_gemm_set_experiment(int e);

/* ---- full-sequence relative self-attention, L <= 1024 (student / teacher / auxiliary-decoder transformers:
 *      teachers/teacher_relative.py:33-47, auxiliary_decoders/auxiliary_decoder_relative.py:58-74; same layer math as the
 *      per-block kernels below: multihead_attention_custom.py:274-345 + subsampled_relative_attention.py:30-122) --------
 * qkv (P*L, 3*H*64) with the q third pre-scaled, e1 / e2 (H*L, 64); out (P*L, H*64) fp32 and / or bf16 split;
 * probs (P, H, L, L) or NULL.  Backward: d_qkv (and / or its split) overwritten, q third = q_scale * dq~; d_e1 / d_e2
 * ACCUMULATED; ds_scratch = (P, H, L, L) floats of caller-owned scratch. */
int vqcpc_attn_long_fwd(const float* qkv, const float* e1, const float* e2, float* out, void* out_hi, void* out_lo,
                        float* probs, int64_t P, int L, int H, int head_dim, void* stream);
int vqcpc_attn_long_bwd(const float* qkv, const float* e1, const float* e2, const float* probs, const float* d_out,
                        float* ds_scratch, float* d_qkv, void* d_qkv_hi, void* d_qkv_lo, float* d_e1, float* d_e2, int64_t P,
                        int L, int H, int head_dim, float q_scale, void* stream);

/* vqcpc_attn_long_fwd with the decoder's attention masks (decoders/decoder.py:295-306): mask_mode 0 none, 1 causal (keys <= query),
 * 2 anticausal (keys >= query); key tiles without a visible key are skipped.  fp32 output only; e1 == e2 == NULL: no relative bias
 * (transformer_type 'absolute'). */
int vqcpc_attn_long_fwd_masked(const float* qkv, const float* e1, const float* e2, float* out, float* probs, int64_t P, int L, int H,
                               int head_dim, int mask_mode, void* stream);

/* ---- general relative attention, forward only (decoder inference: decoders/decoder.py:431-543, :552-722;
 *      transformer_custom.py:355-384 decoder layer; multihead_attention_custom.py:168-353 incl. the encoder-decoder branch;
 *      subsampled_relative_attention.py:30-122 with seq_len_tgt != seq_len_src) -------------------------------------------
 * q (P*Lq, ldq) head h at columns [64h, 64h+64) of the pointer, pre-scaled; k, v (P*Lk, ldkv) likewise; e1 / e2 (H*Lk, 64);
 * map1 / map2 (Lq, Lk) int32: >= 0 -> bias term q~[m / Lk] . e[h, m % Lk], -1 -> 0, -2 -> -100 (the reference's pad / view
 * skewing replayed on flat indices, VQCPCB/transformer/relative_maps.py); mask (Lq, Lk) additive (-inf = hidden) or NULL;
 * out (P*Lq, H*64); probs (P, H, Lq, Lk) or NULL. */
int vqcpc_attn_map_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* e1, const float* e2,
                       const int* map1, const int* map2, const float* mask, float* out, float* probs, int64_t P, int Lq, int Lk,
                       int H, int head_dim, void* stream);

/* ---- incremental (key / value cached) generation step (decoders/decoder.py:607-644: the reference recomputes the whole
 *      teacher-forced forward per token and samples on the host; csrc/decode.cu explains why the cache is exact) -------------
 * Every entry point reads the current target position from DEVICE memory (`pos`, one int32) so that the launch sequence of
 * one event is position independent and can be replayed as a CUDA graph. */
/* feat (B, E + 2 pe) = [ embedding of token pos-1 | channel embedding | event embedding ] of position pos-1 (decoder.py:442-457
 * shifted right by one, :476-482); zeros at pos == 0.  tables: the per-channel embedding tables concatenated by rows,
 * table_off (C) int32 device = first row of each channel. */
int vqcpc_decode_features(const int64_t* tokens, const int* pos, const float* tables, const int* table_off, const float* chan,
                          const float* evt, float* feat, int B, int T, int C, int total_upscaling, int E, int pe, int pos_mode,
                          void* stream);
/* pos_mode 1 (transformer_type 'absolute', decoder.py:458-466): feat (B, E + pe) = [ embedding | chan[position] ] with chan the
 * (T, pe) table of absolute positional embeddings; evt is not read. */
/* nn.Linear on M <= 32 rows in exact fp32: out = [row0 if *pos == 0 else] relu?( (x W^T + bias) * scale on the first
 * scale_cols columns ) + residual.  W (N, K) row-major; bias / residual / row0 may be NULL. */
int vqcpc_linear_small(const float* x, int ldx, const float* W, const float* bias, const float* residual, int ldr, float* out,
                       int ldo, int M, int N, int K, int relu, float scale, int scale_cols, const float* row0, const int* pos,
                       void* stream);
/* the same with nn.LayerNorm(K) applied to the input rows while they are staged (post-LN blocks, transformer_custom.py:283-289:
 * x holds the pre-norm sum); ln_out (M, K) != NULL additionally receives the normalised rows (residual of a later linear). */
int vqcpc_ln_linear_small(const float* x, int ldx, const float* ln_gamma, const float* ln_beta, float ln_eps, float* ln_out,
                          const float* W, const float* bias, const float* residual, int ldr, float* out, int ldo, int M, int N, int K,
                          int relu, float scale, int scale_cols, void* stream);
/* one query row per (sequence, head) at position *pos (multihead_attention_custom.py:274-345 restricted to that row): q (P, ldq)
 * pre-scaled; k_new / v_new (P, ldn) appended to the cache at row *pos when non-NULL; kcache / vcache (P * Lmax rows, ldc);
 * e1 / e2 (H * Lk, 64) with map1 / map2 (Lq, Lk) int32 as in vqcpc_attn_map_fwd (row *pos is used); mask (Lq, Lk) additive or
 * NULL; causal != 0: keys 0 .. *pos, else all Lk keys; out (P, ldo). */
int vqcpc_attn_decode(const float* q, int ldq, const float* k_new, const float* v_new, int ldn, float* kcache, float* vcache,
                      int ldc, int Lmax, const float* e1, const float* e2, const int* map1, const int* map2, const float* mask,
                      int Lk, const int* pos, int causal, float* out, int ldo, int64_t P, int H, int head_dim, void* stream);
/* keys are split over a thread-block cluster of ns CTAs per (sequence, head): 0 = automatic (default), 1, 2, 4, 8. */
int vqcpc_attn_decode_set_split(int ns);
/* tokens[b, *pos] = one draw per row of logits (B, V; ld): logits / temperature, top-k and top-p filtering (utils.py:100-128),
 * softmax, then numpy.random.choice's inverse-CDF draw (decoder.py:636-644) with the caller's uniform u[*pos * u_stride + b]
 * (double, [0, 1)); greedy != 0: argmax. */
int vqcpc_sample_step(const float* logits, int ld, int V, float temperature, int top_k, float top_p, const double* u, int u_stride,
                      int64_t* tokens, int T, const int* pos, int greedy, int B, void* stream);
/* vqcpc_sample_step + vqcpc_decode_features of the NEXT position (from the token just drawn) + *pos += 1, in one launch;
 * done: one zero-initialised int32 (arrival counter of the B CTAs, left at zero). */
int vqcpc_decode_sample_embed(const float* logits, int ld, int V, float temperature, int top_k, float top_p, const double* u,
                              int u_stride, int64_t* tokens, int T, int* pos, int greedy, int B, const float* tables,
                              const int* table_off, const float* chan, const float* evt, float* feat, int C, int total_upscaling,
                              int E, int pe, int* done, int pos_mode, void* stream);
/* aligned ('diagonal') decoder layer, transformer_custom.py:389-479: its cross term is an MLP of the aligned memory event
 * (Linear, ELU, Linear - :428-432), i.e. a per-position constant add (B, T, d) during generation:
 * out (B, d) = LayerNorm(y) + add[:, *pos, :] (:447-449 norm1, :472 residual). */
int vqcpc_decode_ln_add(const float* y, const float* gamma, const float* beta, float eps, const float* add, int T, const int* pos,
                        float* out, int B, int d, void* stream);
int vqcpc_elu_fwd(const float* x, float* out, int64_t n, void* stream);
int vqcpc_decode_advance(int* pos, int delta, void* stream);
/* programmatic dependent launch between the kernels of the generation step (default on): each kernel requests its static
 * parameters (weight rows, biases, LayerNorm parameters) while its predecessor is still running. */
int vqcpc_decode_set_pdl(int on);

/* ---- CPC batch assembly from a chorale tensor dataset resident in HBM (datasets/chorale_dataset.py:542-595 data_loaders over
 *      the cached tensor_dataset; dataloaders/bach_cpc_dataloader.py:110-252) -------------------------------------------------
 * data (n_rows, V, ticks) int32 (elem_bytes 4: the reference's cache) or int64 (8); rows: sampled row ids (int64, device).
 * out (n, n_ev, V) int64 = data[rows[r], :, t0 : t0 + n_ev] with the voice dimension last (:227-233, :242-245). */
int vqcpc_cpc_window_gather(const void* data, int elem_bytes, int64_t n_rows, int V, int ticks, const int64_t* rows, int64_t n, int t0,
                            int n_ev, int64_t* out, void* stream);
/* _build_negatives_sameSeq (:162-184): out (B, KA + KB - 1, nk, evb, V): for every k < nk all KA blocks of part A (ticks from a0),
 * then the KB blocks of part B (ticks from b0) except block k (the reference loops k over num_blocks_right for both directions). */
int vqcpc_cpc_same_seq_negatives(const void* data, int elem_bytes, int64_t n_rows, int V, int ticks, const int64_t* rows, int64_t B,
                                 int a0, int KA, int b0, int KB, int nk, int evb, int64_t* out, void* stream);

/* ---- softmax cross-entropy rows (utils.py:24-49 categorical_crossentropy via nn.CrossEntropyLoss; utils.py:131-159
 *      distilled_categorical_crossentropy) --------------------------------------------------------------------------
 * logits (rows, C; ld).  Exactly one of target_idx (int64 class per row) / target_logits (rows, C; ldt: the teacher's
 * logits, target distribution = their softmax).  loss (rows) and / or d_logits (rows, C; ldd) = scale * dloss/dlogits. */
int vqcpc_softmax_ce(const float* logits, int ld, const int64_t* target_idx, const float* target_logits, int ldt, float* loss,
                     float* d_logits, int ldd, float scale, int64_t rows, int C, void* stream);

/* ---- per-block relative self-attention (multihead_attention_custom.py:274-345 +
 *      subsampled_relative_attention.py:30-122) ---------------------------------------------------------------------
 * qkv (P*L, 3*H*64), q third pre-scaled by hd^-0.5; e1, e2 (H*L, 64).  Only the query rows with token % keep == 0 are
 * computed (keep = 1: all; keep = downscale factor in a stack's last layer = exact dead-row elimination) and stored
 * compacted: out fp32 and/or bf16 split (out_hi, out_lo), (P*L/keep, H*64); probs (P, H, L, L) or NULL. */
int vqcpc_attn_fwd(const float* qkv, const float* e1, const float* e2, float* out, void* out_hi, void* out_lo, float* probs,
                   int64_t P, int L, int H, int head_dim, int keep, void* stream);
/* d_out (P*L/keep, H*64); d_qkv fp32 and/or its bf16 split overwritten (q third = q_scale * dq~); d_e1 / d_e2 ACCUMULATED. */
int vqcpc_attn_bwd(const float* qkv, const float* e1, const float* e2, const float* probs, const float* d_out,
                   float* d_qkv, void* d_qkv_hi, void* d_qkv_lo, float* d_e1, float* d_e2, int64_t P, int L, int H,
                   int head_dim, float q_scale, int keep, void* stream);

/* Attention kernel choice (process-wide): 0 = tensor cores (mma.sync 3xTF32, ~2^-20 per product; default), 1 = CUDA cores
 * (plain fp32 FMA; what the Python side selects in the exact 'fp32' GEMM mode). */
int vqcpc_attn_set_fwd_algo(int algo);
int vqcpc_attn_set_bwd_algo(int algo);
/* Same with dropout on the attention weights (F.dropout at multihead_attention_custom.py:339): drop_keep (P, H, L, L) uint8
 * keep mask or NULL, drop_scale = 1 / (1 - p).  `probs` always holds the softmax output; the mask acts on the P.V product
 * (forward) and on dA / dV (backward). */
int vqcpc_attn_fwd_ex(const float* qkv, const float* e1, const float* e2, float* out, void* out_hi, void* out_lo, float* probs,
                      int64_t P, int L, int H, int head_dim, int keep, const unsigned char* drop_keep, float drop_scale,
                      void* stream);
int vqcpc_attn_bwd_ex(const float* qkv, const float* e1, const float* e2, const float* probs, const float* d_out,
                      float* d_qkv, void* d_qkv_hi, void* d_qkv_lo, float* d_e1, float* d_e2, int64_t P, int L, int H,
                      int head_dim, float q_scale, int keep, const unsigned char* drop_keep, float drop_scale, void* stream);

/* ---- vector quantizer (quantizer/vector_quantizer.py:85-159) --------------------------------------------------
 * x (n, D); books: host array of ncb device pointers, each (Kc, D/ncb); idx (n, ncb) int64. */
int vqcpc_vq_search(const float* x, int64_t n, int D, const float* const* books /*host*/, int ncb, int Kc, int64_t* idx,
                    void* stream);
/* q = E[idx]; loss (n) = q_latent + beta * e_latent (:72-83); z_sg = x + (q - x) (:148); counts (ncb*Kc) int32
 * usage histogram ACCUMULATED (may be NULL). */
int vqcpc_vq_gather(const float* x, int64_t n, int D, const float* const* books /*host*/, int ncb, int Kc,
                    const int64_t* idx, float beta, int squared, float* z_sg, float* loss, int* counts, void* stream);
/* d_x overwritten = g_zsg + g_loss * dloss/dx; d_books[cb] ACCUMULATED (host array, entries may be NULL). */
int vqcpc_vq_bwd(const float* x, int64_t n, int D, const float* const* books /*host*/, float* const* d_books /*host*/,
                 int ncb, int Kc, const int64_t* idx, float beta, int squared, const float* g_zsg, const float* g_loss,
                 float* d_x, void* stream);

/* Fused single-codebook bottleneck with a dropout-free MLP upscaler folded into a table T = upscaler(E) (Kc, zd):
 * one pass: exact search -> idx (n) int64, loss (n), optional z_sg (n, D), y = T[idx] (n, zd) (T and y may both be NULL),
 * counts (Kc) ACCUMULATED or NULL.  Replaces vector_quantizer.py:104-157 + mlp_upscaler.py:21-35 on every vector. */
int vqcpc_vq_fused_fwd(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta,
                       int squared, int64_t* idx, float* loss, float* z_sg, float* y, int* counts, void* stream);
/* Same, with the search algorithm selectable: algo 0 = automatic, 1 = direct-form scan on the CUDA cores (the
 * reference's arithmetic for every (vector, code) pair), 2 = tcgen05 tensor-core ranking of z.e - |e|^2/2 followed by
 * exact re-evaluation of the codes within the error margin (D in {3, 4}, Kc <= 2048).  Indices are bit-identical between
 * the two.  dbg_scores (n, ceil(Kc/256)*256) and stats[4098] (exact group re-evaluations, full-scan fallbacks, then 8
 * clock64 stamps per CTA) are test / profiling hooks of algo 2 and are NULL in production. */
int vqcpc_vq_fused_fwd_ex(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta,
                          int squared, int64_t* idx, float* loss, float* z_sg, float* y, int* counts, int algo,
                          float* dbg_scores, int* stats, void* stream);
/* backward: d_x = J[idx]^T dy + g_loss * dloss/dx (J (Kc, zd, D) = upscaler Jacobian at each code; straight-through);
 * d_T[idx] += dy and d_E[idx] += g_loss * dloss/dq are ACCUMULATED. */
int vqcpc_vq_fused_bwd(const float* x, int64_t n, int D, const float* E, const float* J, int Kc, int zd, const int64_t* idx,
                       float beta, int squared, const float* dy, const float* g_loss, float* d_x, float* d_T, float* d_E,
                       void* stream);

/* ---- fused InfoNCE (vqcpc_helper.py:5-29, :79-98; vqcpc_encoder_trainer.py:240-272) ---------------------------
 * c (B, cd); W (zd, cd, KR); y_pos (B, KR, zd); y_neg (B, N, KR, zd).  Outputs: u (B, KR, zd), pos (B, KR),
 * neg (B, KR, N), lse (B, KR), nll = lse - pos (B, KR), correct (B, KR) int32 = [pos > max_n neg]. */
int vqcpc_infonce_fwd(const float* c, const float* W, const float* y_pos, const float* y_neg, int B, int N, int KR, int zd,
                      int cd, float* u, float* pos, float* neg, float* lse, float* nll, int* correct, void* stream);
/* tuning / test hook: forward kernel choice, 0 = automatic, 1 = warp per (b, k), 2 = streaming cp.async.bulk ring. */
int vqcpc_infonce_set_algo(int algo);
/* gradient of sum(nll) scaled by scale * g_dev[0] (g_dev device scalar or NULL); d_c, d_W ACCUMULATED. */
int vqcpc_infonce_bwd(const float* c, const float* W, const float* y_pos, const float* y_neg, const float* u,
                      const float* pos, const float* neg, const float* lse, const float* g_dev, float scale, int B, int N,
                      int KR, int zd, int cd, float* d_c, float* d_W, float* d_ypos, float* d_yneg, void* stream);

/* ---- index-gather form of the quantizer -> upscaler -> InfoNCE path (csrc/idxpath.cu) ------------------------------
 * Table mode (one codebook, no active upscaler dropout): every upscaled vector is a row of T = upscaler(E) (Kc, zd), so
 * the scores f_k = u_k . y of vqcpc_helper.py:79-98 and the loss of :5-29 are evaluated from the code indices, and the
 * backward never materialises d y.  Supported: zd == cd == 32, 1 <= N <= 15, 1 <= KR <= 16, Kc <= 698 (T and d T in shared
 * memory), D in {3, 4}; vqcpc_idxpath_supported() returns 1 for such shapes (callers keep the per-vector entry points
 * above for everything else). */
int vqcpc_idxpath_supported(int N, int KR, int zd, int cd, int Kc, int D);
/* debugging aid: clock64 phase timers of block 0 of the last vqcpc_infonce_idx_fwd launch (16 values, host buffer). */
int vqcpc_idx_debug_read(long long* out16 /*host*/);
/* vqcpc_vq_fused_fwd with the upscaled rows y = T[idx] produced only for the rows [y_row0, y_row0 + y_rows) (the left
 * blocks, which the context GRU reads; y has y_rows rows).  Replaces vector_quantizer.py:104-157 + mlp_upscaler.py:21-35. */
int vqcpc_vq_fused_fwd_rows(const float* x, int64_t n, int D, const float* E, const float* T, int Kc, int zd, float beta,
                            int squared, int64_t* idx, float* loss, float* z_sg, float* y, int64_t y_row0, int64_t y_rows,
                            int* counts, void* stream);
/* c (B, cd); W (zd, cd, KR); T (Kc, zd); idx_pos (B, KR), idx_neg (B, N, KR) int64.  Outputs as vqcpc_infonce_fwd:
 * u (B, KR, zd), pos (B, KR), neg (B, KR, N), lse, nll = lse - pos (B, KR), correct (B, KR) int32.
 * vqcpc_helper.py:79-98, :5-29; vqcpc_encoder_trainer.py:240-272. */
int vqcpc_infonce_idx_fwd(const float* c, const float* W, const float* T, const int64_t* idx_pos, const int64_t* idx_neg,
                          int B, int N, int KR, int zd, int cd, int Kc, float* u, float* pos, float* neg, float* lse,
                          float* nll, int* correct, void* stream);
/* per-CTA partial tables vqcpc_infonce_idx_bwd writes, and the floats of scratch it needs in total
 * (parts * (Kc * zd + zd * cd * KR) for the partial d_T | d_W tables + B * KR * cd per-row d_c contributions). */
int vqcpc_infonce_idx_parts(int B, int KR);
int64_t vqcpc_infonce_idx_scratch(int B, int KR, int Kc);
/* gradient of sum(nll) scaled by scale * g_dev[0] (g_dev device scalar or NULL): d_c (B, cd) overwritten; coef_pos (B, KR)
 * and coef_neg (B, N, KR) = d loss / d score of every candidate, overwritten (vqcpc_vq_idx_bwd turns them into d z);
 * d_W (zd, cd, KR) and d_T (Kc, zd) ACCUMULATED (per-CTA partial tables in the scratch `part`, summed in a fixed order by a
 * second kernel).  d_T = C^T U runs on the tensor cores: tcgen05.mma with the accumulators in TMEM (bf16 hi/lo operand
 * split, fp32 accumulation); VQCPC_NCE_BWD_ALGO=1 selects the CUDA-core owner-computes kernel instead. */
int vqcpc_infonce_idx_bwd(const float* c, const float* W, const float* T, const int64_t* idx_pos, const int64_t* idx_neg,
                          const float* u, const float* pos, const float* neg, const float* lse, const float* g_dev,
                          float scale, int B, int N, int KR, int zd, int cd, int Kc, float* d_c, float* coef_pos,
                          float* coef_neg, float* part, float* d_W, float* d_T, void* stream);
/* test / tuning hooks: kernel choice of the two backward entry points.  infonce_idx: 0 = tcgen05 d T (default), 1 = CUDA-core
 * owner-computes kernel;  vq_idx: 0 = streaming d z + counting-sort d E kernels (default), 1 = single kernel;  -1 = follow
 * the environment (VQCPC_NCE_BWD_ALGO / VQCPC_VQ_BWD_ALGO). */
int vqcpc_infonce_idx_set_bwd_algo(int algo);
int vqcpc_vq_idx_set_bwd_algo(int algo);
/* out[idx[i]] += rows[i] (rows (n, zd), out (Kc, zd) ACCUMULATED) without atomics, deterministic: the gradient of the
 * table gather y = T[idx] of the left blocks. */
int vqcpc_table_scatter_add(const int64_t* idx, int64_t n, const float* rows, int zd, int Kc, float* out, void* stream);
/* backward of the quantizer in index form.  d_x (n, D) overwritten = coef_p * J[idx_p]^T v_p + g_loss_p * dloss/dx, where
 * the upscaled-side gradient v_p of the vectors of segment s (rows [seg_row0[s], +seg_rows[s]), at most 4 segments, other
 * vectors get none) is  kind 1: the dense row seg_a[s][(p - row0) * zd ..] (coef = 1)
 *                       kind 2: coef = seg_a[s][p - row0],  v = seg_b[s][((p - row0) / per_b * KR + (p - row0) % KR) * zd ..]
 * (per_b = N * KR for the negatives laid out (B, N, KR), KR for the right blocks; seg_b = the u of vqcpc_infonce_idx_fwd).
 * d_E (Kc, D) ACCUMULATED (may be NULL); d_T (Kc, zd) ACCUMULATED with the kind-1 rows (d_T[idx_p] += row: the gradient
 * of the table gather of the left blocks) or NULL.  J (Kc, zd, D), g_loss (n) or NULL.  seg_* are HOST arrays. */
int vqcpc_vq_idx_bwd(const float* x, int64_t n, int D, const float* E, const float* J, int Kc, int zd, const int64_t* idx,
                     float beta, int squared, const float* g_loss, int nseg, const int64_t* seg_row0 /*host*/,
                     const int64_t* seg_rows /*host*/, const int* seg_kind /*host*/, const float* const* seg_a /*host*/,
                     const float* const* seg_b /*host*/, const int* seg_per_b /*host*/, int KR, float* d_x, float* d_E,
                     float* d_T, void* stream);

/* ---- GRU (torch.nn.GRU semantics; CModule vqcpc_helper.py:54-76, LstmDownscaler lstm_downscaler.py:31-48) ---------------
 * The products GI = X W_ih^T + b_ih (all steps at once) and GH = h_{t-1} W_hh^T + b_hh (per step) are vqcpc_gemm_* calls;
 * these are the point-wise stages of step t, gate order (r, z, n), batch-first layouts: gi (B*T, 3H) rows b*T + t,
 * gh (B, 3H), h_prev (B, H) with row stride hprev_ld (NULL = zero state), out (B, T, H) written at [:, t],
 * gates (B*T, 3H) and ghn (B*T, H) saved for backward (both NULL for inference); out_hi / out_lo: optional bf16 split
 * of h_t (same layout as out), which the next step's tensor-core GEMM reads directly. */
int vqcpc_gru_gate_fwd(const float* gi, const float* gh, const float* h_prev, int64_t hprev_ld, float* out, void* out_hi,
                       void* out_lo, float* gates, float* ghn, int B, int T, int H, int t, void* stream);
/* dh (B, H): recurrent part of the gradient wrt h_t on entry (the upstream d_out[:, t], (B, T, H) or NULL, is added here),
 * overwritten with (dh + d_out_t) * z (the caller adds dGH_t W_hh to obtain the gradient wrt h_{t-1}); d_gi / d_gh
 * (B*T, 3H) receive the gradients of GI_t / GH_t at rows b*T + t; d_gh may be given as fp32 and / or as its bf16 split. */
int vqcpc_gru_gate_bwd(const float* gates, const float* ghn, const float* h_prev, int64_t hprev_ld, float* dh,
                       const float* d_out, float* d_gi, float* d_gh, void* d_gh_hi, void* d_gh_lo, int B, int T, int H,
                       int t, void* stream);

/* ---- optimizer tail (clip_grad_norm_(…, 5) + Adam: vqcpc_encoder_trainer.py:92, :313-314) ----------------------
 * out[0] = ||grad_scale * g||_2^2 (device scalar, deterministic). */
int vqcpc_grad_sumsq(const float* g, int64_t n, float grad_scale, float* out, void* stream);
/* p, m, v updated in place from grad_scale * g clipped by min(1, max_norm / (sqrt(gnorm_sq[0]) + 1e-6)); step >= 1. */
int vqcpc_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* gnorm_sq, float max_norm,
                    float grad_scale, float lr, float beta1, float beta2, float eps, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQCPC_B200_H */
