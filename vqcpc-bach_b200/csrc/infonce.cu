// Fused InfoNCE scoring: Fks bilinear scores + log-sum-exp + accuracy bit, forward and backward.
//
// Reference: VQCPCB/vqcpc_helper.py:79-98 (FksModule: u[b,k,:] = W[:,:,k] c[b];  f = u . z) called for the positive
// (vqcpc_encoder_trainer.py:240) and for the N negatives (:245-263); nce_loss vqcpc_helper.py:5-29;
// score matrix pos > max(neg) vqcpc_encoder_trainer.py:269.
//
// Negatives are per sample (negative_samples[b] is only scored against c[b]), so scoring is a batched GEMV:
// per (b, k) one u vector against N+1 vectors of zd floats -> 0.5 FLOP/B, HBM-bound.  One warp per (b, k) with
// lane == z index (zd <= 32 per pass); W[:, :, k] is staged once per CTA (blockIdx.y = k).  The positive and every
// negative go through the SAME dot-product code so equal vectors give bit-equal scores (ties are not "correct").
//
// Layouts: c (B, cd); W (zd, cd, KR) [reference parameter layout]; y_pos (B, KR, zd); y_neg (B, N, KR, zd)
// (= the encoder output on negative_samples.view(B*N*KR, T, V), no permute needed);
// outputs u (B, KR, zd), pos (B, KR), neg (B, KR, N), lse (B, KR), correct (B, KR) int32.
#include "common.cuh"

namespace {

constexpr int MAXZ = 128;   // zd <= 128 (4 floats per lane)
int g_nce_algo = 0;         // 0 = automatic, 1 = warp-per-(b,k) kernel, 2 = streaming kernel (test / tuning hook)

template <int ZPL>   // z elements per lane: zd <= 32*ZPL
__global__ void __launch_bounds__(256)
infonce_fwd_kernel(const float* __restrict__ c, const float* __restrict__ W, const float* __restrict__ ypos,
                   const float* __restrict__ yneg, int B, int N, int KR, int zd, int cd, float* __restrict__ u_out,
                   float* __restrict__ pos_out, float* __restrict__ neg_out, float* __restrict__ lse_out,
                   float* __restrict__ nll_out, int* __restrict__ correct_out) {
  extern __shared__ float Wk[];                  // [zd][cd + 1]
  const int k = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cs = cd + 1;
  for (int i = threadIdx.x; i < zd * cd; i += blockDim.x) {
    int z = i / cd, cc = i % cd;
    Wk[z * cs + cc] = W[((size_t)z * cd + cc) * KR + k];
  }
  __syncthreads();
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int b = blockIdx.x * (blockDim.x >> 5) + warp; b < B; b += warps_total) {
    float u[ZPL];
#pragma unroll
    for (int t = 0; t < ZPL; ++t) {
      int z = lane + 32 * t;
      float a = 0.f;
      if (z < zd)
        for (int cc = 0; cc < cd; ++cc) a = fmaf(__ldg(c + (size_t)b * cd + cc), Wk[z * cs + cc], a);
      u[t] = a;
      if (z < zd) u_out[((size_t)b * KR + k) * zd + z] = a;
    }
    auto score = [&](const float* __restrict__ y) {
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < ZPL; ++t) {
        int z = lane + 32 * t;
        if (z < zd) a = fmaf(u[t], y[z], a);
      }
      return warp_sum(a);
    };
    float p = score(ypos + ((size_t)b * KR + k) * zd);
    // online log-sum-exp over [neg_0 .. neg_{N-1}, pos]; every lane carries the same running (max, sum).
    // Negatives are fetched 8 at a time (independent 128-byte loads in flight) before the dependent reduce chain.
    float mx = p, se = 1.f, nmax = -INFINITY;
    for (int n0 = 0; n0 < N; n0 += 8) {
      float yv[8][ZPL];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int t = 0; t < ZPL; ++t) {
          int z = lane + 32 * t;
          yv[j][t] = (n0 + j < N && z < zd) ? yneg[(((size_t)b * N + n0 + j) * KR + k) * zd + z] : 0.f;
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n0 + j < N) {
          float a = 0.f;
#pragma unroll
          for (int t = 0; t < ZPL; ++t) a = fmaf(u[t], yv[j][t], a);     // same fma order as score()
          float s = warp_sum(a);
          nmax = fmaxf(nmax, s);
          if (lane == 0) neg_out[((size_t)b * KR + k) * N + n0 + j] = s;
          float m2 = fmaxf(mx, s);
          se = se * expf(mx - m2) + expf(s - m2);
          mx = m2;
        }
      }
    }
    if (lane == 0) {
      pos_out[(size_t)b * KR + k] = p;
      const float l = mx + logf(se);
      lse_out[(size_t)b * KR + k] = l;
      nll_out[(size_t)b * KR + k] = l - p;         // -(pos - lse): summed by the caller (vqcpc_reduce_sum)
      correct_out[(size_t)b * KR + k] = (p > nmax) ? 1 : 0;
    }
  }
}

// backward.  scale = dL/dnce / B_total  (nce = -(1/B) sum_{b,k} (pos - lse)).
//   pi = softmax([neg, pos]); dpos = -(1 - pi_pos) * scale ; dneg_n = pi_n * scale
//   du = dpos * y_pos + sum_n dneg_n * y_neg[n] ; dy_pos = dpos * u ; dy_neg[n] = dneg_n * u
//   dc[b] += sum_k W[:, :, k]^T du[b, k] ;  dW[:, :, k] += sum_b du[b, k] (x) c[b]
template <int ZPL>
__global__ void __launch_bounds__(256)
infonce_bwd_kernel(const float* __restrict__ c, const float* __restrict__ W, const float* __restrict__ ypos,
                   const float* __restrict__ yneg, const float* __restrict__ u_in, const float* __restrict__ pos_in,
                   const float* __restrict__ neg_in, const float* __restrict__ lse_in, const float* __restrict__ g_dev,
                   float scale_host, int B, int N, int KR,
                   int zd, int cd, float* __restrict__ dc, float* __restrict__ dW, float* __restrict__ dypos,
                   float* __restrict__ dyneg) {
  extern __shared__ float sm[];
  const int cs = cd + 1;
  float* Wk = sm;                     // [zd][cs]
  float* dWk = sm + zd * cs;          // [zd][cs] CTA-private accumulator
  const int k = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = scale_host * (g_dev ? __ldg(g_dev) : 1.f);
  for (int i = threadIdx.x; i < zd * cd; i += blockDim.x) {
    int z = i / cd, cc = i % cd;
    Wk[z * cs + cc] = W[((size_t)z * cd + cc) * KR + k];
    dWk[z * cs + cc] = 0.f;
  }
  __syncthreads();
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int b = blockIdx.x * (blockDim.x >> 5) + warp; b < B; b += warps_total) {
    const size_t bk = (size_t)b * KR + k;
    const float lse = lse_in[bk];
    const float p = pos_in[bk];
    const float dpos = -(1.f - expf(p - lse)) * scale;
    float u[ZPL], du[ZPL];
#pragma unroll
    for (int t = 0; t < ZPL; ++t) {
      int z = lane + 32 * t;
      u[t] = (z < zd) ? u_in[bk * zd + z] : 0.f;
      du[t] = (z < zd) ? dpos * ypos[bk * zd + z] : 0.f;
      if (z < zd) dypos[bk * zd + z] = dpos * u[t];
    }
    for (int n0 = 0; n0 < N; n0 += 8) {
      float yv[8][ZPL], dn[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dn[j] = (n0 + j < N) ? expf(neg_in[bk * N + n0 + j] - lse) * scale : 0.f;
#pragma unroll
        for (int t = 0; t < ZPL; ++t) {
          int z = lane + 32 * t;
          yv[j][t] = (n0 + j < N && z < zd) ? yneg[(((size_t)b * N + n0 + j) * KR + k) * zd + z] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n0 + j < N) {
          size_t off = (((size_t)b * N + n0 + j) * KR + k) * zd;
#pragma unroll
          for (int t = 0; t < ZPL; ++t) {
            int z = lane + 32 * t;
            if (z < zd) {
              du[t] = fmaf(dn[j], yv[j][t], du[t]);
              dyneg[off + z] = dn[j] * u[t];
            }
          }
        }
      }
    }
    // dW_k[z][cc] += du[z] * c[b][cc]   (CTA-private smem, lane owns rows z)
    // dc[b][cc]   += sum_z Wk[z][cc] * du[z]
    for (int cc = 0; cc < cd; ++cc) {
      float cv = __ldg(c + (size_t)b * cd + cc);
      float part = 0.f;
#pragma unroll
      for (int t = 0; t < ZPL; ++t) {
        int z = lane + 32 * t;
        if (z < zd) {
          atomicAdd(&dWk[z * cs + cc], du[t] * cv);
          part = fmaf(Wk[z * cs + cc], du[t], part);
        }
      }
      part = warp_sum(part);
      if (lane == 0) atomicAdd(dc + (size_t)b * cd + cc, part);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < zd * cd; i += blockDim.x) {
    int z = i / cd, cc = i % cd;
    atomicAdd(dW + ((size_t)z * cd + cc) * KR + k, dWk[z * cs + cc]);
  }
}


// ------------------------------------------------------------------------------------------------------------
// Streaming forward: the scoring is a batched GEMV at ~0.5 FLOP/B, i.e. purely HBM-bound (16.6 KB per sample at
// KR = 8, N = 15, zd = 32).  Persistent CTAs pull GROUPS of 8/KR samples through rings of shared-memory stages with
// cp.async.bulk (three 1-D TMA copies per group: the contiguous negatives, positives and context vectors;
// completion on an mbarrier), so several stages (>= 100 KB per SM) are always in flight and no thread ever waits on a
// global load:
//   U phase: u[k, z] = sum_c W[z, c, k] c[c]           (thread per (k, z); its W row stays in registers)
//   S phase: score[k, n] = sum_z u[k, z] y[n, k, z]    (thread per (k, n); n = N is the positive), then the
//            log-sum-exp / accuracy bit of a (sample, k) with shuffles among its 16 lanes
// Every row block starts at bank 0, so lane l walks the 16-byte chunks of its row in the rotated order (j + l) mod 8;
// the chunk partial sums are combined by the rotation-invariant tree ((p0+p4)+(p2+p6))+((p1+p5)+(p3+p7)) (generalised to
// 8 chunks), so every lane computes the SAME floating-point expression of its row whatever its rotation: identical
// vectors give bit-equal scores and a negative equal to the positive is never counted "correct".
__device__ __forceinline__ uint32_t nce_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void nce_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void nce_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nce_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void nce_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if (it > (1u << 26)) { printf("vqcpc infonce: mbarrier wait timed out (block %d)\n", blockIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void nce_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// The streaming kernel is specialised for the shapes every shipped configuration uses - zd = cd = 32 (upscaler output,
// c_net output), N = 15 negatives - and KR in {2, 4, 8}.  A CTA is NCE_TEAMS independent teams of 128 threads; each team
// streams its own sample groups (8 / KR samples = 8 (sample, k) rows x 16 scores = one score per thread, 256 u values =
// two per thread) through its own ring and synchronises on its own named barrier, so the load -> fma -> shuffle
// latency chains of different teams overlap.  Other shapes use infonce_fwd_kernel.
constexpr int NCE_TEAMS = 4, NCE_TT = 128, NCE_THREADS = NCE_TEAMS * NCE_TT;
constexpr int NCE_ZD = 32, NCE_CD = 32, NCE_GRP = 16, NCE_N = NCE_GRP - 1, NCE_NC = NCE_ZD / 4;

struct NceParams {
  const float *c, *W, *ypos, *yneg;
  int B, nst;
  float *u, *pos, *neg, *lse, *nll;
  int* correct;
};

// dot product of two rows of NC 16-byte chunks, chunk order rotated by `rot`, rotation-invariant combination (see above)
template <int NC>
__device__ __forceinline__ float rot_dot(const float* __restrict__ uv, const float* __restrict__ yv, int rot) {
  float q[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int ch = (j + rot) & (NC - 1);
    const float4 y4 = *reinterpret_cast<const float4*>(yv + ch * 4);
    const float4 u4 = *reinterpret_cast<const float4*>(uv + ch * 4);
    q[j] = fmaf(u4.w, y4.w, fmaf(u4.z, y4.z, fmaf(u4.y, y4.y, u4.x * y4.x)));
  }
  // pairs (j, j + NC/2), then (j, j + NC/4), ...: each level is mapped onto itself by a cyclic shift of the indices
#pragma unroll
  for (int h = NC / 2; h >= 1; h >>= 1)
#pragma unroll
    for (int j = 0; j < h; ++j) q[j] = q[j] + q[j + h];
  return q[0];
}

template <int KR>
struct NceShape {
  static constexpr int SPS = 8 / KR;                          // samples per stage
  static constexpr int ROWB = KR * NCE_ZD;                    // floats per (sample, n) row block
  static constexpr int NEG_F = NCE_N * ROWB;                  // floats of negatives per sample
  static constexpr int STAGE_F = SPS * (NEG_F + ROWB + NCE_CD);   // [SPS][N][KR][zd] | [SPS][KR][zd] | [SPS][cd]
  static constexpr int TEAM_US = 8 * NCE_ZD;                  // u of the 8 (sample, k) rows of the current stage
  static size_t smem(int nst) { return (size_t)NCE_TEAMS * ((size_t)nst * STAGE_F + TEAM_US) * 4 + NCE_TEAMS * 2 * (size_t)nst * 8 + 16; }
};

template <int KR>
__global__ void __launch_bounds__(NCE_THREADS, 1)
infonce_stream_fwd_kernel(NceParams p) {
  using S = NceShape<KR>;
  extern __shared__ __align__(16) uint8_t nce_smem[];
  const int team = threadIdx.x / NCE_TT, tt = threadIdx.x % NCE_TT, lane = threadIdx.x & 31;
  float* stg = reinterpret_cast<float*>(nce_smem) + (size_t)team * ((size_t)p.nst * S::STAGE_F + S::TEAM_US);   // [nst][STAGE_F]
  float* us = stg + (size_t)p.nst * S::STAGE_F;                                                                 // [8][zd]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(nce_smem) +
                                               (size_t)NCE_TEAMS * ((size_t)p.nst * S::STAGE_F + S::TEAM_US)) + team * 2 * p.nst;
  const uint32_t bar0 = nce_smem_u32(bars);            // full[nst], empty[nst] of this team

  if (tt == 0) {
    for (int i = 0; i < 2 * p.nst; ++i) nce_mbar_init(bar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // sample groups are dealt round-robin to (CTA, team) streams
  const int ngroups = (p.B + S::SPS - 1) / S::SPS;
  const int stream = blockIdx.x * NCE_TEAMS + team, nstreams = gridDim.x * NCE_TEAMS;
  const int my = stream < ngroups ? (ngroups - stream + nstreams - 1) / nstreams : 0;
  // thread 0 of the team issues the three copies of its group i into ring slot i % nst
  auto issue = [&](int i) {
    if (i >= my) return;
    const int slot = i % p.nst;
    if (i >= p.nst) nce_mbar_wait(bar0 + 8u * (p.nst + slot), ((i / p.nst) & 1) ^ 1);
    const int b0 = (stream + i * nstreams) * S::SPS;
    const uint32_t ns = (uint32_t)min(S::SPS, p.B - b0);
    const uint32_t fb = bar0 + 8u * slot;
    float* d = stg + (size_t)slot * S::STAGE_F;
    nce_mbar_expect_tx(fb, ns * (uint32_t)(S::NEG_F + S::ROWB + NCE_CD) * 4u);
    nce_bulk_g2s(nce_smem_u32(d), p.yneg + (size_t)b0 * S::NEG_F, ns * (uint32_t)S::NEG_F * 4u, fb);
    nce_bulk_g2s(nce_smem_u32(d + S::SPS * S::NEG_F), p.ypos + (size_t)b0 * S::ROWB, ns * (uint32_t)S::ROWB * 4u, fb);
    nce_bulk_g2s(nce_smem_u32(d + S::SPS * (S::NEG_F + S::ROWB)), p.c + (size_t)b0 * NCE_CD, ns * (uint32_t)NCE_CD * 4u, fb);
  };
  // roles: U phase - rows ru and ru + 4 of the 8 (sample, k) rows, column zu (the W rows of the thread never change: they
  // live in registers for the whole kernel);  S phase - row rs, score nn (15 = positive)
  const int ru = tt >> 5, zu = tt & 31;
  const int rs = tt >> 4, nn = tt & 15;
  constexpr bool TWO_K = KR == 8;                      // rows ru and ru + 4 belong to different k only when KR = 8
  float w0[NCE_CD], w1[TWO_K ? NCE_CD : 1];
  {
    // W (zd, cd, KR) is read once per CTA with coalesced loads - issued BEFORE the first bulk copies, so they do not
    // queue behind megabytes of streaming traffic - into the ring slots the prologue does not fill (slot nst - 1 of
    // every team holds a quarter of it); then each thread picks its strided rows out of shared memory
    constexpr int W_F = NCE_ZD * NCE_CD * KR, QUART = W_F / NCE_TEAMS, PER = W_F / 4 / NCE_THREADS;
    static_assert(QUART + 8 + 24 <= S::STAGE_F && NCE_ZD == 8 * NCE_TEAMS, "W quarter (8 z rows) must fit in a ring slot");
    static_assert(W_F % (4 * NCE_THREADS) == 0, "W must split evenly over the CTA");
    float4 wv[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) wv[j] = __ldg(reinterpret_cast<const float4*>(p.W) + threadIdx.x + j * NCE_THREADS);
    __syncwarp();                                      // (lane 0 of each team initialised barriers: reconverge first - synccheck)
    __syncthreads();                                   // barriers initialised
    if (tt == 0)
      for (int i = 0; i < p.nst - 1; ++i) issue(i);
    // quarter q of W (z rows 8q .. 8q+7, each cd * KR floats + one word of padding so that the 32 lanes of a warp - 32
    // different z - read 32 different banks; quarters are skewed by 8 words for the same reason) sits in slot nst - 1 of team q
    constexpr int ROW = NCE_CD * KR, ROWP = ROW + 1;
    float* base = reinterpret_cast<float*>(nce_smem) + (size_t)(p.nst - 1) * S::STAGE_F;
    const int team_f = p.nst * S::STAGE_F + S::TEAM_US;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = (threadIdx.x + j * NCE_THREADS) * 4;              // 4 consecutive floats of one z row
      const int z = i / ROW, w = i % ROW;
      float* d = base + (z >> 3) * team_f + (z >> 3) * 8 + (z & 7) * ROWP + w;
      d[0] = wv[j].x; d[1] = wv[j].y; d[2] = wv[j].z; d[3] = wv[j].w;
    }
    __syncthreads();
    const float* wr = base + (zu >> 3) * team_f + (zu >> 3) * 8 + (zu & 7) * ROWP;    // bank of lane zu: zu + const
#pragma unroll
    for (int cc = 0; cc < NCE_CD; ++cc) {
      w0[cc] = wr[cc * KR + (ru % KR)];
      if (TWO_K) w1[cc] = wr[cc * KR + ((ru + 4) % KR)];
    }
    __syncthreads();                                   // slots nst - 1 may now be overwritten by the first in-loop issue
  }
  const int bar_id = 1 + team;

  for (int it = 0; it < my; ++it) {
    if (tt == 0) issue(it + p.nst - 1);
    const int slot = it % p.nst;
    const int b0 = (stream + it * nstreams) * S::SPS;
    const int ns = min(S::SPS, p.B - b0);
    nce_mbar_wait(bar0 + 8u * slot, (it / p.nst) & 1);
    const float* st = stg + (size_t)slot * S::STAGE_F;
    const float* st_pos = st + S::SPS * S::NEG_F;
    const float* st_c = st_pos + S::SPS * S::ROWB;
    // ---- U phase: u[row][z] = sum_c W[z][c][k] c[sample][c]   (four partial sums per output: short dependency chains)
    {
      const int r0 = ru, r1 = ru + 4;
      const float* c0 = st_c + (r0 / KR) * NCE_CD;
      const float* c1 = st_c + (r1 / KR) * NCE_CD;
      float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int cc = 0; cc < NCE_CD; cc += 4) {
        const float4 ca = *reinterpret_cast<const float4*>(c0 + cc);
        const float4 cb = *reinterpret_cast<const float4*>(c1 + cc);
        const float* wb = TWO_K ? w1 : w0;
        a0[0] = fmaf(ca.x, w0[cc], a0[0]); a0[1] = fmaf(ca.y, w0[cc + 1], a0[1]);
        a0[2] = fmaf(ca.z, w0[cc + 2], a0[2]); a0[3] = fmaf(ca.w, w0[cc + 3], a0[3]);
        a1[0] = fmaf(cb.x, wb[cc], a1[0]); a1[1] = fmaf(cb.y, wb[cc + 1], a1[1]);
        a1[2] = fmaf(cb.z, wb[cc + 2], a1[2]); a1[3] = fmaf(cb.w, wb[cc + 3], a1[3]);
      }
      const float u0 = (a0[0] + a0[1]) + (a0[2] + a0[3]), u1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
      us[r0 * NCE_ZD + zu] = u0;
      us[r1 * NCE_ZD + zu] = u1;
      float* ug = p.u + (size_t)b0 * S::ROWB;
      if (r0 / KR < ns) ug[r0 * NCE_ZD + zu] = u0;
      if (r1 / KR < ns) ug[r1 * NCE_ZD + zu] = u1;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NCE_TT) : "memory");
    // ---- S phase: one score per thread, then log-sum-exp / accuracy among the 16 lanes of a (sample, k) row
    {
      const int sI = rs / KR, k = rs % KR;
      const bool act = sI < ns;
      const float* yv = (nn < NCE_N) ? st + sI * S::NEG_F + nn * S::ROWB + k * NCE_ZD : st_pos + sI * S::ROWB + k * NCE_ZD;
      const float a = act ? rot_dot<NCE_NC>(us + rs * NCE_ZD, yv, lane) : 0.f;
      const size_t bk = (size_t)(b0 + sI) * KR + k;
      if (act) {
        if (nn < NCE_N) p.neg[bk * NCE_N + nn] = a;
        else p.pos[bk] = a;
      }
      const float pv = __shfl_sync(0xffffffffu, a, (lane & 16) + NCE_N);
      float nm = (nn < NCE_N) ? a : -INFINITY;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) nm = fmaxf(nm, __shfl_xor_sync(0xffffffffu, nm, o));
      const float mx = fmaxf(nm, pv);
      float se = expf(a - mx);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
      if (act && nn == NCE_N) {
        const float l = mx + logf(se);
        p.lse[bk] = l;
        p.nll[bk] = l - pv;                         // -(pos - lse): summed by the caller (vqcpc_reduce_sum)
        p.correct[bk] = (pv > nm) ? 1 : 0;
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NCE_TT) : "memory");
    if (tt == 0) nce_mbar_arrive(bar0 + 8u * (p.nst + slot));       // every read of the stage is done
  }
}

// ------------------------------------------------------------------------------------------------------------
// Streaming backward (same shapes, same team / ring structure as the forward).  Nothing saved by the forward is read:
// u, the scores and the log-sum-exp are recomputed from the staged rows with the forward's own code (a few hundred
// flops per sample against 32 KB of traffic), so the kernel streams y once and writes dy once:
//   U, S phases as in the forward; coefficient of score n of row (sample, k):  pi_n * scale for a negative,
//                                  -(1 - pi_pos) * scale for the positive                     (pi = softmax of the row)
//   G phase (thread per (row, z)): du[z] = sum_n coef_n y[n][z];  dy[n][z] = coef_n u[z]  (coalesced 128-byte stores);
//                                  dW[k][z][:] += du[z] c[sample][:]  in REGISTERS for the whole kernel
//   C phase (thread per (c, pair of rows)): dc[sample][c] = sum_{k,z} W[z][c][k] du[k][z], W resident in shared memory
// End: the per-thread dW blocks are summed over the 4 teams in shared memory, over the 4 CTAs of a cluster through
// distributed shared memory, and only then added to global memory (the L2 atomic units sustain ~16 red/clk: per-thread
// reds would cost more than the whole stream).
struct NceBwdParams {
  const float *c, *W, *ypos, *yneg, *g_dev;
  float scale;
  int B, nst;
  float *dc, *dW, *dypos, *dyneg;
};

template <int KR>
struct NceBwdShape {
  using F = NceShape<KR>;
  static constexpr int WS = NCE_CD + 4;                       // padded W row: Wsm[k][z][c]
  static constexpr int W_F = KR * NCE_ZD * WS;
  static constexpr int TEAM_X = 8 * NCE_ZD + 8 * NCE_GRP + 8 * NCE_ZD + 4 * NCE_CD;   // us, cf, dus, dcp
  static constexpr int RED_F = KR * NCE_CD * NCE_ZD;          // one team's dW block [k][c][z]
  __host__ __device__ static constexpr int team_f(int nst) { return (nst * F::STAGE_F > RED_F ? nst * F::STAGE_F : RED_F) + TEAM_X; }
  static size_t smem(int nst) { return ((size_t)W_F + (size_t)NCE_TEAMS * team_f(nst)) * 4 + NCE_TEAMS * 2 * (size_t)nst * 8 + 16; }
};

template <int KR>
__global__ void __launch_bounds__(NCE_THREADS, 1)
infonce_stream_bwd_kernel(NceBwdParams p) {
  using S = NceShape<KR>;
  using SB = NceBwdShape<KR>;
  extern __shared__ __align__(16) uint8_t nce_smem[];
  const int team = threadIdx.x / NCE_TT, tt = threadIdx.x % NCE_TT, lane = threadIdx.x & 31;
  const int team_f = SB::team_f(p.nst);
  float* Wsm = reinterpret_cast<float*>(nce_smem);                          // [KR][zd][cd + 4]
  float* tbase = Wsm + SB::W_F + (size_t)team * team_f;
  float* stg = tbase;                                                        // [nst][STAGE_F] (later: dW block of the team)
  float* us = tbase + team_f - SB::TEAM_X;                                   // [8][zd]
  float* cf = us + 8 * NCE_ZD;                                               // [8][16] coefficients
  float* dus = cf + 8 * NCE_GRP;                                             // [8][zd]
  float* dcp = dus + 8 * NCE_ZD;                                             // [4 parts][cd]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Wsm + SB::W_F + (size_t)NCE_TEAMS * team_f) + team * 2 * p.nst;
  const uint32_t bar0 = nce_smem_u32(bars);            // full[nst], empty[nst] of this team

  if (tt == 0) {
    for (int i = 0; i < 2 * p.nst; ++i) nce_mbar_init(bar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int ngroups = (p.B + S::SPS - 1) / S::SPS;
  const int stream = blockIdx.x * NCE_TEAMS + team, nstreams = gridDim.x * NCE_TEAMS;
  const int my = stream < ngroups ? (ngroups - stream + nstreams - 1) / nstreams : 0;
  auto issue = [&](int i) {
    if (i >= my) return;
    const int slot = i % p.nst;
    if (i >= p.nst) nce_mbar_wait(bar0 + 8u * (p.nst + slot), ((i / p.nst) & 1) ^ 1);
    const int b0 = (stream + i * nstreams) * S::SPS;
    const uint32_t ns = (uint32_t)min(S::SPS, p.B - b0);
    const uint32_t fb = bar0 + 8u * slot;
    float* d = stg + (size_t)slot * S::STAGE_F;
    nce_mbar_expect_tx(fb, ns * (uint32_t)(S::NEG_F + S::ROWB + NCE_CD) * 4u);
    nce_bulk_g2s(nce_smem_u32(d), p.yneg + (size_t)b0 * S::NEG_F, ns * (uint32_t)S::NEG_F * 4u, fb);
    nce_bulk_g2s(nce_smem_u32(d + S::SPS * S::NEG_F), p.ypos + (size_t)b0 * S::ROWB, ns * (uint32_t)S::ROWB * 4u, fb);
    nce_bulk_g2s(nce_smem_u32(d + S::SPS * (S::NEG_F + S::ROWB)), p.c + (size_t)b0 * NCE_CD, ns * (uint32_t)NCE_CD * 4u, fb);
  };
  {
    // W (zd, cd, KR) -> Wsm[k][z][c]: coalesced float4 loads issued before the first bulk copies, scattered stores
    constexpr int W_N = NCE_ZD * NCE_CD * KR, PER = W_N / 4 / NCE_THREADS;
    static_assert(W_N % (4 * NCE_THREADS) == 0 && (KR % 4 == 0 || KR == 2), "W must split evenly over the CTA");
    float4 wv[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) wv[j] = __ldg(reinterpret_cast<const float4*>(p.W) + threadIdx.x + j * NCE_THREADS);
    __syncwarp();                                      // (lane 0 of each team initialised barriers: reconverge first - synccheck)
    __syncthreads();                                   // barriers initialised
    if (tt == 0)
      for (int i = 0; i < p.nst - 1; ++i) issue(i);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const float v4[4] = {wv[j].x, wv[j].y, wv[j].z, wv[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = (threadIdx.x + j * NCE_THREADS) * 4 + e;
        const int k = i % KR, cc = (i / KR) % NCE_CD, z = i / (KR * NCE_CD);
        Wsm[(k * NCE_ZD + z) * SB::WS + cc] = v4[e];
      }
    }
    __syncthreads();
  }
  const float scale = p.scale * (p.g_dev ? __ldg(p.g_dev) : 1.f);
  const int ru = tt >> 5, zu = tt & 31;                // G phase: rows ru, ru + 4, column zu
  const int rs = tt >> 4, nn = tt & 15;                // S phase: row rs, score nn (15 = positive)
  constexpr bool TWO_K = KR == 8;                      // rows ru and ru + 4 belong to different k only when KR = 8
  const int k0 = ru % KR, k1 = (ru + 4) % KR;
  float dw0[NCE_CD], dw1[TWO_K ? NCE_CD : 1];
#pragma unroll
  for (int cc = 0; cc < NCE_CD; ++cc) { dw0[cc] = 0.f; if (TWO_K) dw1[cc] = 0.f; }
  const int bar_id = 1 + team;
#define TEAM_SYNC() asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NCE_TT) : "memory")

  for (int it = 0; it < my; ++it) {
    if (tt == 0) issue(it + p.nst - 1);
    const int slot = it % p.nst;
    const int b0 = (stream + it * nstreams) * S::SPS;
    const int ns = min(S::SPS, p.B - b0);
    nce_mbar_wait(bar0 + 8u * slot, (it / p.nst) & 1);
    const float* st = stg + (size_t)slot * S::STAGE_F;
    const float* st_pos = st + S::SPS * S::NEG_F;
    const float* st_c = st_pos + S::SPS * S::ROWB;
    const int r0 = ru, r1 = ru + 4;
    const int s0 = r0 / KR, s1 = r1 / KR;
    // ---- U phase (same arithmetic as the forward)
    float u0, u1;
    {
      const float* w0 = Wsm + (k0 * NCE_ZD + zu) * SB::WS;
      const float* w1 = Wsm + (k1 * NCE_ZD + zu) * SB::WS;
      const float* c0 = st_c + s0 * NCE_CD;
      const float* c1 = st_c + s1 * NCE_CD;
      float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int cc = 0; cc < NCE_CD; cc += 4) {
        const float4 wa = *reinterpret_cast<const float4*>(w0 + cc);
        const float4 wb = TWO_K ? *reinterpret_cast<const float4*>(w1 + cc) : wa;
        const float4 ca = *reinterpret_cast<const float4*>(c0 + cc);
        const float4 cb = *reinterpret_cast<const float4*>(c1 + cc);
        a0[0] = fmaf(ca.x, wa.x, a0[0]); a0[1] = fmaf(ca.y, wa.y, a0[1]); a0[2] = fmaf(ca.z, wa.z, a0[2]); a0[3] = fmaf(ca.w, wa.w, a0[3]);
        a1[0] = fmaf(cb.x, wb.x, a1[0]); a1[1] = fmaf(cb.y, wb.y, a1[1]); a1[2] = fmaf(cb.z, wb.z, a1[2]); a1[3] = fmaf(cb.w, wb.w, a1[3]);
      }
      u0 = (a0[0] + a0[1]) + (a0[2] + a0[3]);
      u1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
      us[r0 * NCE_ZD + zu] = u0;
      us[r1 * NCE_ZD + zu] = u1;
    }
    TEAM_SYNC();
    // ---- S phase: score, softmax of the row, coefficient
    {
      const int sI = rs / KR, k = rs % KR;
      const bool act = sI < ns;
      const float* yv = (nn < NCE_N) ? st + sI * S::NEG_F + nn * S::ROWB + k * NCE_ZD : st_pos + sI * S::ROWB + k * NCE_ZD;
      const float a = act ? rot_dot<NCE_NC>(us + rs * NCE_ZD, yv, lane) : 0.f;
      const float pv = __shfl_sync(0xffffffffu, a, (lane & 16) + NCE_N);
      float nm = (nn < NCE_N) ? a : -INFINITY;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) nm = fmaxf(nm, __shfl_xor_sync(0xffffffffu, nm, o));
      const float mx = fmaxf(nm, pv);
      const float ex = expf(a - mx);
      float se = ex;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
      const float pi = expf(a - (mx + logf(se)));      // exp(score - lse), as the warp-per-(b, k) kernel computes it
      cf[rs * NCE_GRP + nn] = act ? ((nn < NCE_N) ? pi * scale : -(1.f - pi) * scale) : 0.f;
    }
    TEAM_SYNC();
    // ---- G phase
    {
      const float* cf0 = cf + r0 * NCE_GRP;
      const float* cf1 = cf + r1 * NCE_GRP;
      float du0 = 0.f, du1 = 0.f;
      float* dn0 = p.dyneg + ((size_t)(b0 + s0) * NCE_N * KR + k0) * NCE_ZD + zu;     // + n * KR * zd
      float* dn1 = p.dyneg + ((size_t)(b0 + s1) * NCE_N * KR + k1) * NCE_ZD + zu;
      const float* y0 = st + s0 * S::NEG_F + k0 * NCE_ZD + zu;                         // + n * ROWB
      const float* y1 = st + s1 * S::NEG_F + k1 * NCE_ZD + zu;
      const bool a0 = s0 < ns, a1 = s1 < ns;
#pragma unroll
      for (int n4 = 0; n4 < NCE_GRP; n4 += 4) {
        const float4 q0 = *reinterpret_cast<const float4*>(cf0 + n4);
        const float4 q1 = *reinterpret_cast<const float4*>(cf1 + n4);
        const float e0[4] = {q0.x, q0.y, q0.z, q0.w}, e1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int n = n4 + e;
          if (n < NCE_N) {
            du0 = fmaf(e0[e], y0[n * S::ROWB], du0);
            du1 = fmaf(e1[e], y1[n * S::ROWB], du1);
            if (a0) dn0[(size_t)n * KR * NCE_ZD] = e0[e] * u0;
            if (a1) dn1[(size_t)n * KR * NCE_ZD] = e1[e] * u1;
          } else {
            du0 = fmaf(e0[e], st_pos[s0 * S::ROWB + k0 * NCE_ZD + zu], du0);
            du1 = fmaf(e1[e], st_pos[s1 * S::ROWB + k1 * NCE_ZD + zu], du1);
            if (a0) p.dypos[((size_t)(b0 + s0) * KR + k0) * NCE_ZD + zu] = e0[e] * u0;
            if (a1) p.dypos[((size_t)(b0 + s1) * KR + k1) * NCE_ZD + zu] = e1[e] * u1;
          }
        }
      }
      if (!a0) du0 = 0.f;                              // rows past the batch end hold stale shared memory
      if (!a1) du1 = 0.f;
      dus[r0 * NCE_ZD + zu] = du0;
      dus[r1 * NCE_ZD + zu] = du1;
      const float* c0 = st_c + s0 * NCE_CD;
      const float* c1 = st_c + s1 * NCE_CD;
#pragma unroll
      for (int cc = 0; cc < NCE_CD; cc += 4) {
        const float4 ca = *reinterpret_cast<const float4*>(c0 + cc);
        const float4 cb = *reinterpret_cast<const float4*>(c1 + cc);
        dw0[cc] = fmaf(du0, ca.x, dw0[cc]); dw0[cc + 1] = fmaf(du0, ca.y, dw0[cc + 1]);
        dw0[cc + 2] = fmaf(du0, ca.z, dw0[cc + 2]); dw0[cc + 3] = fmaf(du0, ca.w, dw0[cc + 3]);
        float* d1 = TWO_K ? dw1 : dw0;
        d1[cc] = fmaf(du1, cb.x, d1[cc]); d1[cc + 1] = fmaf(du1, cb.y, d1[cc + 1]);
        d1[cc + 2] = fmaf(du1, cb.z, d1[cc + 2]); d1[cc + 3] = fmaf(du1, cb.w, d1[cc + 3]);
      }
    }
    TEAM_SYNC();
    if (tt == 0) nce_mbar_arrive(bar0 + 8u * (p.nst + slot));       // every read of the stage is done
    // ---- C phase: thread (c, part): rows 2 part, 2 part + 1
    {
      const int cc = tt & 31, part = tt >> 5;
      float acc = 0.f;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = 2 * part + rr;
        const float* wk = Wsm + (size_t)((r % KR) * NCE_ZD) * SB::WS + cc;
        const float* dr = dus + r * NCE_ZD;
#pragma unroll
        for (int z = 0; z < NCE_ZD; z += 4) {
          const float4 d4 = *reinterpret_cast<const float4*>(dr + z);
          acc = fmaf(wk[(z + 0) * SB::WS], d4.x, acc);
          acc = fmaf(wk[(z + 1) * SB::WS], d4.y, acc);
          acc = fmaf(wk[(z + 2) * SB::WS], d4.z, acc);
          acc = fmaf(wk[(z + 3) * SB::WS], d4.w, acc);
        }
      }
      dcp[part * NCE_CD + cc] = acc;
    }
    TEAM_SYNC();
    if (tt < S::SPS * NCE_CD) {
      const int sI = tt >> 5, cc = tt & 31;
      if (sI < ns) {
        constexpr int PPS = KR / 2;                    // parts per sample
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < PPS; ++q) acc += dcp[(sI * PPS + q) * NCE_CD + cc];
        p.dc[(size_t)(b0 + sI) * NCE_CD + cc] += acc;   // this team owns the sample: plain read-modify-write
      }
    }
    // (dcp / cf / us / dus are next written after barriers of the next iteration that these threads also pass)
  }

  // ---- dW: registers -> team block [k][c][z] (reusing the ring), sum over teams, over the cluster, one red per word
  __syncthreads();                                     // every stage of every team consumed, no bulk copy in flight
  {
    float* blk = tbase;                                // RED_F floats
    constexpr int ROUNDS = KR >= 4 ? 1 : 4 / KR;       // warps of a team that share k take turns
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      if (ru / KR == rd || KR >= 4) {
#pragma unroll
        for (int cc = 0; cc < NCE_CD; ++cc) {
          float* d = blk + (k0 * NCE_CD + cc) * NCE_ZD + zu;
          if (rd == 0) *d = dw0[cc]; else *d += dw0[cc];
          if (TWO_K) blk[(k1 * NCE_CD + cc) * NCE_ZD + zu] = dw1[cc];
        }
      }
      TEAM_SYNC();
    }
  }
  __syncthreads();
  float* red0 = Wsm + SB::W_F;                         // team 0's block becomes the CTA's block
  for (int i = threadIdx.x; i < SB::RED_F; i += NCE_THREADS) {
    float a = red0[i];
#pragma unroll
    for (int t = 1; t < NCE_TEAMS; ++t) a += red0[(size_t)t * team_f + i];
    red0[i] = a;
  }
  unsigned crank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  {
    const int slice = (SB::RED_F + (int)csize - 1) / (int)csize;
    const int lo = (int)crank * slice, hi = min((int)SB::RED_F, lo + slice);
    const uint32_t sbase = nce_smem_u32(red0);
    for (int i = lo + threadIdx.x; i < hi; i += NCE_THREADS) {
      float acc = 0.f;
      for (unsigned r = 0; r < csize; ++r) {
        uint32_t ra;
        float t;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sbase + 4u * (uint32_t)i), "r"(r));
        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t) : "r"(ra));
        acc += t;
      }
      const int z = i % NCE_ZD, cc = (i / NCE_ZD) % NCE_CD, k = i / (NCE_ZD * NCE_CD);
      if (acc != 0.f) atomicAdd(p.dW + ((size_t)z * NCE_CD + cc) * KR + k, acc);
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#undef TEAM_SYNC
}

template <int KR>
int nce_stream_bwd_launch(NceBwdParams p, cudaStream_t st) {
  using S = NceShape<KR>;
  using SB = NceBwdShape<KR>;
  int nst = 0;
  for (int t = 3; t >= 2; --t)
    if (SB::smem(t) <= 220 * 1024) { nst = t; break; }
  if (!nst) return 1;
  p.nst = nst;
  const size_t sm = SB::smem(nst);
  VQ_CUDA_CALL(cudaFuncSetAttribute(infonce_stream_bwd_kernel<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const int ngroups = (p.B + S::SPS - 1) / S::SPS;
  int grid = std::min((ngroups + NCE_TEAMS - 1) / NCE_TEAMS, vq_sm_count());
  const int cl = grid >= 4 ? 4 : 1;
  grid = std::max(1, grid / cl * cl);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NCE_THREADS);
  cfg.dynamicSmemBytes = sm;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VQ_CUDA_CALL(cudaLaunchKernelEx(&cfg, infonce_stream_bwd_kernel<KR>, p));
  VQ_CUDA_LAUNCH_CHECK("infonce_stream_bwd_kernel");
  return VQ_OK;
}

template <int KR>
int nce_stream_launch(NceParams p, cudaStream_t st) {
  using S = NceShape<KR>;
  int nst = 0;
  for (int t = 4; t >= 2; --t)
    if (S::smem(t) <= 220 * 1024) { nst = t; break; }
  if (!nst) return 1;
  p.nst = nst;
  const size_t sm = S::smem(nst);
  VQ_CUDA_CALL(cudaFuncSetAttribute(infonce_stream_fwd_kernel<KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const int ngroups = (p.B + S::SPS - 1) / S::SPS;
  const int grid = std::max(1, std::min((ngroups + NCE_TEAMS - 1) / NCE_TEAMS, vq_sm_count()));
  infonce_stream_fwd_kernel<KR><<<grid, NCE_THREADS, sm, st>>>(p);
  VQ_CUDA_LAUNCH_CHECK("infonce_stream_fwd_kernel");
  return VQ_OK;
}

}  // namespace

/* tuning / test hook: 0 = automatic, 1 = warp-per-(b, k) kernel, 2 = streaming (cp.async.bulk ring) kernel */
VQ_API int vqcpc_infonce_set_algo(int algo) {
  g_nce_algo = (algo == 1 || algo == 2) ? algo : 0;
  return VQ_OK;
}

VQ_API int vqcpc_infonce_fwd(const float* c, const float* W, const float* y_pos, const float* y_neg, int B, int N, int KR,
                             int zd, int cd, float* u, float* pos, float* neg, float* lse, float* nll, int* correct, void* stream) {
  VQ_CHECK_ARG(zd >= 1 && zd <= MAXZ, "infonce_fwd: z_dim must be in 1..%d (got %d)", MAXZ, zd);
  VQ_CHECK_ARG(KR >= 1 && KR <= 65535 && N >= 1, "infonce_fwd: bad KR=%d / N=%d", KR, N);
  if (B == 0) return VQ_OK;
  cudaStream_t st = vq_stream(stream);
  // streaming kernel: 16-byte granular rows, at least two ring stages in shared memory, enough samples to fill the GPU
  const bool aligned = (((uintptr_t)c | (uintptr_t)y_pos | (uintptr_t)y_neg) & 15) == 0;
  if (g_nce_algo != 1 && aligned && zd == NCE_ZD && cd == NCE_CD && N == NCE_N && (KR == 2 || KR == 4 || KR == 8) &&
      (g_nce_algo == 2 || B >= 64)) {
    NceParams p{c, W, y_pos, y_neg, B, 0, u, pos, neg, lse, nll, correct};
    return KR == 8 ? nce_stream_launch<8>(p, st) : KR == 4 ? nce_stream_launch<4>(p, st) : nce_stream_launch<2>(p, st);
  }
  VQ_CHECK_ARG(g_nce_algo != 2, "infonce_fwd: the streaming kernel needs 16-byte aligned inputs, zd = cd = 32, N = 15 and KR in {2, 4, 8}");
  size_t smem = (size_t)zd * (cd + 1) * sizeof(float);
  VQ_CHECK_ARG(smem <= 48 * 1024, "infonce_fwd: W[:, :, k] tile (%zu B) exceeds 48 KB", smem);
  int gx = (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(B, 8), (int64_t)vq_sm_count() * 8 / KR + 1));
  dim3 grid(gx, KR);
  int zpl = (zd + 31) / 32;
  if (zpl == 1) infonce_fwd_kernel<1><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, B, N, KR, zd, cd, u, pos, neg, lse, nll, correct);
  else if (zpl == 2) infonce_fwd_kernel<2><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, B, N, KR, zd, cd, u, pos, neg, lse, nll, correct);
  else infonce_fwd_kernel<4><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, B, N, KR, zd, cd, u, pos, neg, lse, nll, correct);
  VQ_CUDA_LAUNCH_CHECK("infonce_fwd_kernel");
  return VQ_OK;
}

// d_c (B, cd) and d_W (zd, cd, KR) are ACCUMULATED; d_ypos / d_yneg are overwritten.
// upstream gradient of sum(nll) = scale * g_dev[0] (g_dev: device scalar or NULL -> 1), so no host sync is needed.
VQ_API int vqcpc_infonce_bwd(const float* c, const float* W, const float* y_pos, const float* y_neg, const float* u,
                             const float* pos, const float* neg, const float* lse, const float* g_dev, float scale, int B, int N,
                             int KR, int zd,
                             int cd, float* d_c, float* d_W, float* d_ypos, float* d_yneg, void* stream) {
  VQ_CHECK_ARG(zd >= 1 && zd <= MAXZ, "infonce_bwd: z_dim must be in 1..%d (got %d)", MAXZ, zd);
  if (B == 0) return VQ_OK;
  {
    const bool aligned = (((uintptr_t)c | (uintptr_t)y_pos | (uintptr_t)y_neg | (uintptr_t)W) & 15) == 0;
    if (g_nce_algo != 1 && aligned && zd == NCE_ZD && cd == NCE_CD && N == NCE_N && (KR == 2 || KR == 4 || KR == 8) &&
        (g_nce_algo == 2 || B >= 64)) {
      NceBwdParams p{c, W, y_pos, y_neg, g_dev, scale, B, 0, d_c, d_W, d_ypos, d_yneg};
      cudaStream_t st2 = vq_stream(stream);
      return KR == 8 ? nce_stream_bwd_launch<8>(p, st2) : KR == 4 ? nce_stream_bwd_launch<4>(p, st2) : nce_stream_bwd_launch<2>(p, st2);
    }
    VQ_CHECK_ARG(g_nce_algo != 2, "infonce_bwd: the streaming kernel needs 16-byte aligned inputs, zd = cd = 32, N = 15 and KR in {2, 4, 8}");
  }
  size_t smem = (size_t)2 * zd * (cd + 1) * sizeof(float);
  VQ_CHECK_ARG(smem <= 48 * 1024, "infonce_bwd: W tile (%zu B) exceeds 48 KB", smem);
  int gx = (int)std::max<int64_t>(1, std::min<int64_t>(vq_cdiv(B, 8), (int64_t)vq_sm_count() * 4 / KR + 1));
  dim3 grid(gx, KR);
  cudaStream_t st = vq_stream(stream);
  int zpl = (zd + 31) / 32;
  if (zpl == 1) infonce_bwd_kernel<1><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, u, pos, neg, lse, g_dev, scale, B, N, KR, zd, cd, d_c, d_W, d_ypos, d_yneg);
  else if (zpl == 2) infonce_bwd_kernel<2><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, u, pos, neg, lse, g_dev, scale, B, N, KR, zd, cd, d_c, d_W, d_ypos, d_yneg);
  else infonce_bwd_kernel<4><<<grid, 256, smem, st>>>(c, W, y_pos, y_neg, u, pos, neg, lse, g_dev, scale, B, N, KR, zd, cd, d_c, d_W, d_ypos, d_yneg);
  VQ_CUDA_LAUNCH_CHECK("infonce_bwd_kernel");
  return VQ_OK;
}
