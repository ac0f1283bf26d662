// Fused multi-head attention BACKWARD for B200 (sm_100a), head dim 64 (3DETR encoder, CLIP) and 128
// (3DETR decoder).  Two tcgen05 kernels; the (Lq x Lk) probability / score-gradient tiles never leave
// the SM: they are produced in registers from TMEM accumulators and handed back to the tensor core as
// TMEM A operands (tcgen05.mma [a_tmem]).
//
//   attn_bwd_dq_kernel   CTA = (bh, 128 queries); loop over 64-key tiles j
//        S_j  = Qs K_j^T              dP_j = dO V_j^T           (TMEM 128 x 64 each, double-buffered)
//        dS_j = P_j o (m_j o dP_j - D)    with P_j = exp(S_j - LSE), m = dropout factor, D = rowsum(dO o O)
//        dQ  += dS_j K_j              (dS_j from TMEM; K_j as an MN-major B operand -- the same smem tile
//                                      that served S_j; accumulated in TMEM over j, scaled at the end)
//   attn_bwd_dkv_kernel  CTA = (bh, 128 keys); loop over 64-query tiles i
//        S_i^T  = K Qs_i^T            dP_i^T = V dO_i^T         (TMEM 128 x 64 each)
//        P~_i^T = P_i^T o m           dS_i^T = P_i^T o (m o dP_i^T - D_i)
//        dV += P~_i^T dO_i            dK += dS_i^T Qs_i         (A from TMEM, dO_i / Qs_i MN-major from the
//                                      tiles that served the scores; accumulated in TMEM over i)
//
// S is recomputed in both kernels (7 MMA groups instead of the 5 of a single fused pass) so that every
// accumulation stays inside one CTA: no atomics, deterministic.  Operands are split-bf16 planes like the
// forward; the backward uses 2 planes (3 cross products, ~16 mantissa bits), enough for the 2e-3
// gradient parity bar.  Only row-major packs of q, k, v, dO are needed (no transposed copies).
//
// The element-wise stage (exp, dropout stream, dS, split into planes, tcgen05.st) is the critical path, not the
// tensor pipe: EIGHT softmax warps per CTA -- two per TMEM lane quadrant, each taking 32 of a tile's 64 columns --
// and, where tensor memory allows (head dim 64), double-buffered S^T / dP^T in the dK/dV kernel so that the
// scores of tile i+1 are computed while tile i is in the softmax warps.
// Optional attention mask (bit-packed, 1 = key not visible to query; MaskedTransformerEncoder of the reference,
// models/transformer.py:146-211): one 64-bit word per (row, 64-column tile).
#include "../../include/coda_attention.h"
#include "attention_common.cuh"

using namespace coda;
using namespace coda::attn;

namespace {

constexpr int NS = 2;           // planes per operand in the backward
constexpr int NPROD = 3;

// D[bh][q] = sum_d dO * O   (both (Lq, B, H*hd)); one warp per (q, b, h)
__global__ void __launch_bounds__(256)
bwd_delta_kernel(int Lq, int B, int H, int hd, const float *__restrict__ dout, const float *__restrict__ out,
                 float *__restrict__ delta) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Lq * B * H) return;
  const int h = warp % H, b = (warp / H) % B, q = warp / (H * B);
  const size_t off = ((size_t)q * B + b) * H * hd + (size_t)h * hd;
  float s = 0.f;
  for (int d = lane; d < hd; d += 32) s += __ldg(dout + off + d) * __ldg(out + off + d);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) delta[(size_t)(b * H + h) * Lq + q] = s;
}

struct BwdMaps {
  CUtensorMap q[NS], k[NS], v[NS], dO[NS];
};

// 16 consecutive values of this thread's row -> NS bf16 planes, 8 TMEM columns (bf16 pairs) per plane.
// `taddr` addresses the row's lane and the first column of plane 0; planes are 32 columns apart.
__device__ __forceinline__ void st_planes16(uint32_t taddr, const float (&v)[16]) {
  uint32_t w[NS][8];
#pragma unroll
  for (int e = 0; e < 16; e += 2) {
    float r0 = v[e], r1 = v[e + 1];
#pragma unroll
    for (int pl = 0; pl < NS; ++pl) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(r0, r1);
      const uint32_t bits = *reinterpret_cast<const uint32_t *>(&h2);
      w[pl][e >> 1] = bits;
      if (pl + 1 < NS) {
        r0 -= __uint_as_float(bits << 16);
        r1 -= __uint_as_float(bits & 0xFFFF0000u);
      }
    }
  }
#pragma unroll
  for (int pl = 0; pl < NS; ++pl) tmem_st_32x8(taddr + (uint32_t)(pl * 32), w[pl]);
}

__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ====================================================================== dQ
template <int HD>
struct DqCfg {
  static constexpr int KB = HD / 64;
  static constexpr int T128 = 128 * 128;              // [128 x 64] bf16 block
  static constexpr int T64 = 64 * 128;                // [64 x 64] bf16 block
  static constexpr int QP = KB * T128;                // one plane of Qs / dO (128 rows x HD)
  static constexpr int KP = KB * T64;                 // one plane of K_j / V_j (64 rows x HD)
  // K_j stays in its stage until dQ_j has run, so with two stages the load of K_{j+1} could only start after
  // dQ_{j-1} had finished: one full TMA latency on the critical path of every tile (ncu: the softmax warps spent
  // half their time waiting for S).  Head dim 64 has the shared memory for a deeper ring; head dim 128 does not.
  static constexpr int KST = HD == 64 ? 4 : 2;        // K stages
  static constexpr int VST = HD == 64 ? 3 : 1;        // V stages
  static constexpr int Q_OFF = 0;
  static constexpr int DO_OFF = Q_OFF + NS * QP;
  static constexpr int K_OFF = DO_OFF + NS * QP;
  static constexpr int V_OFF = K_OFF + KST * NS * KP;
  static constexpr int TOTAL = V_OFF + VST * NS * KP;
  // TMEM columns: S / dP double-buffered, dQ accumulator, dS planes (A operand)
  static constexpr int S_COL = 0, DP_COL = 64, BUF_COLS = 128, DQ_COL = 256, DS_COL = 256 + HD;
  static constexpr int TMEM_COLS = 512;
  static_assert(DS_COL + NS * 32 <= 512, "TMEM budget");
};

template <int HD>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dq_kernel(const __grid_constant__ BwdMaps maps, int Lq, int Lk, int B, int H, float scale,
                   const float *__restrict__ lse, const float *__restrict__ delta, float *__restrict__ dq,
                   long long ld_dq, const unsigned long long *__restrict__ mask_q, float drop_p, uint32_t seed,
                   const uint32_t *__restrict__ seed_dev) {
  using C = DqCfg<HD>;
  constexpr int KB = C::KB;
  if (seed_dev) seed += __ldg(seed_dev);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t q_full, k_full[C::KST], k_empty[C::KST], v_full[C::VST], v_empty[C::VST],
      s_full[2], ds_full, ds_empty, dq_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, bh = blockIdx.y;
  const int ntiles = (Lk + 63) / 64;

  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < C::KST; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < C::VST; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
    mbar_init(&ds_full, 256); mbar_init(&ds_empty, 1); mbar_init(&dq_done, 1);
    mbar_fence_init_cluster();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (warp-uniform control flow, one elected lane issues) =====
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&q_full, (uint32_t)(2 * NS * C::QP));
#pragma unroll
      for (int p = 0; p < NS; ++p)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_3d(smem + C::Q_OFF + p * C::QP + kb * C::T128, &maps.q[p], &q_full, kb * 64, q0, bh);
          tma_load_3d(smem + C::DO_OFF + p * C::QP + kb * C::T128, &maps.dO[p], &q_full, kb * 64, q0, bh);
        }
    }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int ks = j % C::KST, vs = j % C::VST;
      mbar_wait(&k_empty[ks], ((uint32_t)(j / C::KST) & 1u) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&k_full[ks], (uint32_t)(NS * C::KP));
#pragma unroll
        for (int p = 0; p < NS; ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(smem + C::K_OFF + (ks * NS + p) * C::KP + kb * C::T64, &maps.k[p], &k_full[ks], kb * 64, j * 64, bh);
      }
      __syncwarp();
      mbar_wait(&v_empty[vs], ((uint32_t)(j / C::VST) & 1u) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&v_full[vs], (uint32_t)(NS * C::KP));
#pragma unroll
        for (int p = 0; p < NS; ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(smem + C::V_OFF + (vs * NS + p) * C::KP + kb * C::T64, &maps.v[p], &v_full[vs], kb * 64, j * 64, bh);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform control flow, one elected lane issues) =====
    constexpr uint32_t idesc_s = umma_idesc_f16(0, 128, 64);
    constexpr uint32_t idesc_dq = umma_idesc_f16(0, 128, HD, 0, 1);  // B = K_j, MN-major (hd contiguous)
    auto issue_scores = [&](int j) {
      const int ks = j % C::KST, vs = j % C::VST;
      mbar_wait(&k_full[ks], (uint32_t)(j / C::KST) & 1u);
      mbar_wait(&v_full[vs], (uint32_t)(j / C::VST) & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t ts = tm + (uint32_t)((j & 1) * C::BUF_COLS + C::S_COL);
        const uint32_t tp = tm + (uint32_t)((j & 1) * C::BUF_COLS + C::DP_COL);
#pragma unroll
        for (int p = 0; p < NPROD; ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t aq = umma_smem_desc_k_sw128(smem + C::Q_OFF + a_pa(NS, p) * C::QP + kb * C::T128);
            const uint64_t bk = umma_smem_desc_k_sw128(smem + C::K_OFF + (ks * NS + a_pb(NS, p)) * C::KP + kb * C::T64);
            const uint64_t ad = umma_smem_desc_k_sw128(smem + C::DO_OFF + a_pa(NS, p) * C::QP + kb * C::T128);
            const uint64_t bv = umma_smem_desc_k_sw128(smem + C::V_OFF + (vs * NS + a_pb(NS, p)) * C::KP + kb * C::T64);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16(ts, umma_desc_advance(aq, kk * 32), umma_desc_advance(bk, kk * 32), idesc_s, (uint32_t)((p | kb | kk) != 0));
              umma_f16(tp, umma_desc_advance(ad, kk * 32), umma_desc_advance(bv, kk * 32), idesc_s, (uint32_t)((p | kb | kk) != 0));
            }
          }
        umma_commit(&v_empty[vs]);
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    mbar_wait(&q_full, 0);
    issue_scores(0);
    for (int j = 0; j < ntiles; ++j) {
      // the other score buffer held tile j-1, consumed when ds_full(j-1) arrived (awaited last iteration)
      if (j + 1 < ntiles) issue_scores(j + 1);
      mbar_wait(&ds_full, (uint32_t)j & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        const int ks = j % C::KST;
#pragma unroll
        for (int p = 0; p < NPROD; ++p) {
          const uint32_t a = tm + (uint32_t)(C::DS_COL + a_pa(NS, p) * 32);
          const uint64_t bk = umma_smem_desc_mn_sw128(smem + C::K_OFF + (ks * NS + a_pb(NS, p)) * C::KP);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // 16 keys: 8 TMEM columns of A, 16 rows x 128 B of K_j
            umma_f16_ts(tm + C::DQ_COL, a + kk * 8, umma_desc_advance(bk, kk * 16 * 128), idesc_dq,
                        (uint32_t)((j | p | kk) != 0));
        }
        umma_commit(&k_empty[ks]);
        umma_commit(&ds_empty);
        if (j == ntiles - 1) umma_commit(&dq_done);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int qq = (warp - 4) & 3, half = (warp - 4) >> 2, row = qq * 32 + lane;   // half: columns [32 half, +32) of a tile
    const uint32_t lane_base = (uint32_t)(qq * 32) << 16;
    const int qrow = q0 + row;
    const bool valid_row = qrow < Lq;
    const unsigned long long *mrow =
        mask_q ? mask_q + ((size_t)(bh / H) * Lq + (valid_row ? qrow : 0)) * (size_t)ntiles : nullptr;
    const float lse2 = (valid_row ? __ldg(lse + (size_t)bh * Lq + qrow) : 0.f) * LOG2E;
    const float d_r = valid_row ? __ldg(delta + (size_t)bh * Lq + qrow) : 0.f;
    const bool dropout = drop_p > 0.f;
    const uint32_t thresh32 = drop_thresh32(drop_p);
    const float keep_scale = dropout ? 1.0f / (1.0f - drop_p) : 1.0f;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&s_full[j & 1], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t sr[32], pr[32];
      const uint32_t tb = tm + (uint32_t)((j & 1) * C::BUF_COLS) + lane_base + (uint32_t)(half * 32);
      tmem_ld_32x32(tb + C::S_COL, sr);
      tmem_ld_32x32(tb + C::DP_COL, pr);
      tmem_ld_wait();
      // the previous tile's dS must have been consumed by its MMAs before it is overwritten
      if (j > 0) { mbar_wait(&ds_empty, ((uint32_t)j & 1u) ^ 1u); tc_fence_after(); }
      const int kvalid = Lk - j * 64;
      uint32_t ts = 0;
      if (dropout) ts = drop_tile_seed(seed, (uint32_t)bh, (uint32_t)qrow, (uint32_t)j);
      uint32_t mbits = 0;       // this half's 32 mask bits (1 = masked)
      if (mrow) mbits = (uint32_t)(__ldg(mrow + j) >> (half * 32));
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int ch = half * 2 + c2;
        float ds[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int cc = c2 * 16 + e, c = half * 32 + cc;     // column within this half / within the tile
          const float s = __uint_as_float(sr[cc]);
          const bool vis = c < kvalid && valid_row && !((mbits >> cc) & 1u);
          const float p = vis ? ex2_approx_b(fmaf(s, LOG2E, -lse2)) : 0.f;
          float dp = __uint_as_float(pr[cc]);
          if (dropout) dp = (ts * kLcgJump.a[c] + kLcgJump.c[c] >= thresh32) ? dp * keep_scale : 0.f;
          ds[e] = p * (dp - d_r);
        }
        st_planes16(tm + lane_base + (uint32_t)(C::DS_COL + ch * 8), ds);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full);
    }
    mbar_wait(&dq_done, 0);
    tc_fence_after();
    {
      const int b = bh / H, h = bh - b * H;
      float *orow = dq + ((size_t)qrow * B + b) * (size_t)ld_dq + (size_t)h * HD;
#pragma unroll
      for (int c1 = 0; c1 < HD / 2; c1 += 32) {
        const int c0 = half * (HD / 2) + c1;       // each half of the softmax warps stores half of the head dim
        uint32_t r[32];
        tmem_ld_32x32(tm + lane_base + (uint32_t)(C::DQ_COL + c0), r);   // warp-collective: every lane executes it
        tmem_ld_wait();
        if (valid_row) {
#pragma unroll
          for (int t = 0; t < 32; t += 4)
            *reinterpret_cast<float4 *>(orow + c0 + t) =
                make_float4(__uint_as_float(r[t]) * scale, __uint_as_float(r[t + 1]) * scale,
                            __uint_as_float(r[t + 2]) * scale, __uint_as_float(r[t + 3]) * scale);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_slot, C::TMEM_COLS);
}

// ====================================================================== dK, dV
template <int HD>
struct DkvCfg {
  static constexpr int KB = HD / 64;
  static constexpr int T128 = 128 * 128;
  static constexpr int T64 = 64 * 128;
  static constexpr int KP = KB * T128;                // one plane of K / V (128 rows x HD), resident
  static constexpr int QP = KB * T64;                 // one plane of Qs_i / dO_i (64 rows x HD)
  static constexpr int QST = HD == 64 ? 3 : 1;        // Qs_i / dO_i stages
  static constexpr int SB = HD == 64 ? 2 : 1;         // S^T / dP^T buffers (tensor memory is full at head dim 128)
  static constexpr int K_OFF = 0;
  static constexpr int V_OFF = K_OFF + NS * KP;
  static constexpr int Q_OFF = V_OFF + NS * KP;
  static constexpr int DO_OFF = Q_OFF + QST * NS * QP;
  static constexpr int TOTAL = DO_OFF + QST * NS * QP;
  // TMEM columns: S^T, dP^T, dV, dK accumulators, then the P~^T and dS^T planes (A operands)
  static constexpr int S_COL = 0, DP_COL = 64, BUF_COLS = 128, DV_COL = SB * 128, DK_COL = DV_COL + HD,
                       P_COL = DV_COL + 2 * HD, DS_COL = P_COL + NS * 32;
  static constexpr int TMEM_COLS = 512;
  static_assert(DS_COL + NS * 32 <= 512, "TMEM budget");
};

template <int HD>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dkv_kernel(const __grid_constant__ BwdMaps maps, int Lq, int Lk, int B, int H,
                    const float *__restrict__ lse, const float *__restrict__ delta, float *__restrict__ dk,
                    float *__restrict__ dv, long long ld_dk, long long ld_dv,
                    const unsigned long long *__restrict__ mask_k, float drop_p, uint32_t seed,
                    const uint32_t *__restrict__ seed_dev) {
  using C = DkvCfg<HD>;
  constexpr int KB = C::KB;
  if (seed_dev) seed += __ldg(seed_dev);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t kv_full, q_full[C::QST], q_empty[C::QST], s_full[C::SB], p_full, p_empty, acc_done;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lse[2][64], s_delta[2][64];
  __shared__ __align__(16) uint32_t s_seed[2][2][64];   // dropout stream seeds of (64 queries) x (this CTA's two key tiles)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, bh = blockIdx.y;
  const int ntiles = (Lq + 63) / 64;

  if (warp == 1 && lane == 0) {
    mbar_init(&kv_full, 1);
    for (int i = 0; i < C::QST; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < C::SB; ++i) mbar_init(&s_full[i], 1);
    mbar_init(&p_full, 256); mbar_init(&p_empty, 1); mbar_init(&acc_done, 1);
    mbar_fence_init_cluster();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&kv_full, (uint32_t)(2 * NS * C::KP));
#pragma unroll
      for (int p = 0; p < NS; ++p)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_3d(smem + C::K_OFF + p * C::KP + kb * C::T128, &maps.k[p], &kv_full, kb * 64, k0, bh);
          tma_load_3d(smem + C::V_OFF + p * C::KP + kb * C::T128, &maps.v[p], &kv_full, kb * 64, k0, bh);
        }
    }
    __syncwarp();
    for (int i = 0; i < ntiles; ++i) {
      const int qs = i % C::QST;
      mbar_wait(&q_empty[qs], ((uint32_t)(i / C::QST) & 1u) ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&q_full[qs], (uint32_t)(2 * NS * C::QP));
#pragma unroll
        for (int p = 0; p < NS; ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            tma_load_3d(smem + C::Q_OFF + (qs * NS + p) * C::QP + kb * C::T64, &maps.q[p], &q_full[qs], kb * 64, i * 64, bh);
            tma_load_3d(smem + C::DO_OFF + (qs * NS + p) * C::QP + kb * C::T64, &maps.dO[p], &q_full[qs], kb * 64, i * 64, bh);
          }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform control flow, one elected lane issues) =====
    constexpr uint32_t idesc_s = umma_idesc_f16(0, 128, 64);
    constexpr uint32_t idesc_acc = umma_idesc_f16(0, 128, HD, 0, 1);  // B = dO_i / Qs_i, MN-major
    mbar_wait(&kv_full, 0);
    auto issue_scores = [&](int i) {
      const int qs = i % C::QST, sb = i % C::SB;
      mbar_wait(&q_full[qs], (uint32_t)(i / C::QST) & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t ts = tm + (uint32_t)(sb * C::BUF_COLS + C::S_COL), tp = tm + (uint32_t)(sb * C::BUF_COLS + C::DP_COL);
#pragma unroll
        for (int p = 0; p < NPROD; ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t ak = umma_smem_desc_k_sw128(smem + C::K_OFF + a_pa(NS, p) * C::KP + kb * C::T128);
            const uint64_t bq = umma_smem_desc_k_sw128(smem + C::Q_OFF + (qs * NS + a_pb(NS, p)) * C::QP + kb * C::T64);
            const uint64_t av = umma_smem_desc_k_sw128(smem + C::V_OFF + a_pa(NS, p) * C::KP + kb * C::T128);
            const uint64_t bd = umma_smem_desc_k_sw128(smem + C::DO_OFF + (qs * NS + a_pb(NS, p)) * C::QP + kb * C::T64);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              umma_f16(ts, umma_desc_advance(ak, kk * 32), umma_desc_advance(bq, kk * 32), idesc_s, (uint32_t)((p | kb | kk) != 0));
              umma_f16(tp, umma_desc_advance(av, kk * 32), umma_desc_advance(bd, kk * 32), idesc_s, (uint32_t)((p | kb | kk) != 0));
            }
          }
        umma_commit(&s_full[sb]);
      }
      __syncwarp();
    };
    if (C::SB == 2) issue_scores(0);
    for (int i = 0; i < ntiles; ++i) {
      const int qs = i % C::QST;
      // two score buffers: tile i+1 is computed while tile i is in the softmax warps (its buffer held tile i-1,
      // read before p_full(i-1) arrived, which was awaited in the previous iteration)
      if (C::SB == 2) { if (i + 1 < ntiles) issue_scores(i + 1); } else issue_scores(i);
      mbar_wait(&p_full, (uint32_t)i & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int p = 0; p < NPROD; ++p) {
          const uint32_t ap = tm + (uint32_t)(C::P_COL + a_pa(NS, p) * 32);
          const uint32_t as = tm + (uint32_t)(C::DS_COL + a_pa(NS, p) * 32);
          const uint64_t bo = umma_smem_desc_mn_sw128(smem + C::DO_OFF + (qs * NS + a_pb(NS, p)) * C::QP);
          const uint64_t bq = umma_smem_desc_mn_sw128(smem + C::Q_OFF + (qs * NS + a_pb(NS, p)) * C::QP);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 16 queries: 8 TMEM columns of A, 16 rows x 128 B of dO_i / Qs_i
            umma_f16_ts(tm + C::DV_COL, ap + kk * 8, umma_desc_advance(bo, kk * 16 * 128), idesc_acc, (uint32_t)((i | p | kk) != 0));
            umma_f16_ts(tm + C::DK_COL, as + kk * 8, umma_desc_advance(bq, kk * 16 * 128), idesc_acc, (uint32_t)((i | p | kk) != 0));
          }
        }
        umma_commit(&q_empty[qs]);  // Qs_i / dO_i stage free
        umma_commit(&p_empty);      // P~^T / dS^T consumed
        if (i == ntiles - 1) umma_commit(&acc_done);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int qq = (warp - 4) & 3, half = (warp - 4) >> 2, row = qq * 32 + lane;  // key row; query columns [32 half, +32)
    const int t256 = (warp - 4) * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qq * 32) << 16;
    const int krow = k0 + row;
    const bool valid_row = krow < Lk;
    const unsigned long long *mrow =
        mask_k ? mask_k + ((size_t)(bh / H) * Lk + (valid_row ? krow : 0)) * (size_t)ntiles : nullptr;
    const bool dropout = drop_p > 0.f;
    const uint32_t thresh32 = drop_thresh32(drop_p);
    const float keep_scale = dropout ? 1.0f / (1.0f - drop_p) : 1.0f;
    const uint32_t ja = kLcgJump.a[krow & 63], jc = kLcgJump.c[krow & 63];  // this key's position in its tile
    // log-sum-exp (log2 units), D and the dropout stream seeds of a tile's 64 queries live in shared memory, double
    // buffered.  Tile i+1's values are fetched into a register at the top of iteration i and stored at its end: the
    // global-load latency hides behind the tile's arithmetic (it used to sit in front of every tile's barrier).
    auto fetch_tile = [&](int i) -> uint32_t {
      const int t = t256 & 63, qi = i * 64 + t;
      if (t256 < 64) return __float_as_uint(qi < Lq ? __ldg(lse + (size_t)bh * Lq + qi) * LOG2E : 0.f);
      if (t256 < 128) return __float_as_uint(qi < Lq ? __ldg(delta + (size_t)bh * Lq + qi) : 0.f);
      // one strong hash per (query, key tile), computed once and shared by the 64 key rows of that tile
      return dropout ? drop_tile_seed(seed, (uint32_t)bh, (uint32_t)qi, (uint32_t)((k0 >> 6) + ((t256 - 128) >> 6))) : 0u;
    };
    auto store_tile = [&](int i, uint32_t v) {
      const int t = t256 & 63;
      if (t256 < 64) s_lse[i & 1][t] = __uint_as_float(v);
      else if (t256 < 128) s_delta[i & 1][t] = __uint_as_float(v);
      else s_seed[i & 1][(t256 - 128) >> 6][t] = v;
    };
    store_tile(0, fetch_tile(0));
    for (int i = 0; i < ntiles; ++i) {
      // orders the writes of tile i's values (end of iteration i-1) before the reads below, and the reads of
      // tile i-1 before the rewrite of its buffer at the end of this iteration
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const uint32_t next_vals = i + 1 < ntiles ? fetch_tile(i + 1) : 0u;
      mbar_wait(&s_full[i % C::SB], (uint32_t)(i / C::SB) & 1u);
      tc_fence_after();
      uint32_t sr[32], pr[32];
      const uint32_t tb = tm + lane_base + (uint32_t)((i % C::SB) * C::BUF_COLS + half * 32);
      tmem_ld_32x32(tb + C::S_COL, sr);
      tmem_ld_32x32(tb + C::DP_COL, pr);
      tmem_ld_wait();
      if (i > 0) { mbar_wait(&p_empty, ((uint32_t)i & 1u) ^ 1u); tc_fence_after(); }  // previous P~^T / dS^T consumed
      const int qvalid = Lq - i * 64;
      uint32_t mbits = 0;       // this half's 32 mask bits (1 = this key is not visible to that query)
      if (mrow) mbits = (uint32_t)(__ldg(mrow + i) >> (half * 32));
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int ch = half * 2 + c2;
        float pt[16], ds[16];
        // the per-query values of these 16 columns: broadcast 128-bit shared loads (4 per array instead of 16)
        float lse16[16], dl16[16];
        uint32_t sd16[16];
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          const int c = half * 32 + c2 * 16 + v4 * 4;
          const float4 l = *reinterpret_cast<const float4 *>(&s_lse[i & 1][c]);
          const float4 d = *reinterpret_cast<const float4 *>(&s_delta[i & 1][c]);
          lse16[v4 * 4] = l.x; lse16[v4 * 4 + 1] = l.y; lse16[v4 * 4 + 2] = l.z; lse16[v4 * 4 + 3] = l.w;
          dl16[v4 * 4] = d.x; dl16[v4 * 4 + 1] = d.y; dl16[v4 * 4 + 2] = d.z; dl16[v4 * 4 + 3] = d.w;
          if (dropout) {
            const uint4 sd = *reinterpret_cast<const uint4 *>(&s_seed[i & 1][row >> 6][c]);
            sd16[v4 * 4] = sd.x; sd16[v4 * 4 + 1] = sd.y; sd16[v4 * 4 + 2] = sd.z; sd16[v4 * 4 + 3] = sd.w;
          }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int cc = c2 * 16 + e, c = half * 32 + cc;
          const float s = __uint_as_float(sr[cc]);
          const bool vis = c < qvalid && valid_row && !((mbits >> cc) & 1u);
          const float p = vis ? ex2_approx_b(fmaf(s, LOG2E, -lse16[e])) : 0.f;
          const float dp = __uint_as_float(pr[cc]);
          float m = 1.0f;
          if (dropout) m = (sd16[e] * ja + jc >= thresh32) ? keep_scale : 0.f;
          pt[e] = p * m;
          ds[e] = p * (dp * m - dl16[e]);
        }
        st_planes16(tm + lane_base + (uint32_t)(C::P_COL + ch * 8), pt);
        st_planes16(tm + lane_base + (uint32_t)(C::DS_COL + ch * 8), ds);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full);
      if (i + 1 < ntiles) store_tile(i + 1, next_vals);
    }
    mbar_wait(&acc_done, 0);
    tc_fence_after();
    const int b = bh / H, h = bh - b * H;
    // half 0 stores dV, half 1 stores dK
    float *orow = half == 0 ? dv + ((size_t)krow * B + b) * (size_t)ld_dv + (size_t)h * HD
                            : dk + ((size_t)krow * B + b) * (size_t)ld_dk + (size_t)h * HD;
    const uint32_t acc_col = half == 0 ? C::DV_COL : C::DK_COL;
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tm + lane_base + acc_col + (uint32_t)c0, r);
      tmem_ld_wait();
      if (valid_row) {
#pragma unroll
        for (int t = 0; t < 32; t += 4)
          *reinterpret_cast<float4 *>(orow + c0 + t) = make_float4(__uint_as_float(r[t]), __uint_as_float(r[t + 1]),
                                                                   __uint_as_float(r[t + 2]), __uint_as_float(r[t + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_slot, C::TMEM_COLS);
}

template <int HD>
int launch_bwd(const BwdMaps &mq, const BwdMaps &mk, int b, int h, int lq, int lk, float scale, const float *lse,
               const float *delta, float *dq, float *dk, float *dv, long long ld_dq, long long ld_dk, long long ld_dv,
               const unsigned long long *mask_q, const unsigned long long *mask_k, float dropout_p, unsigned int seed,
               const unsigned int *seed_dev, cudaStream_t s) {
  static bool configured = false;  // once per template instance
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         DqCfg<HD>::TOTAL + 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             DkvCfg<HD>::TOTAL + 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int bh = b * h;
  attn_bwd_dq_kernel<HD><<<dim3((lq + 127) / 128, bh), 384, DqCfg<HD>::TOTAL + 1024, s>>>(
      mq, lq, lk, b, h, scale, lse, delta, dq, ld_dq, mask_q, dropout_p, seed, seed_dev);
  attn_bwd_dkv_kernel<HD><<<dim3((lk + 127) / 128, bh), 384, DkvCfg<HD>::TOTAL + 1024, s>>>(
      mk, lq, lk, b, h, lse, delta, dk, dv, ld_dk, ld_dv, mask_k, dropout_p, seed, seed_dev);
  return launch_status();
}

}  // namespace

extern "C" {

long long coda_attention_bwd_workspace_bytes(int b, int h, int lq, int lk, int hd) {
  const long long bh = (long long)b * h;
  // q, dO rows; k, v rows (NS bf16 planes each) + delta
  return 2LL * NS * bh * hd * (2LL * lq + 2LL * lk) + 4LL * bh * lq + 4096;
}

int coda_attention_bwd(int b, int h, int lq, int lk, int hd, float scale, const float *q, const float *k,
                       const float *v, const float *out, const float *dout, const float *lse, float *dq,
                       float *dk, float *dv, float dropout_p, unsigned int seed, const unsigned int *seed_dev,
                       void *workspace, void *stream) {
  const long long e = (long long)h * hd;
  return coda_attention_bwd_ex(b, h, lq, lk, hd, scale, q, k, v, e, e, e, out, dout, lse, dq, dk, dv, e, e, e, nullptr,
                               nullptr, dropout_p, seed, seed_dev, workspace, stream);
}

int coda_attention_bwd_ex(int b, int h, int lq, int lk, int hd, float scale, const float *q, const float *k,
                          const float *v, long long ld_q, long long ld_k, long long ld_v, const float *out,
                          const float *dout, const float *lse, float *dq, float *dk, float *dv, long long ld_dq,
                          long long ld_dk, long long ld_dv, const unsigned long long *mask_q,
                          const unsigned long long *mask_k, float dropout_p, unsigned int seed,
                          const unsigned int *seed_dev, void *workspace, void *stream) {
  if ((hd != 64 && hd != 128) || b < 0 || h <= 0 || lq <= 0 || lk <= 0 || (long long)b * h > 65535) return CODA_EINVAL;
  if (b == 0) return CODA_OK;
  if (!q || !k || !v || !out || !dout || !lse || !dq || !dk || !dv || !workspace) return CODA_EINVAL;
  if (dropout_p < 0.f || dropout_p >= 1.f) return CODA_EINVAL;
  if ((mask_q == nullptr) != (mask_k == nullptr)) return CODA_EINVAL;
  {
    const long long e0 = (long long)h * hd;
    if (ld_q < e0 || ld_k < e0 || ld_v < e0 || ld_dq < e0 || ld_dk < e0 || ld_dv < e0) return CODA_EINVAL;
    if ((ld_q | ld_k | ld_v | ld_dq | ld_dk | ld_dv) & 3) return CODA_EINVAL;
    if ((((uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) != 0) return CODA_EINVAL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int bh = b * h;
  __nv_bfloat16 *w = (__nv_bfloat16 *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  __nv_bfloat16 *qp = w;                                   w += (size_t)NS * bh * lq * hd;
  __nv_bfloat16 *dop = w;                                  w += (size_t)NS * bh * lq * hd;
  __nv_bfloat16 *kp = w;                                   w += (size_t)NS * bh * lk * hd;
  __nv_bfloat16 *vp = w;                                   w += (size_t)NS * bh * lk * hd;
  float *delta = (float *)(((uintptr_t)w + 255) & ~(uintptr_t)255);
  if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)dout) & 15) != 0) return CODA_EINVAL;
  PackJobs jobs = {};
  const long long e = (long long)h * hd;
  jobs.job[0] = {q, qp, lq, scale, ld_q};
  jobs.job[1] = {dout, dop, lq, 1.0f, e};
  jobs.job[2] = {k, kp, lk, 1.0f, ld_k};
  jobs.job[3] = {v, vp, lk, 1.0f, ld_v};
  const long long t4 = (long long)(lq > lk ? lq : lk) * bh * hd / 4;
  pack_rows_multi_kernel<NS, false><<<dim3((unsigned)((t4 + 255) / 256), 4), 256, 0, s>>>(jobs, b, h, hd);
  bwd_delta_kernel<<<(unsigned)(((long long)lq * bh * 32 + 255) / 256), 256, 0, s>>>(lq, b, h, hd, dout, out, delta);
  int st = launch_status();
  if (st != CODA_OK) return st;

  BwdMaps mq, mk;  // dq kernel: 128-row q/dO boxes, 64-row k/v boxes; dkv kernel: the opposite
  for (int p = 0; p < NS; ++p) {
    const size_t oq = (size_t)p * bh * lq * hd, ok = (size_t)p * bh * lk * hd;
#define MAP(dst, base, rows, box) \
  if ((st = make_tmap_k_major_16b(&(dst), (base), 0, hd, (rows), bh, hd, (long long)(rows) * hd, (box))) != CODA_OK) return st
    MAP(mq.q[p], qp + oq, lq, 128);
    MAP(mq.dO[p], dop + oq, lq, 128);
    MAP(mq.k[p], kp + ok, lk, 64);
    MAP(mq.v[p], vp + ok, lk, 64);
    MAP(mk.k[p], kp + ok, lk, 128);
    MAP(mk.v[p], vp + ok, lk, 128);
    MAP(mk.q[p], qp + oq, lq, 64);
    MAP(mk.dO[p], dop + oq, lq, 64);
#undef MAP
  }
  if (hd == 64)
    return launch_bwd<64>(mq, mk, b, h, lq, lk, scale, lse, delta, dq, dk, dv, ld_dq, ld_dk, ld_dv, mask_q, mask_k,
                          dropout_p, seed, seed_dev, s);
  return launch_bwd<128>(mq, mk, b, h, lq, lk, scale, lse, delta, dq, dk, dv, ld_dq, ld_dk, ld_dv, mask_q, mask_k,
                         dropout_p, seed, seed_dev, s);
}

}  // extern "C"
