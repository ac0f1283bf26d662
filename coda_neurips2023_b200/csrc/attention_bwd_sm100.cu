// Fused multi-head attention BACKWARD for B200 (sm_100a), head dim 64 (the 3DETR encoder, where the
// (B*H, 2048, 2048) probability tensor dominates).  Two tcgen05 kernels, every MMA operand K-major so
// that no transposed accumulation or atomics are needed:
//
//   attn_bwd_dq_kernel   CTA = (bh, 128 queries); loop over 64-key tiles j
//        S_j  = Qs K_j^T              dP_j = dO V_j^T                    (TMEM 128 x 64 each)
//        dS_j = P_j o (m_j o dP_j - D)    with P_j = exp(S_j - LSE), m = dropout factor, D = rowsum(dO o O)
//        dQ  += dS_j K_j              (accumulated in TMEM over j; scaled by 1/sqrt(hd) at the end)
//   attn_bwd_dkv_kernel  CTA = (bh, 128 keys); loop over 64-query tiles i
//        S_i^T  = K Qs_i^T            dP_i^T = V dO_i^T                  (TMEM 128 x 64 each)
//        P~_i^T = P_i^T o m           dS_i^T = P_i^T o (m o dP_i^T - D_i)
//        dV += P~_i^T dO_i            dK += dS_i^T Qs_i                  (accumulated in TMEM over i)
//
// S is recomputed in both kernels (7 MMA groups instead of the 5 of a single fused pass) in exchange
// for K-major operands everywhere.  Operands are split-bf16 planes like the forward; the backward
// uses 2 planes (3 cross products, ~16 mantissa bits), enough for the 2e-3 gradient parity bar.
#include "../../include/coda_attention.h"
#include "attention_common.cuh"

using namespace coda;
using namespace coda::attn;

namespace {

constexpr int HD = 64;
constexpr int NS = 2;           // planes per operand in the backward
constexpr int NPROD = 3;

// rows pack: src (L, B, H*HD) -> planes [NS][B*H][L][HD]
__global__ void __launch_bounds__(256)
bwd_pack_rows_kernel(int L, int B, int H, float scale, const float *__restrict__ src,
                     __nv_bfloat16 *__restrict__ planes) {
  const long long total = (long long)L * B * H * HD;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int d = (int)(i % HD);
  long long t = i / HD;
  const int h = (int)(t % H); t /= H;
  const int b = (int)(t % B);
  const int l = (int)(t / B);
  split3<NS>(__ldg(src + i) * scale, planes + (((size_t)(b * H + h)) * L + l) * HD + d, (size_t)total);
}
// transposed pack: src (L, B, H*HD) -> planes [NS][B*H][HD][Lpad]
__global__ void __launch_bounds__(256)
bwd_pack_t_kernel(int L, int Lpad, int B, int H, float scale, const float *__restrict__ src,
                  __nv_bfloat16 *__restrict__ planes) {
  __shared__ float tile[32][33];
  const int bh = blockIdx.z, b = bh / H, h = bh % H;
  const int l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, d = d0 + tx;
    tile[i][tx] = (l < L && d < HD) ? __ldg(src + ((size_t)l * B + b) * H * HD + (size_t)h * HD + d) * scale : 0.f;
  }
  __syncthreads();
  const size_t plane = (size_t)B * H * HD * Lpad;
  for (int i = ty; i < 32; i += 8) {
    const int d = d0 + i, l = l0 + tx;
    if (d < HD && l < Lpad) split3<NS>(tile[tx][i], planes + ((size_t)bh * HD + d) * Lpad + l, plane);
  }
}
// D[bh][q] = sum_d dO * O   (both (Lq, B, H*HD))
__global__ void __launch_bounds__(256)
bwd_delta_kernel(int Lq, int B, int H, const float *__restrict__ dout, const float *__restrict__ out,
                 float *__restrict__ delta) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Lq * B * H) return;
  const int h = warp % H, b = (warp / H) % B, q = warp / (H * B);
  const size_t off = ((size_t)q * B + b) * H * HD + (size_t)h * HD;
  float s = __ldg(dout + off + lane) * __ldg(out + off + lane) + __ldg(dout + off + lane + 32) * __ldg(out + off + lane + 32);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) delta[(size_t)(b * H + h) * Lq + q] = s;
}

struct BwdMaps {
  CUtensorMap q[NS], k[NS], v[NS], dO[NS], kt[NS], qt[NS], dOt[NS];
};

// write one row (64 values) of a [128 x 64] K-major 128B-swizzled tile as NS bf16 planes
__device__ __forceinline__ void store_row_planes(unsigned char *tile_row0, int plane_bytes, int row, const float (&v)[64]) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint32_t w[NS][4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float r0 = v[ch * 8 + e], r1 = v[ch * 8 + e + 1];
#pragma unroll
      for (int pl = 0; pl < NS; ++pl) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(r0), h1 = __float2bfloat16_rn(r1);
        w[pl][e >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        r0 -= __bfloat162float(h0);
        r1 -= __bfloat162float(h1);
      }
    }
    const uint32_t off = (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) << 4);
#pragma unroll
    for (int pl = 0; pl < NS; ++pl)
      *reinterpret_cast<uint4 *>(tile_row0 + pl * plane_bytes + off) = make_uint4(w[pl][0], w[pl][1], w[pl][2], w[pl][3]);
  }
}

// ====================================================================== dQ
struct DqSmem {
  static constexpr int T128 = 128 * 128;  // [128 x 64] bf16 tile
  static constexpr int T64 = 64 * 128;    // [64 x 64] bf16 tile
  static constexpr int Q_OFF = 0;                   // Qs   [128 q x 64]   resident
  static constexpr int DO_OFF = Q_OFF + NS * T128;  // dO   [128 q x 64]   resident
  static constexpr int K_OFF = DO_OFF + NS * T128;  // K_j  [64 k x 64]
  static constexpr int V_OFF = K_OFF + NS * T64;    // V_j  [64 k x 64]
  static constexpr int KT_OFF = V_OFF + NS * T64;   // K_j^T [64 hd x 64 k]
  static constexpr int DS_OFF = KT_OFF + NS * T64;  // dS_j [128 q x 64 k]
  static constexpr int TOTAL = DS_OFF + NS * T128;
};

__global__ void __launch_bounds__(256, 1)
attn_bwd_dq_kernel(const __grid_constant__ BwdMaps maps, int Lq, int Lk, int B, int H, float scale,
                   const float *__restrict__ lse, const float *__restrict__ delta, float *__restrict__ dq,
                   float drop_p, uint32_t seed, const uint32_t *__restrict__ seed_dev) {
  using SM = DqSmem;
  if (seed_dev) seed += __ldg(seed_dev);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t q_full, kv_full, kv_empty, kt_full, kt_empty, s_full, ds_full, dq_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, bh = blockIdx.y;
  const int ntiles = (Lk + 63) / 64;

  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1); mbar_init(&kv_full, 1); mbar_init(&kv_empty, 1); mbar_init(&kt_full, 1);
    mbar_init(&kt_empty, 1); mbar_init(&s_full, 1); mbar_init(&ds_full, 128); mbar_init(&dq_done, 1);
    mbar_fence_init_cluster();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm_s = tmem_slot, tm_dp = tmem_slot + 64, tm_dq = tmem_slot + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full, (uint32_t)(2 * NS * SM::T128));
      for (int p = 0; p < NS; ++p) {
        tma_load_3d(smem + SM::Q_OFF + p * SM::T128, &maps.q[p], &q_full, 0, q0, bh);
        tma_load_3d(smem + SM::DO_OFF + p * SM::T128, &maps.dO[p], &q_full, 0, q0, bh);
      }
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(&kv_empty, ph ^ 1u);
        mbar_arrive_expect_tx(&kv_full, (uint32_t)(2 * NS * SM::T64));
        for (int p = 0; p < NS; ++p) {
          tma_load_3d(smem + SM::K_OFF + p * SM::T64, &maps.k[p], &kv_full, 0, j * 64, bh);
          tma_load_3d(smem + SM::V_OFF + p * SM::T64, &maps.v[p], &kv_full, 0, j * 64, bh);
        }
        mbar_wait(&kt_empty, ph ^ 1u);
        mbar_arrive_expect_tx(&kt_full, (uint32_t)(NS * SM::T64));
        for (int p = 0; p < NS; ++p) tma_load_3d(smem + SM::KT_OFF + p * SM::T64, &maps.kt[p], &kt_full, j * 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    // warp-uniform control flow, one elected lane issues (see elect_one_sync)
    constexpr uint32_t idesc = umma_idesc_f16(0, 128, 64);
    mbar_wait(&q_full, 0);
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      mbar_wait(&kv_full, ph);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int p = 0; p < NPROD; ++p) {
          const uint64_t aq = umma_smem_desc_k_sw128(smem + SM::Q_OFF + a_pa(NS, p) * SM::T128);
          const uint64_t bk = umma_smem_desc_k_sw128(smem + SM::K_OFF + a_pb(NS, p) * SM::T64);
          const uint64_t ad = umma_smem_desc_k_sw128(smem + SM::DO_OFF + a_pa(NS, p) * SM::T128);
          const uint64_t bv = umma_smem_desc_k_sw128(smem + SM::V_OFF + a_pb(NS, p) * SM::T64);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_f16(tm_s, umma_desc_advance(aq, kk * 32), umma_desc_advance(bk, kk * 32), idesc, (uint32_t)((p | kk) != 0));
            umma_f16(tm_dp, umma_desc_advance(ad, kk * 32), umma_desc_advance(bv, kk * 32), idesc, (uint32_t)((p | kk) != 0));
          }
        }
        umma_commit(&kv_empty);
        umma_commit(&s_full);
      }
      __syncwarp();
      mbar_wait(&ds_full, ph);
      mbar_wait(&kt_full, ph);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int p = 0; p < NPROD; ++p) {
          const uint64_t as = umma_smem_desc_k_sw128(smem + SM::DS_OFF + a_pa(NS, p) * SM::T128);
          const uint64_t bt = umma_smem_desc_k_sw128(smem + SM::KT_OFF + a_pb(NS, p) * SM::T64);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tm_dq, umma_desc_advance(as, kk * 32), umma_desc_advance(bt, kk * 32), idesc, (uint32_t)((j | p | kk) != 0));
        }
        umma_commit(&kt_empty);  // also: dS_j consumed (next dS write waits on s_full of j+1, issued after this)
        if (j == ntiles - 1) umma_commit(&dq_done);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int qq = warp - 4, row = qq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qq * 32) << 16;
    const int qrow = q0 + row;
    const bool valid_row = qrow < Lq;
    const float lse_r = valid_row ? __ldg(lse + (size_t)bh * Lq + qrow) : 0.f;
    const float d_r = valid_row ? __ldg(delta + (size_t)bh * Lq + qrow) : 0.f;
    const bool dropout = drop_p > 0.f;
    const uint32_t thresh32 = drop_thresh32(drop_p);
    const float keep_scale = dropout ? 1.0f / (1.0f - drop_p) : 1.0f;
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      mbar_wait(&s_full, ph);
      tc_fence_after();
      uint32_t sr[2][32], pr[2][32];
      tmem_ld_32x32(tm_s + lane_base, sr[0]);
      tmem_ld_32x32(tm_s + lane_base + 32, sr[1]);
      tmem_ld_32x32(tm_dp + lane_base, pr[0]);
      tmem_ld_32x32(tm_dp + lane_base + 32, pr[1]);
      tmem_ld_wait();
      // the previous tile's dS must have been consumed by its MMAs before it is overwritten
      if (j > 0) mbar_wait(&kt_empty, ph ^ 1u);
      float ds[64];
      const int kvalid = Lk - j * 64;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const float s = __uint_as_float(sr[c >> 5][c & 31]);
        float p = (c < kvalid && valid_row) ? exp2f((s - lse_r) * LOG2E) : 0.f;
        float dp = __uint_as_float(pr[c >> 5][c & 31]);
        if (dropout) dp = drop_keep(seed, (uint32_t)bh, (uint32_t)qrow, (uint32_t)(j * 64 + c), thresh32) ? dp * keep_scale : 0.f;
        ds[c] = p * (dp - d_r);
      }
      store_row_planes(smem + SM::DS_OFF, SM::T128, row, ds);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(&ds_full);
    }
    mbar_wait(&dq_done, 0);
    tc_fence_after();
    {
      const int b = bh / H, h = bh - b * H;
      float *orow = dq + ((size_t)qrow * B + b) * (size_t)(H * HD) + (size_t)h * HD;
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tm_dq + lane_base + c0, r);   // warp-collective: every lane executes it
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
  if (warp == 2) tmem_dealloc(tmem_slot, 256);
}

// ====================================================================== dK, dV
struct DkvSmem {
  static constexpr int T128 = 128 * 128;
  static constexpr int T64 = 64 * 128;
  static constexpr int K_OFF = 0;                    // K    [128 k x 64]  resident
  static constexpr int V_OFF = K_OFF + NS * T128;    // V    [128 k x 64]  resident
  static constexpr int Q_OFF = V_OFF + NS * T128;    // Qs_i [64 q x 64]
  static constexpr int DO_OFF = Q_OFF + NS * T64;    // dO_i [64 q x 64]
  static constexpr int QT_OFF = DO_OFF + NS * T64;   // Qs_i^T [64 hd x 64 q]
  static constexpr int DOT_OFF = QT_OFF + NS * T64;  // dO_i^T [64 hd x 64 q]
  static constexpr int P_OFF = DOT_OFF + NS * T64;   // P~_i^T [128 k x 64 q]
  static constexpr int DS_OFF = P_OFF + NS * T128;   // dS_i^T [128 k x 64 q]
  static constexpr int TOTAL = DS_OFF + NS * T128;
};

__global__ void __launch_bounds__(256, 1)
attn_bwd_dkv_kernel(const __grid_constant__ BwdMaps maps, int Lq, int Lk, int B, int H,
                    const float *__restrict__ lse, const float *__restrict__ delta, float *__restrict__ dk,
                    float *__restrict__ dv, float drop_p, uint32_t seed, const uint32_t *__restrict__ seed_dev) {
  using SM = DkvSmem;
  if (seed_dev) seed += __ldg(seed_dev);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t kv_full, q_full, q_empty, t_full, t_empty, s_full, p_full, acc_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_lse[2][64], s_delta[2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, bh = blockIdx.y;
  const int ntiles = (Lq + 63) / 64;

  if (warp == 1 && lane == 0) {
    mbar_init(&kv_full, 1); mbar_init(&q_full, 1); mbar_init(&q_empty, 1); mbar_init(&t_full, 1);
    mbar_init(&t_empty, 1); mbar_init(&s_full, 1); mbar_init(&p_full, 128); mbar_init(&acc_done, 1);
    mbar_fence_init_cluster();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm_s = tmem_slot, tm_dp = tmem_slot + 64, tm_dv = tmem_slot + 128, tm_dk = tmem_slot + 192;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&kv_full, (uint32_t)(2 * NS * SM::T128));
      for (int p = 0; p < NS; ++p) {
        tma_load_3d(smem + SM::K_OFF + p * SM::T128, &maps.k[p], &kv_full, 0, k0, bh);
        tma_load_3d(smem + SM::V_OFF + p * SM::T128, &maps.v[p], &kv_full, 0, k0, bh);
      }
      for (int i = 0; i < ntiles; ++i) {
        const uint32_t ph = (uint32_t)i & 1u;
        mbar_wait(&q_empty, ph ^ 1u);
        mbar_arrive_expect_tx(&q_full, (uint32_t)(2 * NS * SM::T64));
        for (int p = 0; p < NS; ++p) {
          tma_load_3d(smem + SM::Q_OFF + p * SM::T64, &maps.q[p], &q_full, 0, i * 64, bh);
          tma_load_3d(smem + SM::DO_OFF + p * SM::T64, &maps.dO[p], &q_full, 0, i * 64, bh);
        }
        mbar_wait(&t_empty, ph ^ 1u);
        mbar_arrive_expect_tx(&t_full, (uint32_t)(2 * NS * SM::T64));
        for (int p = 0; p < NS; ++p) {
          tma_load_3d(smem + SM::QT_OFF + p * SM::T64, &maps.qt[p], &t_full, i * 64, 0, bh);
          tma_load_3d(smem + SM::DOT_OFF + p * SM::T64, &maps.dOt[p], &t_full, i * 64, 0, bh);
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform control flow, one elected lane issues (see elect_one_sync)
    constexpr uint32_t idesc = umma_idesc_f16(0, 128, 64);
    mbar_wait(&kv_full, 0);
    for (int i = 0; i < ntiles; ++i) {
      const uint32_t ph = (uint32_t)i & 1u;
      mbar_wait(&q_full, ph);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int p = 0; p < NPROD; ++p) {
          const uint64_t ak = umma_smem_desc_k_sw128(smem + SM::K_OFF + a_pa(NS, p) * SM::T128);
          const uint64_t bq = umma_smem_desc_k_sw128(smem + SM::Q_OFF + a_pb(NS, p) * SM::T64);
          const uint64_t av = umma_smem_desc_k_sw128(smem + SM::V_OFF + a_pa(NS, p) * SM::T128);
          const uint64_t bd = umma_smem_desc_k_sw128(smem + SM::DO_OFF + a_pb(NS, p) * SM::T64);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_f16(tm_s, umma_desc_advance(ak, kk * 32), umma_desc_advance(bq, kk * 32), idesc, (uint32_t)((p | kk) != 0));
            umma_f16(tm_dp, umma_desc_advance(av, kk * 32), umma_desc_advance(bd, kk * 32), idesc, (uint32_t)((p | kk) != 0));
          }
        }
        umma_commit(&q_empty);
        umma_commit(&s_full);
      }
      __syncwarp();
      mbar_wait(&p_full, ph);
      mbar_wait(&t_full, ph);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int p = 0; p < NPROD; ++p) {
          const uint64_t ap = umma_smem_desc_k_sw128(smem + SM::P_OFF + a_pa(NS, p) * SM::T128);
          const uint64_t bo = umma_smem_desc_k_sw128(smem + SM::DOT_OFF + a_pb(NS, p) * SM::T64);
          const uint64_t as = umma_smem_desc_k_sw128(smem + SM::DS_OFF + a_pa(NS, p) * SM::T128);
          const uint64_t bq = umma_smem_desc_k_sw128(smem + SM::QT_OFF + a_pb(NS, p) * SM::T64);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            umma_f16(tm_dv, umma_desc_advance(ap, kk * 32), umma_desc_advance(bo, kk * 32), idesc, (uint32_t)((i | p | kk) != 0));
            umma_f16(tm_dk, umma_desc_advance(as, kk * 32), umma_desc_advance(bq, kk * 32), idesc, (uint32_t)((i | p | kk) != 0));
          }
        }
        umma_commit(&t_empty);  // Q^T/dO^T stage free, and P~^T / dS^T consumed
        if (i == ntiles - 1) umma_commit(&acc_done);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int qq = warp - 4, row = qq * 32 + lane;  // key row
    const uint32_t lane_base = (uint32_t)(qq * 32) << 16;
    const int krow = k0 + row;
    const bool valid_row = krow < Lk;
    const bool dropout = drop_p > 0.f;
    const uint32_t thresh32 = drop_thresh32(drop_p);
    const float keep_scale = dropout ? 1.0f / (1.0f - drop_p) : 1.0f;
    for (int i = 0; i < ntiles; ++i) {
      const uint32_t ph = (uint32_t)i & 1u;
      // log-sum-exp and D of this tile's 64 queries -> shared (double buffered; the named barrier
      // orders these writes before the reads below, and the reads of tile i before the rewrite at i+2)
      {
        const int t = row & 63, qi = i * 64 + t;
        float *dst = row < 64 ? s_lse[i & 1] : s_delta[i & 1];
        const float *src = row < 64 ? lse : delta;
        dst[t] = qi < Lq ? __ldg(src + (size_t)bh * Lq + qi) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&s_full, ph);
      tc_fence_after();
      uint32_t sr[2][32], pr[2][32];
      tmem_ld_32x32(tm_s + lane_base, sr[0]);
      tmem_ld_32x32(tm_s + lane_base + 32, sr[1]);
      tmem_ld_32x32(tm_dp + lane_base, pr[0]);
      tmem_ld_32x32(tm_dp + lane_base + 32, pr[1]);
      tmem_ld_wait();
      if (i > 0) mbar_wait(&t_empty, ph ^ 1u);  // previous P~^T / dS^T consumed by their MMAs
      float pt[64], ds[64];
      const int qvalid = Lq - i * 64;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const float s = __uint_as_float(sr[c >> 5][c & 31]);
        const float p = (c < qvalid && valid_row) ? exp2f((s - s_lse[i & 1][c]) * LOG2E) : 0.f;
        float dp = __uint_as_float(pr[c >> 5][c & 31]);
        float m = 1.0f;
        if (dropout) m = drop_keep(seed, (uint32_t)bh, (uint32_t)(i * 64 + c), (uint32_t)krow, thresh32) ? keep_scale : 0.f;
        pt[c] = p * m;
        ds[c] = p * (dp * m - s_delta[i & 1][c]);
      }
      store_row_planes(smem + SM::P_OFF, SM::T128, row, pt);
      store_row_planes(smem + SM::DS_OFF, SM::T128, row, ds);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      mbar_arrive(&p_full);
    }
    mbar_wait(&acc_done, 0);
    tc_fence_after();
    const int b = bh / H, h = bh - b * H;
    float *vrow = dv + ((size_t)krow * B + b) * (size_t)(H * HD) + (size_t)h * HD;
    float *krw = dk + ((size_t)krow * B + b) * (size_t)(H * HD) + (size_t)h * HD;
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t r[32], r2[32];
      tmem_ld_32x32(tm_dv + lane_base + c0, r);
      tmem_ld_32x32(tm_dk + lane_base + c0, r2);
      tmem_ld_wait();
      if (valid_row) {
#pragma unroll
        for (int t = 0; t < 32; t += 4) {
          *reinterpret_cast<float4 *>(vrow + c0 + t) = make_float4(__uint_as_float(r[t]), __uint_as_float(r[t + 1]),
                                                                   __uint_as_float(r[t + 2]), __uint_as_float(r[t + 3]));
          *reinterpret_cast<float4 *>(krw + c0 + t) = make_float4(__uint_as_float(r2[t]), __uint_as_float(r2[t + 1]),
                                                                  __uint_as_float(r2[t + 2]), __uint_as_float(r2[t + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_slot, 256);
}

}  // namespace

extern "C" {

long long coda_attention_bwd_workspace_bytes(int b, int h, int lq, int lk) {
  const long long lqp = (lq + 63) / 64 * 64, lkp = (lk + 63) / 64 * 64, bh = (long long)b * h;
  // q, dO rows; k, v rows; k^T; q^T, dO^T  (+ delta)
  return 2LL * NS * bh * HD * (2 * lq + 2 * lk + lkp + 2 * lqp) + 4LL * bh * lq + 4096;
}

int coda_attention_bwd(int b, int h, int lq, int lk, int hd, float scale, const float *q, const float *k,
                       const float *v, const float *out, const float *dout, const float *lse, float *dq,
                       float *dk, float *dv, float dropout_p, unsigned int seed, const unsigned int *seed_dev,
                       void *workspace, void *stream) {
  if (hd != HD || b < 0 || h <= 0 || lq <= 0 || lk <= 0 || (long long)b * h > 65535) return CODA_EINVAL;
  if (b == 0) return CODA_OK;
  if (!q || !k || !v || !out || !dout || !lse || !dq || !dk || !dv || !workspace) return CODA_EINVAL;
  if (dropout_p < 0.f || dropout_p >= 1.f) return CODA_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int bh = b * h;
  const int lqp = (lq + 63) / 64 * 64, lkp = (lk + 63) / 64 * 64;
  __nv_bfloat16 *w = (__nv_bfloat16 *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  __nv_bfloat16 *qp = w;                                   w += (size_t)NS * bh * lq * HD;
  __nv_bfloat16 *dop = w;                                  w += (size_t)NS * bh * lq * HD;
  __nv_bfloat16 *kp = w;                                   w += (size_t)NS * bh * lk * HD;
  __nv_bfloat16 *vp = w;                                   w += (size_t)NS * bh * lk * HD;
  __nv_bfloat16 *ktp = w;                                  w += (size_t)NS * bh * HD * lkp;
  __nv_bfloat16 *qtp = w;                                  w += (size_t)NS * bh * HD * lqp;
  __nv_bfloat16 *dotp = w;                                 w += (size_t)NS * bh * HD * lqp;
  float *delta = (float *)(((uintptr_t)w + 255) & ~(uintptr_t)255);
  const long long tq = (long long)lq * bh * HD, tk = (long long)lk * bh * HD;
  bwd_pack_rows_kernel<<<(unsigned)((tq + 255) / 256), 256, 0, s>>>(lq, b, h, scale, q, qp);
  bwd_pack_rows_kernel<<<(unsigned)((tq + 255) / 256), 256, 0, s>>>(lq, b, h, 1.0f, dout, dop);
  bwd_pack_rows_kernel<<<(unsigned)((tk + 255) / 256), 256, 0, s>>>(lk, b, h, 1.0f, k, kp);
  bwd_pack_rows_kernel<<<(unsigned)((tk + 255) / 256), 256, 0, s>>>(lk, b, h, 1.0f, v, vp);
  bwd_pack_t_kernel<<<dim3((lkp + 31) / 32, 2, bh), 256, 0, s>>>(lk, lkp, b, h, 1.0f, k, ktp);
  bwd_pack_t_kernel<<<dim3((lqp + 31) / 32, 2, bh), 256, 0, s>>>(lq, lqp, b, h, scale, q, qtp);
  bwd_pack_t_kernel<<<dim3((lqp + 31) / 32, 2, bh), 256, 0, s>>>(lq, lqp, b, h, 1.0f, dout, dotp);
  bwd_delta_kernel<<<(unsigned)(((long long)lq * bh * 32 + 255) / 256), 256, 0, s>>>(lq, b, h, dout, out, delta);
  int st = launch_status();
  if (st != CODA_OK) return st;

  BwdMaps mq, mk;  // dq kernel: 128-row q/dO boxes, 64-row k/v boxes; dkv kernel: the opposite
  for (int p = 0; p < NS; ++p) {
    const size_t oq = (size_t)p * bh * lq * HD, ok = (size_t)p * bh * lk * HD;
    const size_t okt = (size_t)p * bh * HD * lkp, oqt = (size_t)p * bh * HD * lqp;
#define MAP(dst, base, kdim, rows, rstride, bstride, box) \
  if ((st = make_tmap_k_major_16b(&(dst), (base), 0, (kdim), (rows), bh, (rstride), (bstride), (box))) != CODA_OK) return st
    MAP(mq.q[p], qp + oq, HD, lq, HD, (long long)lq * HD, 128);
    MAP(mq.dO[p], dop + oq, HD, lq, HD, (long long)lq * HD, 128);
    MAP(mq.k[p], kp + ok, HD, lk, HD, (long long)lk * HD, 64);
    MAP(mq.v[p], vp + ok, HD, lk, HD, (long long)lk * HD, 64);
    MAP(mq.kt[p], ktp + okt, lkp, HD, lkp, (long long)HD * lkp, 64);
    mq.qt[p] = mq.kt[p]; mq.dOt[p] = mq.kt[p];
    MAP(mk.k[p], kp + ok, HD, lk, HD, (long long)lk * HD, 128);
    MAP(mk.v[p], vp + ok, HD, lk, HD, (long long)lk * HD, 128);
    MAP(mk.q[p], qp + oq, HD, lq, HD, (long long)lq * HD, 64);
    MAP(mk.dO[p], dop + oq, HD, lq, HD, (long long)lq * HD, 64);
    MAP(mk.qt[p], qtp + oqt, lqp, HD, lqp, (long long)HD * lqp, 64);
    MAP(mk.dOt[p], dotp + oqt, lqp, HD, lqp, (long long)HD * lqp, 64);
    mk.kt[p] = mk.qt[p];
#undef MAP
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::TOTAL + 1024);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::TOTAL + 1024);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  attn_bwd_dq_kernel<<<dim3((lq + 127) / 128, bh), 256, DqSmem::TOTAL + 1024, s>>>(
      mq, lq, lk, b, h, scale, lse, delta, dq, dropout_p, seed, seed_dev);
  attn_bwd_dkv_kernel<<<dim3((lk + 127) / 128, bh), 256, DkvSmem::TOTAL + 1024, s>>>(
      mk, lq, lk, b, h, lse, delta, dk, dv, dropout_p, seed, seed_dev);
  return launch_status();
}

}  // extern "C"
