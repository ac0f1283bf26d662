// Fused multi-head attention forward for B200 (sm_100a): S = Q K^T and O = P V on
// tcgen05 tensor cores with TMEM accumulators, operands staged by TMA into
// 128B-swizzled shared memory, online softmax in registers.  The (Lq x Lk)
// probability matrix never leaves the SM.  C-ABI in include/coda_attention.h.
//
// Numerics: q (pre-scaled), k, v are fp32; each is split into NSPLIT bf16 planes
// (x = p0 + p1 (+ p2)) by attn_pack_kernel and the 1 / 3 / 6 significant cross
// products are accumulated in fp32 -- fp32-class results at bf16 tensor-core rate
// (same scheme as gemm_sm100.cu).  P (in [0, 1]) is split the same way.
//
// One CTA = one (batch*head, 128-query tile); loop over 64-key tiles:
//   warp 0 lane 0 : TMA producer (Q once; K_j, V^T_j per tile)
//   warp 1 lane 0 : MMA issuer   S_j = Q K_j^T  -> TMEM[0,64);  O_j = P_j V_j -> TMEM[64, 64+HD)
//   warp 2        : TMEM alloc / dealloc
//   warps 4..7    : softmax (thread = query row = TMEM lane): tcgen05.ld S_j, running max /
//                   sum, P_j planes -> swizzled smem, then o = o * alpha + O_j from TMEM.
#include <math.h>

#include "../../include/coda_attention.h"
#include "attention_common.cuh"

using namespace coda;
using namespace coda::attn;

namespace {

template <int NSPLIT>
__global__ void __launch_bounds__(256)
attn_pack_rows_kernel(int L, int B, int H, int HD, float scale, const float *__restrict__ src,
                      __nv_bfloat16 *__restrict__ planes) {
  const long long total = (long long)L * B * H * HD;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i enumerates src order: ((l * B + b) * H + h) * HD + d
  const int d = (int)(i % HD);
  long long t = i / HD;
  const int h = (int)(t % H); t /= H;
  const int b = (int)(t % B);
  const int l = (int)(t / B);
  const size_t o = (((size_t)(b * H + h)) * L + l) * HD + d;
  split3<NSPLIT>(__ldg(src + i) * scale, planes + o, (size_t)total);
}

template <int NSPLIT>
__global__ void __launch_bounds__(256)
attn_pack_vt_kernel(int L, int Lpad, int B, int H, int HD, const float *__restrict__ src,
                    __nv_bfloat16 *__restrict__ planes) {
  __shared__ float tile[32][33];
  const int bh = blockIdx.z, b = bh / H, h = bh % H;
  const int l0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int l = l0 + i, d = d0 + tx;
    tile[i][tx] = (l < L && d < HD) ? __ldg(src + ((size_t)l * B + b) * H * HD + (size_t)h * HD + d) : 0.f;
  }
  __syncthreads();
  const size_t plane = (size_t)B * H * HD * Lpad;
  for (int i = ty; i < 32; i += 8) {
    const int d = d0 + i, l = l0 + tx;
    if (d < HD && l < Lpad) split3<NSPLIT>(tile[tx][i], planes + ((size_t)bh * HD + d) * Lpad + l, plane);
  }
}

// materialised keep-mask * 1/(1-p) for the interim (cuBLAS) backward: mult[bh][q][k] in {0, 1/(1-p)}
__global__ void __launch_bounds__(256)
dropout_mult_kernel(long long total, int lq, int lk, uint32_t seed, const uint32_t *__restrict__ seed_dev,
                    uint32_t thresh24, float keep_scale, float *__restrict__ mult) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (seed_dev) seed += __ldg(seed_dev);
  const uint32_t k = (uint32_t)(i % lk);
  const long long t = i / lk;
  const uint32_t q = (uint32_t)(t % lq), bh = (uint32_t)(t / lq);
  mult[i] = drop_keep(seed, bh, q, k, thresh24) ? keep_scale : 0.f;
}

// ------------------------------------------------------------------ the kernel
struct AttnMaps {
  CUtensorMap q[3], k[3], v[3];
};

template <int HD, int NSPLIT>
struct AttnSmem {
  static constexpr int KB = HD / 64;                       // 64-wide k-blocks of the head dim
  static constexpr int Q_PLANE = QT * HD * 2;              // KB blocks of [128 x 64]
  static constexpr int K_PLANE = KT * HD * 2;              // KB blocks of [64 x 64]
  static constexpr int V_PLANE = HD * KT * 2;              // [HD x 64]
  static constexpr int P_PLANE = QT * KT * 2;              // [128 x 64]
  static constexpr bool P_ALIASES_K = (NSPLIT * (Q_PLANE + K_PLANE + V_PLANE + P_PLANE) > 220 * 1024);
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = NSPLIT * Q_PLANE;
  static constexpr int V_OFF = K_OFF + NSPLIT * (K_PLANE > P_PLANE || !P_ALIASES_K ? K_PLANE : P_PLANE);
  static constexpr int P_OFF = P_ALIASES_K ? K_OFF : V_OFF + NSPLIT * V_PLANE;
  static constexpr int TOTAL = (P_ALIASES_K ? V_OFF + NSPLIT * V_PLANE : P_OFF + NSPLIT * P_PLANE);
};

template <int HD, int NSPLIT>
__global__ void __launch_bounds__(256, 1)
attn_fwd_kernel(const __grid_constant__ AttnMaps maps, int Lq, int Lk, int B, int H, float *__restrict__ out,
                float *__restrict__ lse, float drop_p, uint32_t seed, const uint32_t *__restrict__ seed_dev) {
  using SM = AttnSmem<HD, NSPLIT>;
  if (seed_dev) seed += __ldg(seed_dev);  // per-step counter kept on the device (CUDA-graph friendly)
  constexpr int KB = SM::KB;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t q_full, k_full, k_empty, v_full, v_empty, s_full, p_full, o_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT, bh = blockIdx.y;
  const int ntiles = (Lk + KT - 1) / KT;
  constexpr uint32_t TMEM_COLS = (64 + HD) <= 128 ? 128 : 256;

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int p = 0; p < NSPLIT; ++p) { prefetch_tmap(&maps.q[p]); prefetch_tmap(&maps.k[p]); prefetch_tmap(&maps.v[p]); }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 1); mbar_init(&k_full, 1); mbar_init(&k_empty, 1); mbar_init(&v_full, 1);
    mbar_init(&v_empty, 1); mbar_init(&s_full, 1); mbar_init(&p_full, 128); mbar_init(&o_full, 1);
    mbar_fence_init_cluster();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot;       // S_j : columns [0, 64)
  const uint32_t tmem_o = tmem_slot + 64;  // O_j : columns [64, 64 + HD)

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_arrive_expect_tx(&q_full, (uint32_t)(NSPLIT * SM::Q_PLANE));
#pragma unroll
      for (int p = 0; p < NSPLIT; ++p)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
          tma_load_3d(smem + SM::Q_OFF + p * SM::Q_PLANE + kb * (QT * 128), &maps.q[p], &q_full, kb * 64, q0, bh);
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(&k_empty, ph ^ 1u);
        if (SM::P_ALIASES_K && j > 0) mbar_wait(&v_empty, ph ^ 1u);  // P_{j-1} (in K's buffer) consumed
        mbar_arrive_expect_tx(&k_full, (uint32_t)(NSPLIT * SM::K_PLANE));
#pragma unroll
        for (int p = 0; p < NSPLIT; ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(smem + SM::K_OFF + p * SM::K_PLANE + kb * (KT * 128), &maps.k[p], &k_full, kb * 64, j * KT, bh);
        mbar_wait(&v_empty, ph ^ 1u);
        mbar_arrive_expect_tx(&v_full, (uint32_t)(NSPLIT * SM::V_PLANE));
#pragma unroll
        for (int p = 0; p < NSPLIT; ++p)
          tma_load_3d(smem + SM::V_OFF + p * SM::V_PLANE, &maps.v[p], &v_full, j * KT, 0, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc_s = umma_idesc_f16(0, QT, KT);  // 128 x 64
      constexpr uint32_t idesc_o = umma_idesc_f16(0, QT, HD);  // 128 x HD
      mbar_wait(&q_full, 0);
      for (int j = 0; j < ntiles; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        mbar_wait(&k_full, ph);
        tc_fence_after();
#pragma unroll
        for (int p = 0; p < a_nprod(NSPLIT); ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t ad = umma_smem_desc_k_sw128(smem + SM::Q_OFF + a_pa(NSPLIT, p) * SM::Q_PLANE + kb * (QT * 128));
            const uint64_t bd = umma_smem_desc_k_sw128(smem + SM::K_OFF + a_pb(NSPLIT, p) * SM::K_PLANE + kb * (KT * 128));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(tmem_s, umma_desc_advance(ad, kk * 32), umma_desc_advance(bd, kk * 32), idesc_s,
                       (uint32_t)((p | kb | kk) != 0));
          }
        if (!SM::P_ALIASES_K) umma_commit(&k_empty);
        umma_commit(&s_full);
        mbar_wait(&p_full, ph);
        mbar_wait(&v_full, ph);
        tc_fence_after();
#pragma unroll
        for (int p = 0; p < a_nprod(NSPLIT); ++p) {
          const uint64_t ad = umma_smem_desc_k_sw128(smem + SM::P_OFF + a_pa(NSPLIT, p) * SM::P_PLANE);
          const uint64_t bd = umma_smem_desc_k_sw128(smem + SM::V_OFF + a_pb(NSPLIT, p) * SM::V_PLANE);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_o, umma_desc_advance(ad, kk * 32), umma_desc_advance(bd, kk * 32), idesc_o,
                     (uint32_t)((p | kk) != 0));
        }
        if (SM::P_ALIASES_K) umma_commit(&k_empty);
        umma_commit(&v_empty);
        umma_commit(&o_full);
      }
    }
  } else if (warp >= 4) {
    // ===== softmax / accumulation: one thread per query row =====
    const int q = warp - 4;
    const int row = q * 32 + lane;             // row in the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float o_acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o_acc[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const bool dropout = drop_p > 0.f;
    const uint32_t thresh24 = (uint32_t)(drop_p * 16777216.0f);
    const float keep_scale = dropout ? 1.0f / (1.0f - drop_p) : 1.0f;
    unsigned char *prow = smem + SM::P_OFF + row * 128;

    for (int j = 0; j < ntiles; ++j) {
      const uint32_t ph = (uint32_t)j & 1u;
      mbar_wait(&s_full, ph);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld_32x32(tmem_s + lane_base, sr[0]);
      tmem_ld_32x32(tmem_s + lane_base + 32, sr[1]);
      tmem_ld_wait();
      const int kvalid = Lk - j * KT;  // keys >= kvalid are padding
      float mloc = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        float s = __uint_as_float(sr[c >> 5][c & 31]);
        if (c >= kvalid) s = -INFINITY;
        sr[c >> 5][c & 31] = __float_as_uint(s);
        mloc = fmaxf(mloc, s);
      }
      const float m_new = fmaxf(m_run, mloc);
      const float alpha = exp2f((m_run - m_new) * LOG2E);  // m_run = -inf on the first tile -> 0
      float lsum = 0.f;
      // P_j -> NSPLIT bf16 planes, K-major 128B-swizzled rows of 64 keys
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t w[3][4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float pv[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int c = ch * 8 + e + t;
            float p = exp2f((__uint_as_float(sr[c >> 5][c & 31]) - m_new) * LOG2E);
            lsum += p;
            if (dropout) p = drop_keep(seed, (uint32_t)bh, (uint32_t)(q0 + row), (uint32_t)(j * KT + c), thresh24) ? p * keep_scale : 0.f;
            pv[t] = p;
          }
          float r0 = pv[0], r1 = pv[1];
#pragma unroll
          for (int pl = 0; pl < NSPLIT; ++pl) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(r0), h1 = __float2bfloat16_rn(r1);
            w[pl][e >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            r0 -= __bfloat162float(h0);
            r1 -= __bfloat162float(h1);
          }
        }
        const uint32_t off = (uint32_t)((ch ^ (row & 7)) << 4);
#pragma unroll
        for (int pl = 0; pl < NSPLIT; ++pl)
          *reinterpret_cast<uint4 *>(prow + pl * SM::P_PLANE + off) = make_uint4(w[pl][0], w[pl][1], w[pl][2], w[pl][3]);
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // P visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(&p_full);
      // O_j = P_j V_j, then o = o * alpha + O_j
      mbar_wait(&o_full, ph);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t orr[32];
        tmem_ld_32x32(tmem_o + lane_base + c0, orr);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) o_acc[c0 + t] = o_acc[c0 + t] * alpha + __uint_as_float(orr[t]);
      }
    }
    // ===== epilogue: normalise, store (Lq, B, H*HD) and the log-sum-exp =====
    const int qrow = q0 + row;
    if (qrow < Lq) {
      const float inv = 1.0f / l_run;
      const int b = bh / H, h = bh - b * H;
      float *orow = out + ((size_t)qrow * B + b) * (size_t)(H * HD) + (size_t)h * HD;
#pragma unroll
      for (int d = 0; d < HD; d += 4)
        *reinterpret_cast<float4 *>(orow + d) =
            make_float4(o_acc[d] * inv, o_acc[d + 1] * inv, o_acc[d + 2] * inv, o_acc[d + 3] * inv);
      if (lse) lse[(size_t)bh * Lq + qrow] = m_run + logf(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_slot, TMEM_COLS);
}

template <int HD, int NSPLIT>
int launch_attn(const AttnMaps &maps, int Lq, int Lk, int B, int H, float *out, float *lse, float drop_p,
                uint32_t seed, const uint32_t *seed_dev, cudaStream_t s) {
  constexpr size_t smem = AttnSmem<HD, NSPLIT>::TOTAL + 1024;
  auto kern = attn_fwd_kernel<HD, NSPLIT>;
  static bool configured = false;  // once per template instance
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const dim3 grid((Lq + QT - 1) / QT, B * H);
  kern<<<grid, 256, smem, s>>>(maps, Lq, Lk, B, H, out, lse, drop_p, seed, seed_dev);
  return launch_status();
}

}  // namespace

extern "C" {

long long coda_attention_workspace_bytes(int b, int h, int lq, int lk, int hd, int nsplit) {
  const long long lkpad = (lk + 63) / 64 * 64;
  return 2LL * nsplit * b * h * ((long long)lq * hd + (long long)lk * hd + hd * lkpad) + 1024;
}

static int attn_check(int b, int h, int lq, int lk, int hd, int nsplit) {
  if (b < 0 || h <= 0 || lq < 0 || lk <= 0 || (hd != 64 && hd != 128) || nsplit < 1 || nsplit > 3) return CODA_EINVAL;
  if ((long long)b * h > 65535) return CODA_EINVAL;
  return CODA_OK;
}

int coda_attention_pack(int b, int h, int lq, int lk, int hd, int nsplit, float scale, const float *q,
                        const float *k, const float *v, void *workspace, void *stream) {
  int st = attn_check(b, h, lq, lk, hd, nsplit);
  if (st != CODA_OK) return st;
  if (b == 0 || lq == 0) return CODA_OK;
  if (!q || !k || !v || !workspace) return CODA_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int bh = b * h;
  const int lkpad = (lk + 63) / 64 * 64;
  __nv_bfloat16 *qp = (__nv_bfloat16 *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  __nv_bfloat16 *kp = qp + (size_t)nsplit * bh * lq * hd;
  __nv_bfloat16 *vp = kp + (size_t)nsplit * bh * lk * hd;
  const long long tq = (long long)lq * bh * hd, tk = (long long)lk * bh * hd;
  const dim3 gvp((lkpad + 31) / 32, (hd + 31) / 32, bh);
#define CODA_PACK(NS)                                                                                          \
  attn_pack_rows_kernel<NS><<<(unsigned)((tq + 255) / 256), 256, 0, s>>>(lq, b, h, hd, scale, q, qp);           \
  attn_pack_rows_kernel<NS><<<(unsigned)((tk + 255) / 256), 256, 0, s>>>(lk, b, h, hd, 1.0f, k, kp);            \
  attn_pack_vt_kernel<NS><<<gvp, 256, 0, s>>>(lk, lkpad, b, h, hd, v, vp);
  if (nsplit == 1) { CODA_PACK(1) } else if (nsplit == 2) { CODA_PACK(2) } else { CODA_PACK(3) }
#undef CODA_PACK
  return launch_status();
}

int coda_attention_fwd_packed(int b, int h, int lq, int lk, int hd, int nsplit, const void *workspace,
                              float *out, float *lse, float dropout_p, unsigned int seed,
                              const unsigned int *seed_dev, void *stream) {
  int st = attn_check(b, h, lq, lk, hd, nsplit);
  if (st != CODA_OK) return st;
  if (b == 0 || lq == 0) return CODA_OK;
  if (!out || !workspace || dropout_p < 0.f || dropout_p >= 1.f) return CODA_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int bh = b * h;
  const int lkpad = (lk + 63) / 64 * 64;
  const __nv_bfloat16 *qp = (const __nv_bfloat16 *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  const __nv_bfloat16 *kp = qp + (size_t)nsplit * bh * lq * hd;
  const __nv_bfloat16 *vp = kp + (size_t)nsplit * bh * lk * hd;
  AttnMaps maps;
  for (int p = 0; p < nsplit; ++p) {
    st = make_tmap_k_major_16b(&maps.q[p], qp + (size_t)p * bh * lq * hd, 0, hd, lq, bh, hd, (long long)lq * hd, QT);
    if (st != CODA_OK) return st;
    st = make_tmap_k_major_16b(&maps.k[p], kp + (size_t)p * bh * lk * hd, 0, hd, lk, bh, hd, (long long)lk * hd, KT);
    if (st != CODA_OK) return st;
    st = make_tmap_k_major_16b(&maps.v[p], vp + (size_t)p * bh * hd * lkpad, 0, lkpad, hd, bh, lkpad,
                               (long long)hd * lkpad, hd);
    if (st != CODA_OK) return st;
  }
#define CODA_ATTN(HD_, NS) return launch_attn<HD_, NS>(maps, lq, lk, b, h, out, lse, dropout_p, seed, seed_dev, s)
  if (hd == 64) {
    if (nsplit == 1) CODA_ATTN(64, 1);
    if (nsplit == 2) CODA_ATTN(64, 2);
    CODA_ATTN(64, 3);
  }
  if (nsplit == 1) CODA_ATTN(128, 1);
  if (nsplit == 2) CODA_ATTN(128, 2);
  CODA_ATTN(128, 3);
#undef CODA_ATTN
}

int coda_attention_dropout_mult(int bh, int lq, int lk, float dropout_p, unsigned int seed,
                                const unsigned int *seed_dev, float *mult, void *stream) {
  if (bh < 0 || lq < 0 || lk < 0 || dropout_p < 0.f || dropout_p >= 1.f) return CODA_EINVAL;
  const long long total = (long long)bh * lq * lk;
  if (total == 0) return CODA_OK;
  if (!mult) return CODA_EINVAL;
  dropout_mult_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      total, lq, lk, seed, seed_dev, (uint32_t)(dropout_p * 16777216.0f), 1.0f / (1.0f - dropout_p), mult);
  return launch_status();
}

int coda_attention_fwd(int b, int h, int lq, int lk, int hd, int nsplit, float scale, const float *q,
                       const float *k, const float *v, float *out, float *lse, float dropout_p,
                       unsigned int seed, const unsigned int *seed_dev, void *workspace, void *stream) {
  int st = coda_attention_pack(b, h, lq, lk, hd, nsplit, scale, q, k, v, workspace, stream);
  if (st != CODA_OK) return st;
  return coda_attention_fwd_packed(b, h, lq, lk, hd, nsplit, workspace, out, lse, dropout_p, seed, seed_dev, stream);
}

}  // extern "C"
