// Fused multi-head attention forward for B200 (sm_100a): S = Q K^T and O = P V on
// tcgen05 tensor cores with TMEM accumulators, operands staged by TMA into
// 128B-swizzled shared memory, online softmax in registers.  The (Lq x Lk)
// probability matrix never leaves the SM.  C-ABI in include/coda_attention.h.
//
// Numerics: q (pre-scaled), k, v are fp32; each is split into NSPLIT bf16 planes
// (x = p0 + p1 (+ p2)) by attn_pack_kernel and the 1 / 3 / 6 significant cross
// products are accumulated in fp32 -- fp32-class results at bf16 tensor-core rate
// (same scheme as gemm_sm100.cu).  P (in [0, 1]) carries at most two planes.  q is
// packed with scale * log2(e), so the scores are in log2 units and the softmax is one
// ex2 per element; the log-sum-exp is returned in natural-log units.
//
// One CTA = one (batch*head, 128-query tile); loop over 64-key tiles:
//   warp 0 lane 0 : TMA producer of K_j;  warp 3 lane 0 : TMA producer of V_j (row-major tiles, consumed as
//                   MN-major B operands: no transposed copy of V exists);  4 / 2 stages for head dim 64 / 128
//   warp 1 lane 0 : MMA issuer   S_j = Q K_j^T -> TMEM[64 (j&1), +64) (double-buffered: S_{j+1} is
//                   queued before O_j);  O_j = P_j V_j -> TMEM[128 + HD (j % NWG), +HD)
//   warp 2        : TMEM alloc / dealloc
//   warps 4..     : softmax (thread = query row = TMEM lane): tcgen05.ld S_j, running max / sum,
//                   P_j planes -> swizzled smem, then o = o * alpha + O_j from TMEM.  Head dim 64
//                   runs two such warpgroups on alternate key tiles (one works on its exponentials
//                   while the tensor pipe serves the other) and merges their states at the end.
#include <math.h>

#include "../../include/coda_attention.h"
#include "attention_common.cuh"

using namespace coda;
using namespace coda::attn;

namespace {

// materialised keep-mask * 1/(1-p) for the interim (cuBLAS) backward: mult[bh][q][k] in {0, 1/(1-p)}
__global__ void __launch_bounds__(256)
dropout_mult_kernel(long long total, int lq, int lk, uint32_t seed, const uint32_t *__restrict__ seed_dev,
                    uint32_t thresh32, float keep_scale, float *__restrict__ mult) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (seed_dev) seed += __ldg(seed_dev);
  const uint32_t k = (uint32_t)(i % lk);
  const long long t = i / lk;
  const uint32_t q = (uint32_t)(t % lq), bh = (uint32_t)(t / lq);
  mult[i] = drop_keep(seed, bh, q, k, thresh32) ? keep_scale : 0.f;
}

// Attention masks travel bit-packed: bits[b][row][tile] (one 64-bit word per 64 columns), bit c set = column
// 64 * tile + c is NOT visible from that row (torch's boolean attn_mask convention).  The forward and the dQ kernel
// index rows by query, the dK/dV kernel by key (the transposed packing), so every softmax thread reads one word
// per tile.  rows / cols and the strides are in the orientation being packed.
__global__ void __launch_bounds__(128)
mask_pack_kernel(int B, int rows, int cols, const unsigned char *__restrict__ mask, long long stride_b,
                 long long stride_r, long long stride_c, unsigned long long *__restrict__ bits) {
  const int nt = (cols + 63) / 64;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * rows * nt) return;
  const int t = (int)(i % nt);
  const long long br = i / nt;
  const int r = (int)(br % rows), b = (int)(br / rows);
  const unsigned char *src = mask + (size_t)b * stride_b + (size_t)r * stride_r;
  unsigned long long w = 0;
  for (int c = 0; c < 64; ++c) {
    const int col = t * 64 + c;
    if (col < cols && __ldg(src + (size_t)col * stride_c)) w |= 1ull << c;
  }
  bits[i] = w;
}

// Radius mask of the reference's MaskedTransformerEncoder (models/transformer.py:155-162): point j is masked for
// point i when |x_i - x_j| >= radius (Euclidean, as torch.cdist).  Symmetric: one packing serves both orientations.
__global__ void __launch_bounds__(128)
mask_radius_kernel(int B, int L, const float *__restrict__ xyz, float radius, unsigned long long *__restrict__ bits) {
  const int nt = (L + 63) / 64;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * L * nt) return;
  const int t = (int)(i % nt);
  const long long br = i / nt;
  const int r = (int)(br % L), b = (int)(br / L);
  const float *base = xyz + (size_t)b * L * 3;
  const float x = __ldg(base + 3 * r), y = __ldg(base + 3 * r + 1), z = __ldg(base + 3 * r + 2);
  unsigned long long w = 0;
  for (int c = 0; c < 64; ++c) {
    const int col = t * 64 + c;
    if (col >= L) break;
    const float dx = x - __ldg(base + 3 * col), dy = y - __ldg(base + 3 * col + 1), dz = z - __ldg(base + 3 * col + 2);
    if (sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx))) >= radius) w |= 1ull << c;
  }
  bits[i] = w;
}

// half in -> ONE plane of half operands [b*h][L][hd] (scaled): the fp16 tower's q / k / v need a re-layout, not a split
__global__ void __launch_bounds__(256)
pack_rows_half_kernel(const __grid_constant__ PackJobs jobs, int B, int H, int hd) {
  const PackJob jb = jobs.job[blockIdx.y];
  const long long total4 = (long long)jb.L * B * H * hd / 4;
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 >= total4) return;
  const int hd4 = hd >> 2;
  const int d = (int)(i4 % hd4) * 4;
  long long t = i4 / hd4;
  const int h = (int)(t % H); t /= H;     // t = l * B + b
  const int b = (int)(t % B);
  const int l = (int)(t / B);
  const long long off = t * jb.ld + (long long)h * hd + d;
  uint2 raw = __ldg(reinterpret_cast<const uint2 *>(static_cast<const __half *>(jb.src) + off));
  if (jb.scale != 1.0f) {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
    const __half2 a = __floats2half2_rn(lo.x * jb.scale, lo.y * jb.scale), c = __floats2half2_rn(hi.x * jb.scale, hi.y * jb.scale);
    raw.x = *reinterpret_cast<const uint32_t *>(&a);
    raw.y = *reinterpret_cast<const uint32_t *>(&c);
  }
  __half *dst = reinterpret_cast<__half *>(jb.planes) + (((size_t)(b * H + h)) * jb.L + l) * hd + d;
  *reinterpret_cast<uint2 *>(dst) = raw;
}

// ------------------------------------------------------------------ the kernel
struct AttnMaps {
  CUtensorMap k[3], v[3];
};

// P lies in [0, 1] and is consumed once: two bf16 planes (relative error 2^-18) are enough even when
// q, k, v carry three, which drops one of the six cross products of O = P V and a third of the
// conversion work in the softmax warps.
__host__ __device__ constexpr int p_planes(int ns) { return ns == 3 ? 2 : ns; }
__host__ __device__ constexpr int pv_nprod(int ns) { return ns == 1 ? 1 : (ns == 2 ? 3 : 5); }
__host__ __device__ constexpr int pv_pa(int ns, int p) {  // plane of P, smallest terms first
  return ns == 1 ? 0 : ns == 2 ? (p == 0 ? 1 : 0) : (p == 0 ? 1 : p == 1 ? 0 : p == 2 ? 1 : 0);
}
__host__ __device__ constexpr int pv_pb(int ns, int p) {  // plane of V
  return ns == 1 ? 0 : ns == 2 ? (p == 1 ? 1 : 0) : (p == 0 ? 1 : p == 1 ? 2 : p == 2 ? 0 : p == 3 ? 1 : 0);
}

// The A operands of both contractions live in TMEM (tcgen05.mma with [a_tmem]): the Q planes are written
// there once per CTA, the P planes by the softmax warps every tile.  Only K and V^T (the B operands) are
// read from shared memory, which takes the 4 KB A-tile read per MMA off the shared-memory port -- with
// eleven plane products per key tile that port, not the tensor pipe, was the limiter.
// ONE_TILE: Lk <= 64 (the CLIP image tower: 50 tokens).  The whole problem is one key tile, so the CTA needs
// a single K / V stage, half the tensor memory and no running accumulator: two CTAs share an SM and hide each
// other's TMA / MMA / store latencies (one CTA per (crop, head) is otherwise all prologue).
template <int HD, int NSPLIT, bool ONE_TILE = false>
struct AttnCfg {
  // head dim 64: two softmax warpgroups take alternate key tiles (each with its own P / O buffers and its
  // own running max / sum, merged at the end).  head dim 128: one warpgroup (the O accumulator of a row
  // already fills its register budget).
  static constexpr int NWG = (HD == 64 && !ONE_TILE) ? 2 : 1;
  static constexpr int NST = ONE_TILE ? 1 : (HD == 64 ? 4 : 2);   // K / V stages
  static constexpr int NP = p_planes(NSPLIT);
  static constexpr int KB = HD / 64;                       // 64-wide k-blocks of the head dim
  static constexpr int K_PLANE = KT * HD * 2;              // KB blocks of [64 x 64]
  static constexpr int V_PLANE = KT * HD * 2;              // KB blocks of [64 keys x 64]: V_j row-major, used MN-major
  static constexpr int K_STAGE = NSPLIT * K_PLANE;
  static constexpr int V_STAGE = NSPLIT * V_PLANE;
  static constexpr int K_OFF = 0;
  static constexpr int V_OFF = NST * K_STAGE;
  static constexpr int TOTAL = NST * (K_STAGE + V_STAGE);
  static constexpr int THREADS = 128 + NWG * 128;
  // TMEM columns (32-bit): S double buffer | O per warpgroup | P planes per warpgroup | Q planes
  static constexpr int S_COL = 0;
  static constexpr int O_COL = ONE_TILE ? 64 : 128;
  static constexpr int TMEM_COLS = ONE_TILE ? 256 : 512;
  static constexpr int MIN_CTAS = ONE_TILE ? 2 : 1;
  static constexpr int P_COL = O_COL + NWG * HD;
  static constexpr int P_COLS = NP * (KT / 2);             // bf16 pairs: 32 columns per plane
  static constexpr int Q_COL = P_COL + NWG * P_COLS;
  static constexpr int Q_COLS = HD / 2;                    // per plane
  static_assert(Q_COL + NSPLIT * Q_COLS <= TMEM_COLS, "TMEM budget");
  static_assert(NWG == 1 || TOTAL >= QT * 64 * 4, "merge scratch must fit in the K/V stages");
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// F16: the operand planes (and P) are IEEE half instead of bf16 -- the fp16 CLIP tower needs no split at all: one
// plane, one MMA per product, where two bf16 planes took three.
template <int HD, int NSPLIT, bool ONE_TILE, bool F16 = false>
__global__ void __launch_bounds__(AttnCfg<HD, NSPLIT, ONE_TILE>::THREADS, AttnCfg<HD, NSPLIT, ONE_TILE>::MIN_CTAS)
attn_fwd_kernel(const __grid_constant__ AttnMaps maps, const __nv_bfloat16 *__restrict__ qplanes, int Lq, int Lk,
                int B, int H, float *__restrict__ out, float *__restrict__ lse, float drop_p, uint32_t seed,
                const uint32_t *__restrict__ seed_dev, int out_half, const unsigned long long *__restrict__ mask_q) {
  using SM = AttnCfg<HD, NSPLIT, ONE_TILE>;
  if (seed_dev) seed += __ldg(seed_dev);  // per-step counter kept on the device (CUDA-graph friendly)
  constexpr int KB = SM::KB, NWG = SM::NWG, NP = SM::NP, NST = SM::NST;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t q_full, k_full[NST], k_empty[NST], v_full[NST], v_empty[NST], s_full[2],
      s_free[2], p_full[NWG], o_full[NWG];
  __shared__ uint32_t tmem_slot;
  __shared__ float merge_ml[NWG == 2 ? QT : 1][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT, bh = blockIdx.y;
  const int ntiles = (Lk + KT - 1) / KT;
  constexpr uint32_t TMEM_COLS = SM::TMEM_COLS;

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int p = 0; p < NSPLIT; ++p) { prefetch_tmap(&maps.k[p]); prefetch_tmap(&maps.v[p]); }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full, 128);
#pragma unroll
    for (int i = 0; i < NST; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
#pragma unroll
    for (int i = 0; i < NWG; ++i) { mbar_init(&p_full[i], 128); mbar_init(&o_full[i], 1); }
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
    mbar_init(&s_free[0], 128); mbar_init(&s_free[1], 128);
    mbar_fence_init_cluster();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot + SM::S_COL;  // S_j : columns [64 (j&1), +64)
  const uint32_t tmem_o = tmem_slot + SM::O_COL;  // O_j : columns [HD (j % NWG), +HD)
  const uint32_t tmem_p = tmem_slot + SM::P_COL;  // P_j planes of warpgroup j % NWG (bf16 pairs)
  const uint32_t tmem_q = tmem_slot + SM::Q_COL;  // Q planes (bf16 pairs), written once

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer of K_j =====
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % NST;
        mbar_wait(&k_empty[st], ((uint32_t)(j / NST) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&k_full[st], (uint32_t)SM::K_STAGE);
#pragma unroll
        for (int p = 0; p < NSPLIT; ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(smem + SM::K_OFF + st * SM::K_STAGE + p * SM::K_PLANE + kb * (KT * 128), &maps.k[p],
                        &k_full[st], kb * 64, j * KT, bh);
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ===== TMA producer of V_j (own thread: a K load never queues behind a V stage still in use) =====
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % NST;
        mbar_wait(&v_empty[st], ((uint32_t)(j / NST) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&v_full[st], (uint32_t)SM::V_STAGE);
#pragma unroll
        for (int p = 0; p < NSPLIT; ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(smem + SM::V_OFF + st * SM::V_STAGE + p * SM::V_PLANE + kb * (KT * 128), &maps.v[p],
                        &v_full[st], kb * 64, j * KT, bh);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp in uniform control flow, one elected lane issues): the scores run two
    //       tiles ahead of O_j = P_j V_j so the tensor pipe works on the next scores while the softmax
    //       warps are busy with the current ones =====
    constexpr uint32_t idesc_s = umma_idesc_f16(F16 ? 1 : 0, QT, KT);  // 128 x 64
    constexpr uint32_t idesc_o = umma_idesc_f16(F16 ? 1 : 0, QT, HD, 0, 1);  // 128 x HD; B = V_j, MN-major (hd contiguous)
    auto issue_s = [&](int j) {
      const int st = j % NST;
      mbar_wait(&k_full[st], (uint32_t)(j / NST) & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t d = tmem_s + (uint32_t)(j & 1) * 64u;
#pragma unroll
        for (int p = 0; p < a_nprod(NSPLIT); ++p)
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t a = tmem_q + (uint32_t)(a_pa(NSPLIT, p) * SM::Q_COLS + kb * 32);
            const uint64_t bd = umma_smem_desc_k_sw128(smem + SM::K_OFF + st * SM::K_STAGE +
                                                       a_pb(NSPLIT, p) * SM::K_PLANE + kb * (KT * 128));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)  // 16 head-dim elements = 8 TMEM columns per MMA
              umma_f16_ts(d, a + kk * 8, umma_desc_advance(bd, kk * 32), idesc_s, (uint32_t)((p | kb | kk) != 0));
          }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    mbar_wait(&q_full, 0);
    tc_fence_after();
    issue_s(0);
    if (ntiles > 1) issue_s(1);
    for (int j = 0; j < ntiles; ++j) {
      const int st = j % NST, g = j % NWG;
      // scores two tiles ahead: their TMEM buffer is free once the softmax warps hold S_j in registers
      // (s_free, long before p_full), so S_{j+2} is ready when its warpgroup finishes tile j
      if (j + 2 < ntiles) {
        mbar_wait(&s_free[j & 1], (uint32_t)(j >> 1) & 1u);
        issue_s(j + 2);
      }
      mbar_wait(&p_full[g], (uint32_t)(j / NWG) & 1u);
      mbar_wait(&v_full[st], (uint32_t)(j / NST) & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t d = tmem_o + (uint32_t)g * HD;
#pragma unroll
        for (int p = 0; p < pv_nprod(NSPLIT); ++p) {
          const uint32_t a = tmem_p + (uint32_t)(g * SM::P_COLS + pv_pa(NSPLIT, p) * (KT / 2));
          const uint64_t bd = umma_smem_desc_mn_sw128(smem + SM::V_OFF + st * SM::V_STAGE + pv_pb(NSPLIT, p) * SM::V_PLANE);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // 16 keys = 8 TMEM columns of P, 16 rows x 128 B of V_j
            umma_f16_ts(d, a + kk * 8, umma_desc_advance(bd, kk * 16 * 128), idesc_o, (uint32_t)((p | kk) != 0));
        }
        umma_commit(&v_empty[st]);
        umma_commit(&o_full[g]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== softmax / accumulation: one thread per query row; warpgroup g owns tiles j = g (mod NWG) =====
    const int g = (warp - 4) >> 2;
    const int q = (warp - 4) & 3;
    const int row = q * 32 + lane;             // row in the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    if (g == 0) {
      // Q planes of this row -> TMEM (A operand of every S_j): bf16 pairs, 16 elements per 8 columns
      const bool valid = q0 + row < Lq;
#pragma unroll
      for (int p = 0; p < NSPLIT; ++p) {
        const uint4 *src = reinterpret_cast<const uint4 *>(
            qplanes + (((size_t)p * gridDim.y + bh) * Lq + (valid ? q0 + row : 0)) * HD);
#pragma unroll
        for (int c = 0; c < HD / 16; ++c) {
          const uint4 lo = valid ? __ldg(src + 2 * c) : make_uint4(0, 0, 0, 0);
          const uint4 hi = valid ? __ldg(src + 2 * c + 1) : make_uint4(0, 0, 0, 0);
          const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          tmem_st_32x8(tmem_q + lane_base + (uint32_t)(p * SM::Q_COLS + c * 8), w);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&q_full);
    }
    float o_acc[ONE_TILE ? 1 : HD];   // running output (not needed when the single tile IS the output)
#pragma unroll
    for (int d = 0; d < (ONE_TILE ? 1 : HD); ++d) o_acc[d] = 0.f;
    // scores arrive in log2 units (q was packed with scale * log2 e): p = 2^(s - m)
    float m_run = -INFINITY, l_run = 0.f;
    const bool dropout = drop_p > 0.f;
    const uint32_t thresh32 = drop_thresh32(drop_p);
    const float keep_scale = dropout ? 1.0f / (1.0f - drop_p) : 1.0f;
    constexpr LcgJump jump = make_lcg_jump();  // four interleaved chains: x_{c+4} = x_c A^4 + C_4
    const uint32_t lcg_a4 = jump.a[3], lcg_c4 = jump.c[3];
    const uint32_t my_o = tmem_o + (uint32_t)g * HD + lane_base;
    const uint32_t my_p = tmem_p + (uint32_t)(g * SM::P_COLS) + lane_base;
    // attention mask: one 64-bit word per (batch, query row, key tile), bit c set = key c of the tile is not visible
    const unsigned long long *mrow =
        mask_q ? mask_q + ((size_t)(bh / H) * Lq + (q0 + row < Lq ? q0 + row : 0)) * (size_t)ntiles : nullptr;

    for (int j = g; j < ntiles; j += NWG) {
      const uint32_t ph = (uint32_t)(j / NWG) & 1u;
      mbar_wait(&s_full[j & 1], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld_32x32(tmem_s + (uint32_t)(j & 1) * 64u + lane_base, sr[0]);
      tmem_ld_32x32(tmem_s + (uint32_t)(j & 1) * 64u + lane_base + 32, sr[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[j & 1]);     // S_j is in registers: its TMEM buffer may take S_{j+2}
      const int kvalid = Lk - j * KT;  // keys >= kvalid are padding (last tile only)
      if (kvalid < KT) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c >= kvalid) sr[c >> 5][c & 31] = __float_as_uint(-INFINITY);
      }
      if (mrow) {
        const unsigned long long mb = __ldg(mrow + j);
        const uint32_t mlo = (uint32_t)mb, mhi = (uint32_t)(mb >> 32);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if ((mlo >> c) & 1u) sr[0][c] = __float_as_uint(-INFINITY);
          if ((mhi >> c) & 1u) sr[1][c] = __float_as_uint(-INFINITY);
        }
      }
      float mloc = __uint_as_float(sr[0][0]);
#pragma unroll
      for (int c = 1; c < 64; ++c) mloc = fmaxf(mloc, __uint_as_float(sr[c >> 5][c & 31]));
      const float m_run_new = fmaxf(m_run, mloc);
      // a row whose keys so far are all masked keeps m = -inf; exponentials are then taken against 0 (all zero)
      const float m_new = m_run_new == -INFINITY ? 0.f : m_run_new;
      const float alpha = ex2_approx(m_run - m_new);  // m_run = -inf on the first tile -> 0
      float lsum = 0.f;
      uint32_t dx[4] = {0, 0, 0, 0};
      if (dropout) {
        const uint32_t ts = drop_tile_seed(seed, (uint32_t)bh, (uint32_t)(q0 + row), (uint32_t)j);
#pragma unroll
        for (int i = 0; i < 4; ++i) dx[i] = ts * jump.a[i] + jump.c[i];
      }
      // P_j -> NP bf16 planes in TMEM (A operand of O_j): 16 keys = 8 columns of bf16 pairs per store
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t w[NP][8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const int c = ch * 16 + e;
          float r0 = ex2_approx(__uint_as_float(sr[c >> 5][c & 31]) - m_new);
          float r1 = ex2_approx(__uint_as_float(sr[(c + 1) >> 5][(c + 1) & 31]) - m_new);
          lsum += r0 + r1;
          if (dropout) {
            r0 = dx[c & 3] >= thresh32 ? r0 : 0.f;   // the 1/(1-p) factor is applied once, at the end
            r1 = dx[(c + 1) & 3] >= thresh32 ? r1 : 0.f;
            dx[c & 3] = dx[c & 3] * lcg_a4 + lcg_c4;
            dx[(c + 1) & 3] = dx[(c + 1) & 3] * lcg_a4 + lcg_c4;
          }
#pragma unroll
          for (int pl = 0; pl < NP; ++pl) {
            uint32_t bits;
            if constexpr (F16) {
              const __half2 h2 = __floats2half2_rn(r0, r1);
              bits = *reinterpret_cast<const uint32_t *>(&h2);
            } else {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(r0, r1);  // one packed conversion
              bits = *reinterpret_cast<const uint32_t *>(&h2);
            }
            w[pl][e >> 1] = bits;
            if (pl + 1 < NP) {
              r0 -= __uint_as_float(bits << 16);
              r1 -= __uint_as_float(bits & 0xFFFF0000u);
            }
          }
        }
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) tmem_st_32x8(my_p + (uint32_t)(pl * (KT / 2) + ch * 8), w[pl]);
      }
      l_run = l_run * alpha + lsum;
      m_run = m_run_new;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      // O_j = P_j V_j, then o = o * alpha + O_j
      mbar_wait(&o_full[g], ph);
      tc_fence_after();
      if constexpr (ONE_TILE) {
        // the only tile: normalise and store straight from tensor memory
        const int qrow = q0 + row;
        const float inv = keep_scale / l_run;
        const int b = bh / H, h = bh - b * H;
        const size_t ooff = ((size_t)qrow * B + b) * (size_t)(H * HD) + (size_t)h * HD;
        float *orow = out + ooff;
        __half *hrow = reinterpret_cast<__half *>(out) + ooff;    // out_half: the fp16 tower takes its dtype back directly
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 32) {
          uint32_t orr[32];
          tmem_ld_32x32(my_o + c0, orr);
          tmem_ld_wait();
          if (qrow < Lq) {
            if (out_half) {
#pragma unroll
              for (int t = 0; t < 32; t += 4) {
                const __half2 lo = __floats2half2_rn(__uint_as_float(orr[t]) * inv, __uint_as_float(orr[t + 1]) * inv);
                const __half2 hi = __floats2half2_rn(__uint_as_float(orr[t + 2]) * inv, __uint_as_float(orr[t + 3]) * inv);
                uint2 w;
                w.x = *reinterpret_cast<const uint32_t *>(&lo);
                w.y = *reinterpret_cast<const uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(hrow + c0 + t) = w;
              }
            } else {
#pragma unroll
              for (int t = 0; t < 32; t += 4)
                *reinterpret_cast<float4 *>(orow + c0 + t) =
                    make_float4(__uint_as_float(orr[t]) * inv, __uint_as_float(orr[t + 1]) * inv,
                                __uint_as_float(orr[t + 2]) * inv, __uint_as_float(orr[t + 3]) * inv);
            }
          }
        }
        if (lse && qrow < Lq) lse[(size_t)bh * Lq + qrow] = m_run * LN2 + logf(l_run);
      } else {
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 32) {
          uint32_t orr[32];
          tmem_ld_32x32(my_o + c0, orr);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) o_acc[c0 + t] = o_acc[c0 + t] * alpha + __uint_as_float(orr[t]);
        }
      }
    }
    if constexpr (NWG == 2) {
      // ===== merge the two warpgroups' partial softmax states (disjoint key subsets) =====
      // every tile's o_full has been awaited by its owner, so after this barrier no MMA reads smem any more
      float *scratch = reinterpret_cast<float *>(smem);  // [HD][128] floats, row fastest
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (g == 1) {
#pragma unroll
        for (int d = 0; d < HD; ++d) scratch[d * QT + row] = o_acc[d];
        merge_ml[row][0] = m_run;
        merge_ml[row][1] = l_run;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (g == 0) {
        const float m1 = merge_ml[row][0], l1 = merge_ml[row][1];
        const float mm = fmaxf(m_run, m1);           // -inf only if every key of the row is masked
        const float m = mm == -INFINITY ? 0.f : mm;
        const float a0 = ex2_approx(m_run - m), a1 = ex2_approx(m1 - m);
#pragma unroll
        for (int d = 0; d < HD; ++d) o_acc[d] = o_acc[d] * a0 + scratch[d * QT + row] * a1;
        l_run = l_run * a0 + l1 * a1;
        m_run = mm;
      }
    }
    // ===== epilogue: normalise, store (Lq, B, H*HD) and the log-sum-exp (natural-log units) =====
    const int qrow = q0 + row;
    if (!ONE_TILE && g == 0 && qrow < Lq) {
      const float inv = keep_scale / l_run;
      const int b = bh / H, h = bh - b * H;
      float *orow = out + ((size_t)qrow * B + b) * (size_t)(H * HD) + (size_t)h * HD;
#pragma unroll
      for (int d = 0; d < HD; d += 4)
        *reinterpret_cast<float4 *>(orow + d) =
            make_float4(o_acc[d] * inv, o_acc[d + 1] * inv, o_acc[d + 2] * inv, o_acc[d + 3] * inv);
      if (lse) lse[(size_t)bh * Lq + qrow] = m_run * LN2 + logf(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_slot, TMEM_COLS);
}

template <int HD, int NSPLIT, bool ONE_TILE = false, bool F16 = false>
int launch_attn(const AttnMaps &maps, const __nv_bfloat16 *qplanes, int Lq, int Lk, int B, int H, float *out,
                float *lse, float drop_p, uint32_t seed, const uint32_t *seed_dev, cudaStream_t s, int out_half = 0,
                const unsigned long long *mask_q = nullptr) {
  if (out_half && !ONE_TILE) return CODA_EINVAL;   // fp16 output exists on the single-tile (CLIP tower) instance
  static_assert(!F16 || NSPLIT == 1, "half operands are a single plane");
  using SM = AttnCfg<HD, NSPLIT, ONE_TILE>;
  constexpr size_t smem = SM::TOTAL + 1024;
  auto kern = attn_fwd_kernel<HD, NSPLIT, ONE_TILE, F16>;
  static bool configured = false;  // once per template instance
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const dim3 grid((Lq + QT - 1) / QT, B * H);
  kern<<<grid, SM::THREADS, smem, s>>>(maps, qplanes, Lq, Lk, B, H, out, lse, drop_p, seed, seed_dev, out_half, mask_q);
  return launch_status();
}

}  // namespace

extern "C" {

long long coda_attention_workspace_bytes(int b, int h, int lq, int lk, int hd, int nsplit) {
  return 2LL * nsplit * b * h * ((long long)lq * hd + 2LL * lk * hd) + 1024;
}

static int attn_check(int b, int h, int lq, int lk, int hd, int nsplit) {
  if (b < 0 || h <= 0 || lq < 0 || lk <= 0 || (hd != 64 && hd != 128) || nsplit < 1 || nsplit > 3) return CODA_EINVAL;
  if ((long long)b * h > 65535) return CODA_EINVAL;
  return CODA_OK;
}

int coda_attention_pack_strided(int b, int h, int lq, int lk, int hd, int nsplit, float scale, const void *q,
                                const void *k, const void *v, long long ld_q, long long ld_k, long long ld_v,
                                int is_half, void *workspace, void *stream) {
  int st = attn_check(b, h, lq, lk, hd, nsplit);
  if (st != CODA_OK) return st;
  if (b == 0 || lq == 0) return CODA_OK;
  if (!q || !k || !v || !workspace) return CODA_EINVAL;
  const uintptr_t amask = is_half ? 7 : 15;    // four elements per load
  if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & amask) != 0 || ((ld_q | ld_k | ld_v) & 3) != 0) return CODA_EINVAL;
  if (ld_q < (long long)h * hd || ld_k < (long long)h * hd || ld_v < (long long)h * hd) return CODA_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int bh = b * h;
  __nv_bfloat16 *qp = (__nv_bfloat16 *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  __nv_bfloat16 *kp = qp + (size_t)nsplit * bh * lq * hd;
  __nv_bfloat16 *vp = kp + (size_t)nsplit * bh * lk * hd;
  // q carries scale * log2(e): the kernel's softmax works in base 2.  All three tensors in one launch.
  PackJobs jobs = {};
  jobs.job[0] = {q, qp, lq, scale * LOG2E, ld_q};
  jobs.job[1] = {k, kp, lk, 1.0f, ld_k};
  jobs.job[2] = {v, vp, lk, 1.0f, ld_v};
  const long long t4 = (long long)(lq > lk ? lq : lk) * bh * hd / 4;
  const dim3 grid((unsigned)((t4 + 255) / 256), 3);
#define CODA_PACK(NS)                                                                 \
  if (is_half) pack_rows_multi_kernel<NS, true><<<grid, 256, 0, s>>>(jobs, b, h, hd); \
  else pack_rows_multi_kernel<NS, false><<<grid, 256, 0, s>>>(jobs, b, h, hd)
  if (nsplit == 1) { CODA_PACK(1); } else if (nsplit == 2) { CODA_PACK(2); } else { CODA_PACK(3); }
#undef CODA_PACK
  return launch_status();
}

int coda_attention_pack(int b, int h, int lq, int lk, int hd, int nsplit, float scale, const float *q,
                        const float *k, const float *v, void *workspace, void *stream) {
  const long long e = (long long)h * hd;
  return coda_attention_pack_strided(b, h, lq, lk, hd, nsplit, scale, q, k, v, e, e, e, 0, workspace, stream);
}

int coda_attention_fwd_packed(int b, int h, int lq, int lk, int hd, int nsplit, const void *workspace,
                              float *out, float *lse, float dropout_p, unsigned int seed,
                              const unsigned int *seed_dev, void *stream) {
  return coda_attention_fwd_packed_ex(b, h, lq, lk, hd, nsplit, workspace, out, 0, lse, dropout_p, seed, seed_dev,
                                      stream);
}

int coda_attention_fwd_packed_ex(int b, int h, int lq, int lk, int hd, int nsplit, const void *workspace,
                                 void *out_v, int out_half, float *lse, float dropout_p, unsigned int seed,
                                 const unsigned int *seed_dev, void *stream) {
  return coda_attention_fwd_packed_masked(b, h, lq, lk, hd, nsplit, workspace, out_v, out_half, lse, nullptr, dropout_p,
                                          seed, seed_dev, stream);
}

int coda_attention_fwd_packed_masked(int b, int h, int lq, int lk, int hd, int nsplit, const void *workspace,
                                     void *out_v, int out_half, float *lse, const unsigned long long *mask_q,
                                     float dropout_p, unsigned int seed, const unsigned int *seed_dev, void *stream) {
  float *out = reinterpret_cast<float *>(out_v);
  int st = attn_check(b, h, lq, lk, hd, nsplit);
  if (st != CODA_OK) return st;
  if (b == 0 || lq == 0) return CODA_OK;
  if (!out || !workspace || dropout_p < 0.f || dropout_p >= 1.f) return CODA_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int bh = b * h;
  const __nv_bfloat16 *qp = (const __nv_bfloat16 *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  const __nv_bfloat16 *kp = qp + (size_t)nsplit * bh * lq * hd;
  const __nv_bfloat16 *vp = kp + (size_t)nsplit * bh * lk * hd;
  AttnMaps maps;
  for (int p = 0; p < nsplit; ++p) {
    st = make_tmap_k_major_16b(&maps.k[p], kp + (size_t)p * bh * lk * hd, 0, hd, lk, bh, hd, (long long)lk * hd, KT);
    if (st != CODA_OK) return st;
    st = make_tmap_k_major_16b(&maps.v[p], vp + (size_t)p * bh * lk * hd, 0, hd, lk, bh, hd, (long long)lk * hd, KT);
    if (st != CODA_OK) return st;
  }
#define CODA_ATTN(HD_, NS) \
  return launch_attn<HD_, NS>(maps, qp, lq, lk, b, h, out, lse, dropout_p, seed, seed_dev, s, 0, mask_q)
  if (hd == 64 && lk <= KT && nsplit <= 2 && !mask_q) {   // single key tile (CLIP image tower): two CTAs per SM
    if (nsplit == 1) return launch_attn<64, 1, true>(maps, qp, lq, lk, b, h, out, lse, dropout_p, seed, seed_dev, s, out_half);
    return launch_attn<64, 2, true>(maps, qp, lq, lk, b, h, out, lse, dropout_p, seed, seed_dev, s, out_half);
  }
  if (out_half) return CODA_EINVAL;
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

int coda_attention_fwd_half(int b, int h, int l, int hd, const void *q, const void *k, const void *v, long long ld_q,
                            long long ld_k, long long ld_v, void *out, void *workspace, void *stream) {
  // fp16 self-attention of at most one key tile (the CLIP image tower: 50 tokens, 12 x 64): q / k / v are read as
  // half (row-strided slices of the fused projection), re-laid as ONE plane of half operands -- fp16 is the tensor
  // core's native type, no bf16 split -- and the output is written as half.
  if (b < 0 || h <= 0 || l <= 0 || l > KT || hd != 64 || (long long)b * h > 65535) return CODA_EINVAL;
  if (b == 0) return CODA_OK;
  if (!q || !k || !v || !out || !workspace) return CODA_EINVAL;
  if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 7) != 0 || ((ld_q | ld_k | ld_v) & 3) != 0) return CODA_EINVAL;
  if (ld_q < (long long)h * hd || ld_k < (long long)h * hd || ld_v < (long long)h * hd) return CODA_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int bh = b * h;
  __nv_bfloat16 *qp = (__nv_bfloat16 *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  __nv_bfloat16 *kp = qp + (size_t)bh * l * hd;
  __nv_bfloat16 *vp = kp + (size_t)bh * l * hd;
  PackJobs jobs = {};
  const float scale = 1.0f / sqrtf((float)hd);
  jobs.job[0] = {q, qp, l, scale * LOG2E, ld_q};
  jobs.job[1] = {k, kp, l, 1.0f, ld_k};
  jobs.job[2] = {v, vp, l, 1.0f, ld_v};
  const long long t4 = (long long)l * bh * hd / 4;
  pack_rows_half_kernel<<<dim3((unsigned)((t4 + 255) / 256), 3), 256, 0, s>>>(jobs, b, h, hd);
  int st = launch_status();
  if (st != CODA_OK) return st;
  AttnMaps maps;
  st = make_tmap_k_major_16b(&maps.k[0], kp, 1, hd, l, bh, hd, (long long)l * hd, KT);
  if (st != CODA_OK) return st;
  st = make_tmap_k_major_16b(&maps.v[0], vp, 1, hd, l, bh, hd, (long long)l * hd, KT);
  if (st != CODA_OK) return st;
  for (int p = 1; p < 3; ++p) { maps.k[p] = maps.k[0]; maps.v[p] = maps.v[0]; }
  return launch_attn<64, 1, true, true>(maps, qp, l, l, b, h, reinterpret_cast<float *>(out), nullptr, 0.f, 0u, nullptr,
                                        s, 1);
}

int coda_attention_mask_pack(int b, int lq, int lk, const unsigned char *mask, long long stride_b, long long stride_q,
                             long long stride_k, unsigned long long *bits_q, unsigned long long *bits_k, void *stream) {
  if (b < 0 || lq <= 0 || lk <= 0) return CODA_EINVAL;
  if (b == 0) return CODA_OK;
  if (!mask || !bits_q || !bits_k) return CODA_EINVAL;
  const long long nq = (long long)b * lq * ((lk + 63) / 64), nk = (long long)b * lk * ((lq + 63) / 64);
  cudaStream_t s = (cudaStream_t)stream;
  mask_pack_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, s>>>(b, lq, lk, mask, stride_b, stride_q, stride_k, bits_q);
  mask_pack_kernel<<<(unsigned)((nk + 127) / 128), 128, 0, s>>>(b, lk, lq, mask, stride_b, stride_k, stride_q, bits_k);
  return launch_status();
}

int coda_attention_mask_radius(int b, int l, const float *xyz, float radius, unsigned long long *bits, void *stream) {
  if (b < 0 || l <= 0) return CODA_EINVAL;
  if (b == 0) return CODA_OK;
  if (!xyz || !bits) return CODA_EINVAL;
  const long long n = (long long)b * l * ((l + 63) / 64);
  mask_radius_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(b, l, xyz, radius, bits);
  return launch_status();
}

int coda_attention_dropout_mult(int bh, int lq, int lk, float dropout_p, unsigned int seed,
                                const unsigned int *seed_dev, float *mult, void *stream) {
  if (bh < 0 || lq < 0 || lk < 0 || dropout_p < 0.f || dropout_p >= 1.f) return CODA_EINVAL;
  const long long total = (long long)bh * lq * lk;
  if (total == 0) return CODA_OK;
  if (!mult) return CODA_EINVAL;
  dropout_mult_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      total, lq, lk, seed, seed_dev, drop_thresh32(dropout_p), 1.0f / (1.0f - dropout_p), mult);
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
