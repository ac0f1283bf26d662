// tcgen05 GEMM whose A operand is read as FP32 ROWS and turned into bf16 operand planes INSIDE the kernel:
//
//     C[m][n] = sum_k T(A)[m][k] * B[n][k]  (+ bias[n]) (ReLU),      T = per-element prologue on A
//
// Why: the split-bf16 GEMM of gemm_sm100.cu needs its A operand as 2-3 bf16 planes in HBM -- every fp32
// activation was written again as 6 B/element by a pack kernel and read back by the GEMM, and the BatchNorm
// between two layers of the set-abstraction MLP cost another full read + plane write.  Here A is loaded once as
// fp32 (4 B/element, TMA, 128B-swizzled), four "transform" warps apply T (identity | per-channel affine + ReLU,
// i.e. BatchNorm+ReLU with batch statistics folded into scale/shift | the BatchNorm-backward forms), split the
// value into NSPLIT bf16 planes and write them to TENSOR MEMORY as the MMA's A operand (tcgen05.st; thread = row =
// TMEM lane).  Only B (the weights) goes through shared memory, so the 1/3/6 plane products of a k-block read
// 16-48 KB of smem instead of 32-96 KB: the kernel is no longer shared-memory-port bound.
//
// Epilogue options: bias / ReLU, TMA store of the fp32 tile, and per-column sum / sum-of-squares partials of the
// OUTPUT (the BatchNorm statistics of the layer just computed: no separate pass over C).
//
// Warp roles (512 threads, persistent CTAs, one per SM):
//   warp 0      TMA producer: raw fp32 A boxes -> raw ring, B plane boxes -> B ring
//   warp 1      MMA issuer (one elected lane): tcgen05.mma kind::f16, A from TMEM, B from smem descriptors
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld accumulator -> registers -> (stats) -> swizzled staging -> TMA store
//   warps 8-15  transform: raw ring -> T -> bf16 planes -> TMEM A ring.  Two warps share each 32-row TMEM lane
//               quadrant (warp % 4) and take one half (32 of 64) of the k-block's columns each: the transform is an
//               instruction-issue-bound stream (load, prologue, split, tcgen05.st) and four warps -- one per
//               scheduler, nothing to hide its dependent-issue latency behind -- left the tensor pipe waiting
// C-ABI in include/coda_gemm.h (coda_gemm_a32*).
#include "../../include/coda_gemm.h"
#include "sm100_primitives.cuh"

using namespace coda;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGES = 2;     // TMEM A ring depth
constexpr int RAW_TILE = BM * BK * 4;   // 32 KB: two [128 rows x 32 fp32] SW128 boxes

struct A32Maps {
  CUtensorMap a;      // fp32 [m][k], box [128][32]
  CUtensorMap a2;     // second fp32 input of the two-input modes (same geometry)
  CUtensorMap b[3];   // bf16 planes
  CUtensorMap c;      // fp32 [m][n], box [32][32]
};

__host__ __device__ constexpr int n_products(int ns) { return ns == 1 ? 1 : (ns == 2 ? 3 : 6); }
__host__ __device__ constexpr int prod_a(int ns, int p) {
  return ns == 1 ? 0 : ns == 2 ? (p == 0 ? 1 : 0) : (p == 0 ? 1 : p == 1 ? 2 : p == 2 ? 0 : p == 3 ? 1 : 0);
}
__host__ __device__ constexpr int prod_b(int ns, int p) {
  return ns == 1 ? 0 : ns == 2 ? (p == 1 ? 1 : 0) : (p == 0 ? 1 : p == 1 ? 0 : p == 2 ? 2 : p == 3 ? 0 : p == 4 ? 1 : 0);
}

inline int make_tmap_f32_box(CUtensorMap *map, const void *base, long long cols, long long rows, long long row_stride,
                             int box_cols, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return CODA_EINVAL;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  cuuint64_t gstride[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)row_stride * rows * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CODA_OK : CODA_EINVAL;
}

struct A32Params {
  int m, n, k;             // k = logical contraction length (A columns); B planes are padded to kpad
  int mode;                // CODA_A32_*
  const float *scale;      // per-k, padded to a multiple of 64: modes 1-3 evaluate z = a * scale + shift
  const float *shift;
  const float *alpha;      // modes 2, 3 (BatchNorm backward): T = [z > 0] * scale * d + a * alpha + beta
  const float *beta;
  const float *dpooled;    // mode 3: (m / group, k) gradient of the max-pooled output
  const unsigned char *argmax;   // mode 3: (m / group, k) arg-max row within the group
  int group;               // mode 3
  const float *bias;
  int act;
  float *stats;            // [gridDim.x][2][n] or null
  int b_resident;          // the whole B operand of this CTA's n-tile stays in shared memory (short contractions)
};

// RAW_KB: size of the raw-fp32 staging region; it holds RAW_KB / 32 stages (one-input prologues) or RAW_KB / 64
// (the two-input BatchNorm-backward prologue).
template <int NSPLIT, int BN, int RAW_KB, int B_STAGES, bool B_MN>
__global__ void __launch_bounds__(512, 1)
gemm_a32_kernel(const __grid_constant__ A32Maps maps, const A32Params P) {
  constexpr int MAX_RAW = RAW_KB / 32;
  constexpr int B_TILE = BN * BK * 2;
  constexpr int B_STAGE = NSPLIT * B_TILE;
  constexpr uint32_t ACC_COLS = BN;
  constexpr uint32_t A_COLS = NSPLIT * 32;                       // one A stage: NSPLIT planes x [128 x 64] bf16
  constexpr uint32_t TMEM_COLS = 512;
  static_assert(2 * ACC_COLS + A_STAGES * A_COLS <= TMEM_COLS, "TMEM budget");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *raw_ring = smem;
  const bool two_in = P.mode == CODA_A32_BN_BWD;
  const bool pooled = P.mode == CODA_A32_BN_BWD_POOLED || P.mode == CODA_A32_BN_BWD_POOLED_PRE;
  // pooled BatchNorm backward: each raw stage carries, after the fp32 tile, the (group, channel) gradient and
  // arg-max rows of the tile's groups for this k-block: [groups in tile][64 floats | 64 bytes]  (<= 1 KB)
  const int raw_stage_bytes = RAW_TILE * (two_in ? 2 : 1) + (pooled ? 1024 : 0);
  const uint32_t nraw = (uint32_t)((RAW_KB * 1024) / raw_stage_bytes);
  const int tile_groups = pooled ? (P.group >= BM ? 1 : BM / P.group) : 0;
  unsigned char *b_ring = smem + (size_t)RAW_KB * 1024;
  unsigned char *epi = b_ring + (size_t)B_STAGES * B_STAGE;                // 4 x [32 x 128 B] staging tiles
  float *s_stats = reinterpret_cast<float *>(epi + 4 * 32 * 128);          // [4 warps][2][n]  (only if P.stats)
  __shared__ __align__(8) uint64_t raw_full[MAX_RAW], raw_empty[MAX_RAW];
  __shared__ __align__(8) uint64_t b_full[B_STAGES], b_empty[B_STAGES];
  __shared__ __align__(8) uint64_t a_full[A_STAGES], a_empty[A_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = P.m, n = P.n;
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  const long long nwork = (long long)tiles_m * tiles_n;
  const int nkb = (P.k + BK - 1) / BK;
  // n runs fastest: the n-tiles of one m-tile are in flight together on neighbouring CTAs, A comes from HBM once
  // B-resident mode: a CTA keeps ONE n-tile for its whole life (its B planes are loaded once and never leave
  // shared memory) and walks the m-tiles with stride gridDim / tiles_n.
  auto tile_at = [&](long long t, int &m0, int &n0) -> bool {
    if (P.b_resident) {
      const int per = (int)gridDim.x / tiles_n;
      const long long tm = (long long)(blockIdx.x / tiles_n) + t * per;
      if ((int)blockIdx.x >= per * tiles_n || tm >= tiles_m) return false;
      m0 = (int)tm * BM;
      n0 = (int)(blockIdx.x % tiles_n) * BN;
      return true;
    }
    const long long w = blockIdx.x + t * gridDim.x;
    if (w >= nwork) return false;
    n0 = (int)(w % tiles_n) * BN;
    m0 = (int)(w / tiles_n) * BM;
    return true;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a);
    if (two_in) prefetch_tmap(&maps.a2);
#pragma unroll
    for (int p = 0; p < NSPLIT; ++p) prefetch_tmap(&maps.b[p]);
    prefetch_tmap(&maps.c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_RAW; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 8); }
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&a_full[s], 8); mbar_init(&a_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    mbar_fence_init_cluster();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (P.stats && warp >= 4 && warp < 8) {
    float *mine = s_stats + (size_t)(warp - 4) * 2 * n;
    for (int i = lane; i < 2 * n; i += 32) mine[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_a0 = tmem_base + 2 * ACC_COLS;

  if (warp == 0) {
    // ===== TMA producer =====
    uint32_t it = 0;
    int m0, n0;
    auto load_b = [&](int kb, int bs) {
      unsigned char *bt = b_ring + (size_t)bs * B_STAGE;
      mbar_arrive_expect_tx(&b_full[bs], (uint32_t)B_STAGE);
#pragma unroll
      for (int p = 0; p < NSPLIT; ++p) {
        if (!B_MN) {
          tma_load_3d(bt + p * B_TILE, &maps.b[p], &b_full[bs], kb * BK, n0, 0);
        } else {
#pragma unroll
          for (int g = 0; g < BN / 64; ++g)
            tma_load_3d(bt + p * B_TILE + g * 8192, &maps.b[p], &b_full[bs], n0 + g * 64, kb * BK, 0);
        }
      }
    };
    if (P.b_resident && tile_at(0, m0, n0)) {
      if (elect_one_sync())
        for (int kb = 0; kb < nkb; ++kb) load_b(kb, kb);
      __syncwarp();
    }
    for (long long t = 0; tile_at(t, m0, n0); ++t) {
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int rs = it % nraw, bs = it % B_STAGES;
        mbar_wait(&raw_empty[rs], ((it / nraw) & 1u) ^ 1u);
        if (!P.b_resident) mbar_wait(&b_empty[bs], ((it / B_STAGES) & 1u) ^ 1u);
        if (elect_one_sync()) {
          unsigned char *rt = raw_ring + (size_t)rs * raw_stage_bytes;
          mbar_arrive_expect_tx(&raw_full[rs], (uint32_t)(RAW_TILE * (two_in ? 2 : 1) + tile_groups * 320));
          if (pooled) {
            const long long ngroups = P.m / P.group;
            for (int gt = 0; gt < tile_groups; ++gt) {
              long long g = (long long)m0 / P.group + gt;
              if (g >= ngroups) g = ngroups - 1;        // rows past m: values unused
              bulk_load_1d(rt + RAW_TILE + gt * 320, P.dpooled + g * P.k + kb * BK, 256, &raw_full[rs]);
              bulk_load_1d(rt + RAW_TILE + gt * 320 + 256, P.argmax + g * P.k + kb * BK, 64, &raw_full[rs]);
            }
          }
          tma_load_3d(rt, &maps.a, &raw_full[rs], kb * BK, m0, 0);
          tma_load_3d(rt + RAW_TILE / 2, &maps.a, &raw_full[rs], kb * BK + 32, m0, 0);
          if (two_in) {
            tma_load_3d(rt + RAW_TILE, &maps.a2, &raw_full[rs], kb * BK, m0, 0);
            tma_load_3d(rt + RAW_TILE + RAW_TILE / 2, &maps.a2, &raw_full[rs], kb * BK + 32, m0, 0);
          }
          if (!P.b_resident) load_b(kb, bs);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc_f16(0, BM, BN, 0, B_MN ? 1 : 0);
    uint32_t it = 0, tile_i = 0;
    int m0, n0;
    for (long long t = 0; tile_at(t, m0, n0); ++t, ++tile_i) {
      const uint32_t buf = tile_i & 1u, use = tile_i >> 1;
      mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + buf * ACC_COLS;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int as = it % A_STAGES, bs = P.b_resident ? kb : (int)(it % B_STAGES);
        mbar_wait(&a_full[as], (it / A_STAGES) & 1u);
        mbar_wait(&b_full[bs], P.b_resident ? 0u : ((it / B_STAGES) & 1u));
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_t = tmem_a0 + (uint32_t)as * A_COLS;
          unsigned char *bt = b_ring + (size_t)bs * B_STAGE;
#pragma unroll
          for (int p = 0; p < n_products(NSPLIT); ++p) {
            const uint32_t at = a_t + (uint32_t)prod_a(NSPLIT, p) * 32u;
            const void *btile = bt + prod_b(NSPLIT, p) * B_TILE;
            const uint64_t bd = B_MN ? umma_smem_desc_mn_sw128(btile) : umma_smem_desc_k_sw128(btile);
            constexpr uint32_t KSTEP = B_MN ? 16 * 128 : 32;
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma_f16_ts(tmem_acc, at + kk * 8, umma_desc_advance(bd, kk * KSTEP), idesc, (uint32_t)((kb | p | kk) != 0));
          }
          umma_commit(&a_empty[as]);
          if (!P.b_resident) umma_commit(&b_empty[bs]);
          if (kb == nkb - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 8) {
    // ===== transform: raw fp32 rows -> T -> bf16 planes -> TMEM =====
    const int q = (warp - 8) & 3;                     // TMEM lane quadrant (a warp reaches lanes 32 * (warp % 4) ...)
    const int khalf = (warp - 8) >> 2;                // which 32 of the k-block's 64 columns
    const int row = q * 32 + lane;                    // row of the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int sw = row & 7;
    uint32_t it = 0;
    int m0, n0;
    for (long long t = 0; tile_at(t, m0, n0); ++t) {
      const long long grow = (long long)m0 + row;     // global row (mode 3: its group / index within the group)
      int pgt = 0, pgi = 0;       // pooled mode: this row's group within the tile, its index within the group
      if (pooled) {
        pgi = (int)(grow % P.group);
        pgt = P.group >= BM ? 0 : row / P.group;
      }
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int rs = it % nraw, as = it % A_STAGES;
        mbar_wait_relaxed(&raw_full[rs], (it / nraw) & 1u);
        mbar_wait_relaxed(&a_empty[as], ((it / A_STAGES) & 1u) ^ 1u);
        tc_fence_after();
        const unsigned char *rt = raw_ring + (size_t)rs * raw_stage_bytes;
        const uint32_t a_t = tmem_a0 + (uint32_t)as * A_COLS + lane_base;
        const int k0 = kb * BK;
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {              // 16 k-elements = 8 TMEM columns per plane
          const int ch = khalf * 2 + c2;
          float x[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c16 = ch * 4 + j;               // 16-byte chunk (4 floats) of the 64-float row
            const unsigned char *box = rt + (c16 >> 3) * (RAW_TILE / 2);
            const float4 v = *reinterpret_cast<const float4 *>(box + row * 128 + (((c16 & 7) ^ sw) << 4));
            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
          }
          const int kc = k0 + ch * 16;
          if (P.mode == CODA_A32_AFFINE_RELU) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 s4 = __ldg(reinterpret_cast<const float4 *>(P.scale + kc) + j);
              const float4 t4 = __ldg(reinterpret_cast<const float4 *>(P.shift + kc) + j);
              x[4 * j] = fmaxf(fmaf(x[4 * j], s4.x, t4.x), 0.f);
              x[4 * j + 1] = fmaxf(fmaf(x[4 * j + 1], s4.y, t4.y), 0.f);
              x[4 * j + 2] = fmaxf(fmaf(x[4 * j + 2], s4.z, t4.z), 0.f);
              x[4 * j + 3] = fmaxf(fmaf(x[4 * j + 3], s4.w, t4.w), 0.f);
            }
          } else if (P.mode == CODA_A32_BN_BWD) {
            // x = y (pre-BN activation), d = gradient of relu(bn(y)).  With s = gamma * invstd, t = beta_bn - mean * s:
            //   dy = s * ([s y + t > 0] d - s1/N - xhat s2/N) = [s y + t > 0] * s * d + alpha * y + beta
            //   alpha = -s * invstd * s2 / N,  beta = -s * s1 / N - alpha * mean        (host: coda_bn_bwd_coefs)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c16 = ch * 4 + j;
              const unsigned char *box = rt + RAW_TILE + (c16 >> 3) * (RAW_TILE / 2);
              const float4 d = *reinterpret_cast<const float4 *>(box + row * 128 + (((c16 & 7) ^ sw) << 4));
              const float4 s4 = __ldg(reinterpret_cast<const float4 *>(P.scale + kc) + j);
              const float4 t4 = __ldg(reinterpret_cast<const float4 *>(P.shift + kc) + j);
              const float4 a4 = __ldg(reinterpret_cast<const float4 *>(P.alpha + kc) + j);
              const float4 b4 = __ldg(reinterpret_cast<const float4 *>(P.beta + kc) + j);
              x[4 * j] = (fmaf(x[4 * j], s4.x, t4.x) > 0.f ? s4.x * d.x : 0.f) + fmaf(x[4 * j], a4.x, b4.x);
              x[4 * j + 1] = (fmaf(x[4 * j + 1], s4.y, t4.y) > 0.f ? s4.y * d.y : 0.f) + fmaf(x[4 * j + 1], a4.y, b4.y);
              x[4 * j + 2] = (fmaf(x[4 * j + 2], s4.z, t4.z) > 0.f ? s4.z * d.z : 0.f) + fmaf(x[4 * j + 2], a4.z, b4.z);
              x[4 * j + 3] = (fmaf(x[4 * j + 3], s4.w, t4.w) > 0.f ? s4.w * d.w : 0.f) + fmaf(x[4 * j + 3], a4.w, b4.w);
            }
          } else if (P.mode == CODA_A32_BN_BWD_POOLED_PRE) {
            // pre-masked, pre-scaled pooled gradient: one compare + select + FMA + add per element
            const int gi = pgi;
            const unsigned char *px = rt + RAW_TILE + pgt * 320;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 a4 = __ldg(reinterpret_cast<const float4 *>(P.alpha + kc) + j);
              const float4 b4 = __ldg(reinterpret_cast<const float4 *>(P.beta + kc) + j);
              const float4 d = *reinterpret_cast<const float4 *>(px + (ch * 16 + 4 * j) * 4);
              const uchar4 id = *reinterpret_cast<const uchar4 *>(px + 256 + ch * 16 + 4 * j);
              x[4 * j] = (id.x == gi ? d.x : 0.f) + fmaf(x[4 * j], a4.x, b4.x);
              x[4 * j + 1] = (id.y == gi ? d.y : 0.f) + fmaf(x[4 * j + 1], a4.y, b4.y);
              x[4 * j + 2] = (id.z == gi ? d.z : 0.f) + fmaf(x[4 * j + 2], a4.z, b4.z);
              x[4 * j + 3] = (id.w == gi ? d.w : 0.f) + fmaf(x[4 * j + 3], a4.w, b4.w);
            }
          } else if (P.mode == CODA_A32_BN_BWD_POOLED) {
            // the layer output was max-pooled over `group` rows: only the arg-max row of a (group, channel)
            // carries the incoming gradient dpooled[g][c]
            const int gi = pgi;
            const unsigned char *px = rt + RAW_TILE + pgt * 320;     // staged by the producer with the raw tile
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 s4 = __ldg(reinterpret_cast<const float4 *>(P.scale + kc) + j);
              const float4 t4 = __ldg(reinterpret_cast<const float4 *>(P.shift + kc) + j);
              const float4 a4 = __ldg(reinterpret_cast<const float4 *>(P.alpha + kc) + j);
              const float4 b4 = __ldg(reinterpret_cast<const float4 *>(P.beta + kc) + j);
              const float4 d = *reinterpret_cast<const float4 *>(px + (ch * 16 + 4 * j) * 4);
              const uchar4 id = *reinterpret_cast<const uchar4 *>(px + 256 + ch * 16 + 4 * j);
              x[4 * j] = ((id.x == gi && fmaf(x[4 * j], s4.x, t4.x) > 0.f) ? s4.x * d.x : 0.f) + fmaf(x[4 * j], a4.x, b4.x);
              x[4 * j + 1] = ((id.y == gi && fmaf(x[4 * j + 1], s4.y, t4.y) > 0.f) ? s4.y * d.y : 0.f) + fmaf(x[4 * j + 1], a4.y, b4.y);
              x[4 * j + 2] = ((id.z == gi && fmaf(x[4 * j + 2], s4.z, t4.z) > 0.f) ? s4.z * d.z : 0.f) + fmaf(x[4 * j + 2], a4.z, b4.z);
              x[4 * j + 3] = ((id.w == gi && fmaf(x[4 * j + 3], s4.w, t4.w) > 0.f) ? s4.w * d.w : 0.f) + fmaf(x[4 * j + 3], a4.w, b4.w);
            }
          }
          uint32_t wv[NSPLIT][8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float r0 = x[e], r1 = x[e + 1];
#pragma unroll
            for (int pl = 0; pl < NSPLIT; ++pl) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(r0, r1);
              const uint32_t bits = *reinterpret_cast<const uint32_t *>(&h2);
              wv[pl][e >> 1] = bits;
              if (pl + 1 < NSPLIT) {
                r0 -= __uint_as_float(bits << 16);
                r1 -= __uint_as_float(bits & 0xFFFF0000u);
              }
            }
          }
#pragma unroll
          for (int pl = 0; pl < NSPLIT; ++pl) tmem_st_32x8(a_t + (uint32_t)(pl * 32 + ch * 8), wv[pl]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&raw_empty[rs]);    // this warp's rows of the raw stage have been read
          mbar_arrive(&a_full[as]);       // ... and its 32 lanes of the A stage are in TMEM
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp - 4;
    uint32_t tile_i = 0;
    unsigned char *stage = epi + (size_t)q * (32 * 128);
    unsigned char *srow = stage + lane * 128;
    const int sw = lane & 7;
    float *my_stats = P.stats ? s_stats + (size_t)q * 2 * n : nullptr;
    int m0, n0;
    for (long long t = 0; tile_at(t, m0, n0); ++t, ++tile_i) {
      const uint32_t buf = tile_i & 1u, use = tile_i >> 1;
      const int row = m0 + q * 32 + lane;
      mbar_wait_relaxed(&acc_full[buf], use & 1u);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + buf * ACC_COLS;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int col0 = n0 + c0;
        if (col0 >= n) break;
        float v[32];
        {
          uint32_t r[32];
          tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
        }
        if (P.bias) {
          if (col0 + 32 <= n) {
#pragma unroll
            for (int t = 0; t < 32; t += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4 *>(P.bias + col0 + t));
              v[t] += bb.x; v[t + 1] += bb.y; v[t + 2] += bb.z; v[t + 3] += bb.w;
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (col0 + t < n) v[t] += __ldg(P.bias + col0 + t);
          }
        }
        if (P.act == 1) {
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = fmaxf(v[t], 0.f);
        }
        if (my_stats && row >= m) {      // rows past m are padding: clipped by the TMA store, excluded from the statistics
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = 0.f;
        }
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4 *>(srow + ((j ^ sw) << 4)) =
              make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                         __float_as_uint(v[4 * j + 3]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&maps.c, stage, col0, m0 + q * 32, 0);
          tma_store_commit();
        }
        if (my_stats) {
          // column sums of this warp's 32 rows, read back from the staging tile the TMA store is draining: lane l
          // walks column l (conflict-free through the swizzle) -- 32 loads + 64 FMAs instead of two 31-shuffle
          // transposition trees per chunk.  ncu had the epilogue warps as this kernel's bottleneck on the 1M-row layers
          // (the MMA warp waiting for a free accumulator).
          float cs0 = 0.f, cs1 = 0.f, cq0 = 0.f, cq1 = 0.f;
          const unsigned char *colp = stage + (lane & 3) * 4;
          const int ch = lane >> 2;
#pragma unroll
          for (int r = 0; r < 32; r += 2) {
            const float x0 = *reinterpret_cast<const float *>(colp + r * 128 + ((ch ^ (r & 7)) << 4));
            const float x1 = *reinterpret_cast<const float *>(colp + (r + 1) * 128 + ((ch ^ ((r + 1) & 7)) << 4));
            cs0 += x0; cq0 = fmaf(x0, x0, cq0);
            cs1 += x1; cq1 = fmaf(x1, x1, cq1);
          }
          if (col0 + lane < n) {
            my_stats[col0 + lane] += cs0 + cs1;
            my_stats[n + col0 + lane] += cq0 + cq1;
          }
        }
      }
      if (lane == 0) tma_store_wait_read();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (P.stats) {
    // fixed-order sum of the four epilogue warps' slices -> this CTA's partial row
    float *out = P.stats + (size_t)blockIdx.x * 2 * n;
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x)
      out[i] = (s_stats[i] + s_stats[2 * n + i]) + (s_stats[4 * n + i] + s_stats[6 * n + i]);
  }
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int NSPLIT, int BN, int RAW_KB, int B_STAGES, bool B_MN>
int launch_a32(const A32Maps &maps, const A32Params &P, cudaStream_t s) {
  static_assert(RAW_KB % 64 == 0 || RAW_KB == 96, "raw region");
  constexpr size_t smem = (size_t)RAW_KB * 1024 + (size_t)B_STAGES * NSPLIT * BN * BK * 2 + 4 * 32 * 128 + 1024;
  const size_t total = smem + (P.stats ? (size_t)8 * P.n * 4 : 0);
  constexpr size_t SMEM_MAX = 227 * 1024 - 2048;     // static shared memory (barriers) shares the 227 KB limit
  if (total > SMEM_MAX) return CODA_ETOOLARGE;
  if (P.mode == CODA_A32_BN_BWD && RAW_KB < 64) return CODA_EINVAL;
  auto kern = gemm_a32_kernel<NSPLIT, BN, RAW_KB, B_STAGES, B_MN>;
  static size_t configured = 0;
  if (configured < total) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
    configured = SMEM_MAX;
  }
  const int tiles_m = (P.m + BM - 1) / BM, tiles_n = (P.n + BN - 1) / BN;
  const long long nwork = (long long)tiles_m * tiles_n;
  unsigned grid = (unsigned)(nwork < num_sms() ? nwork : num_sms());
  A32Params Q = P;
  const int nkb = (P.k + BK - 1) / BK;
  // short contraction, many m-tiles: keep the weights of one n-tile resident per CTA
  Q.b_resident = (nkb <= B_STAGES && tiles_n <= num_sms() && tiles_m >= 4 * (num_sms() / tiles_n)) ? 1 : 0;
  if (Q.b_resident) grid = (unsigned)((num_sms() / tiles_n) * tiles_n);
  kern<<<grid, 512, total, s>>>(maps, Q);
  return launch_status();
}

}  // namespace

extern "C" {

int coda_gemm_a32_grid(int m, int n) {
  // upper bound of the launch grid = rows of the (zero-initialised) col_stats buffer the caller provides
  if (m <= 0 || n <= 0) return 0;
  return num_sms();
}

int coda_gemm_a32(int nsplit, int m, int n, int k, const float *a, long long lda, int a_mode, const float *a_scale,
                  const float *a_shift, const float *a_alpha, const float *a_beta, const float *a2, long long lda2,
                  const unsigned char *a_argmax, int a_group, const void *b_planes, long long b_plane_stride,
                  int b_ld, int b_mn, const float *bias, int act, float *c, long long ldc, float *col_stats,
                  void *stream) {
  if (nsplit < 2 || nsplit > 3 || m < 0 || n < 0 || k <= 0) return CODA_EINVAL;
  if (m == 0 || n == 0) return CODA_OK;
  if (!a || !b_planes || !c || b_ld % 64 != 0 || (lda & 3) || (ldc & 3) || ((uintptr_t)a & 15) || ((uintptr_t)c & 15))
    return CODA_EINVAL;
  if (a_mode < CODA_A32_PLAIN || a_mode > CODA_A32_BN_BWD_POOLED_PRE) return CODA_EINVAL;
  if (a_mode != CODA_A32_PLAIN && (!a_scale || !a_shift || (k & 3))) return CODA_EINVAL;
  if (a_mode >= CODA_A32_BN_BWD && (!a2 || !a_alpha || !a_beta || ((uintptr_t)a2 & 15))) return CODA_EINVAL;
  if (a_mode == CODA_A32_BN_BWD && (lda2 & 3)) return CODA_EINVAL;
  if ((a_mode == CODA_A32_BN_BWD_POOLED || a_mode == CODA_A32_BN_BWD_POOLED_PRE) &&
      (!a_argmax || a_group < 32 || a_group > 256 || m % a_group != 0 || k % 64 != 0 ||
       !(a_group % 128 == 0 || 128 % a_group == 0) || a_group % 32 != 0))
    return CODA_EINVAL;     // groups must tile the 128-row blocks (warp = 32 rows of one group)
  // few output tiles (the decoder's 2048-row linears: 16 x 4 tiles of 128 x 128 on 148 SMs): 64-wide tiles double
  // the number of CTAs at work; each of these launches is bounded by one CTA's serial tile time, not by throughput
  const long long tiles128 = (long long)((m + BM - 1) / BM) * ((n + 127) / 128);
  const int bn = (n <= 64 || tiles128 <= num_sms() / 2) ? 64 : 128;
  A32Maps maps;
  int st = make_tmap_f32_box(&maps.a, a, k, m, lda, 32, BM);
  if (st != CODA_OK) return st;
  maps.a2 = maps.a;
  if (a_mode == CODA_A32_BN_BWD) {
    st = make_tmap_f32_box(&maps.a2, a2, k, m, lda2, 32, BM);
    if (st != CODA_OK) return st;
  }
  const int kpad = (k + 63) / 64 * 64;
  const char *bp = (const char *)b_planes;
  for (int p = 0; p < nsplit; ++p) {
    if (!b_mn) {
      // planes [n rows][b_ld >= kpad], K contiguous; box [bn rows][64]
      if (b_ld < kpad) return CODA_EINVAL;
      st = make_tmap_k_major_16b(&maps.b[p], bp + (size_t)p * b_plane_stride * 2, 0, b_ld, n, 1, b_ld, 0, bn);
    } else {
      // planes [k rows][b_ld >= n], N contiguous (the forward weight planes, used for dX = dY W); box [64][64]
      if (b_ld < n) return CODA_EINVAL;
      st = make_tmap_k_major_16b(&maps.b[p], bp + (size_t)p * b_plane_stride * 2, 0, b_ld, k, 1, b_ld, 0, 64);
    }
    if (st != CODA_OK) return st;
  }
  for (int p = nsplit; p < 3; ++p) maps.b[p] = maps.b[0];
  st = make_tmap_rows_f32(&maps.c, c, n, m, 1, ldc, 0);
  if (st != CODA_OK) return st;
  A32Params P;
  P.m = m; P.n = n; P.k = k; P.mode = a_mode; P.scale = a_scale; P.shift = a_shift; P.alpha = a_alpha; P.beta = a_beta;
  P.dpooled = a2; P.argmax = a_argmax; P.group = a_group; P.bias = bias; P.act = act; P.stats = col_stats;
  P.b_resident = 0;
  cudaStream_t s = (cudaStream_t)stream;
#define CODA_A32(NS, BN_, RS, BS)                                                   \
  return b_mn ? launch_a32<NS, BN_, RS, BS, true>(maps, P, s) : launch_a32<NS, BN_, RS, BS, false>(maps, P, s)
  // smem: raw region + B ring + 16 KB staging (+ stats): <= 225 KB.  The B (weight) tiles come from L2 with ~1 us
  // latency: long contractions want a deep B ring, the two-input prologue a wide raw region.
  const bool deep_b = !col_stats && k > 128 && a_mode != CODA_A32_BN_BWD;
  if (nsplit == 2) {
    if (bn == 64) CODA_A32(2, 64, 128, 4);
    CODA_A32(2, 128, 128, 2);
  }
  if (bn == 64) CODA_A32(3, 64, 128, 3);
  if (deep_b) CODA_A32(3, 128, 64, 3);
  CODA_A32(3, 128, 96, 2);
#undef CODA_A32
}

}  // extern "C"
