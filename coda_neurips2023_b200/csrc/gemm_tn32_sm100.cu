// Weight-gradient GEMM on fp32 ROWS with in-kernel prologues (tcgen05, sm_100a):
//
//     C[m][n] = sum_r  TA(A)[r][m] * TB(B)[r][n]          (dW = dY^T X, contraction over the ROWS of both operands)
//
// A (R, M) and B (R, N) are the row-major fp32 activations as they sit in HBM.  In gemm_sm100.cu's TN form both had
// to exist as row-packed bf16 planes first (a pack kernel per operand: read 4 B, write 4 B, read 4 B again per
// element); here every 32-row slab is TMA-loaded as fp32, eight transform warps apply the prologue
//     TA: identity | BatchNorm+ReLU backward of the layer (dense two-input form, or max-pooled form)
//     TB: identity | BatchNorm+ReLU forward of the previous layer (relu(b * scale + shift))
// split the values into two bf16 planes and write them as MN-major, 128B-swizzled tensor-core operands in shared
// memory.  The kernel is a pure stream over R: split-K across all SMs, partial tiles leave through
// cp.reduce.async.bulk.tensor (.add) into the zero-initialised output.
//
// Warp roles (640 threads): 0 TMA producer | 1 MMA issuer | 2 TMEM allocator | 4-19 transform (two rows of each operand
// slab per warp); warps 4-7 drain the accumulator once the stream has ended.  The transform is an instruction-issue
// bound stream: with eight warps (two per scheduler) the kernel ran at 0.44 of the HBM roofline with the tensor pipe
// 25 % busy; sixteen hide the shared-memory and conversion latencies of one another.
// Optional: the column sums of TA(A) (the bias gradient of the layer whose dW this is) are accumulated by the
// transform warps on the way -- the values are in registers already -- instead of a separate pass over dY.
// C-ABI in include/coda_gemm.h (coda_gemm_tn32).
#include "../../include/coda_gemm.h"
#include "sm100_primitives.cuh"

using namespace coda;

namespace {

constexpr int BM = 128;      // output rows per tile  (columns of A)
constexpr int BKR = 32;      // contraction rows per pipeline stage
constexpr int NS = 2;        // bf16 planes per operand (gradient precision, as the packed TN path)
constexpr int RAW_STAGES = 3, PL_STAGES = 2;   // fp32 slabs in flight (HBM latency) / converted operand slabs
constexpr int TW = 16;       // transform warps

struct TN32Maps {
  CUtensorMap a, a2, b, c;
};

struct TN32Params {
  long long rows;
  int m, n;
  int a_mode, b_mode;        // CODA_A32_*
  const float *a_scale, *a_shift, *a_alpha, *a_beta;    // per column of A (padded to a multiple of 128)
  const float *dpooled;      // pooled form: (rows / group, m)
  const unsigned char *argmax;
  int group;
  const float *b_scale, *b_shift;                       // per column of B (padded)
  float *a_colsum;           // optional (m): += column sums of TA(A), pre-zeroed by the launcher
  int ksplit;
};

// MN-major, 128B-swizzled operand made of [BKR k-rows x 64 mn] boxes that are `lbo` bytes apart
__device__ __forceinline__ uint64_t desc_mn_sw128(const void *tile, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(tile) & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void store_planes4(float4 v, unsigned char *dst, uint32_t plane_bytes) {
  float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int p = 0; p < NS; ++p) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(r[0], r[1]), hi = __floats2bfloat162_rn(r[2], r[3]);
    uint2 w;
    w.x = *reinterpret_cast<const uint32_t *>(&lo);
    w.y = *reinterpret_cast<const uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(dst + (size_t)p * plane_bytes) = w;
    if (p + 1 < NS) {
      r[0] -= __uint_as_float(w.x << 16); r[1] -= __uint_as_float(w.x & 0xFFFF0000u);
      r[2] -= __uint_as_float(w.y << 16); r[3] -= __uint_as_float(w.y & 0xFFFF0000u);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(640, 1)
gemm_tn32_kernel(const __grid_constant__ TN32Maps maps, const TN32Params P) {
  constexpr int RAW_A = BKR * BM * 4;            // 16 KB: four [32 rows x 32 fp32] SW128 boxes
  constexpr int RAW_B = BKR * BN * 4;
  constexpr int RAW_X = 1024;                    // pooled mode: [128 floats dpooled | 128 bytes argmax] of the slab's group
  constexpr int RAW_STAGE = 2 * RAW_A + RAW_B + RAW_X;   // (second A input only in the dense BN-backward mode)
  constexpr int PL_A = BKR * BM * 2;             // one bf16 plane of the A slab: two [32 x 64] boxes
  constexpr int PL_B = BKR * BN * 2;
  constexpr int PL_STAGE = NS * (PL_A + PL_B);
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *raw_ring = smem;
  unsigned char *pl_ring = smem + (size_t)RAW_STAGES * RAW_STAGE;
  unsigned char *epi = raw_ring;      // 4 x [32 x 128 B] staging tiles: the raw ring is idle once the last MMA is done
  __shared__ __align__(8) uint64_t raw_full[RAW_STAGES], raw_empty[RAW_STAGES];
  __shared__ __align__(8) uint64_t pl_full[PL_STAGES], pl_empty[PL_STAGES];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_colsum[BM];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < BM) s_colsum[threadIdx.x] = 0.f;
  const int tiles_n = (P.n + BN - 1) / BN;
  const int tile = blockIdx.x / P.ksplit, ks = blockIdx.x % P.ksplit;
  const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
  const long long nkb_total = (P.rows + BKR - 1) / BKR;
  const long long per = (nkb_total + P.ksplit - 1) / P.ksplit;
  const long long kb0 = (long long)ks * per;
  const long long nkb = max(0ll, min(per, nkb_total - kb0));
  const bool two_in = P.a_mode == CODA_A32_BN_BWD;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a);
    prefetch_tmap(&maps.b);
    prefetch_tmap(&maps.c);
    if (two_in) prefetch_tmap(&maps.a2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < RAW_STAGES; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], TW); }
    for (int s = 0; s < PL_STAGES; ++s) { mbar_init(&pl_full[s], TW); mbar_init(&pl_empty[s], 1); }
    mbar_init(&acc_full, 1);
    mbar_fence_init_cluster();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_slot;
  if (nkb == 0) {   // more splits than k-blocks: nothing to add
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_acc, TMEM_COLS);
    return;
  }

  if (warp == 0) {
    // ===== TMA producer =====
    const bool pooled = P.a_mode == CODA_A32_BN_BWD_POOLED || P.a_mode == CODA_A32_BN_BWD_POOLED_PRE;
    const uint32_t bytes = (uint32_t)(RAW_A * (two_in ? 2 : 1) + RAW_B + (pooled ? 640 : 0));
    const long long ngroups = pooled ? P.rows / P.group : 0;
    for (long long i = 0; i < nkb; ++i) {
      const int rs = (int)(i % RAW_STAGES);
      mbar_wait(&raw_empty[rs], (uint32_t)((i / RAW_STAGES) & 1) ^ 1u);
      if (elect_one_sync()) {
        unsigned char *st = raw_ring + (size_t)rs * RAW_STAGE;
        const int r0 = (int)((kb0 + i) * BKR);
        mbar_arrive_expect_tx(&raw_full[rs], bytes);
        if (pooled) {      // the slab (32 rows) lies in one group (group % 32 == 0)
          long long g = (long long)r0 / P.group;
          if (g >= ngroups) g = ngroups - 1;
          // m0 + 128 <= padded m: the host guarantees m % 128 == 0 in this mode
          bulk_load_1d(st + 2 * RAW_A + RAW_B, P.dpooled + g * P.m + m0, 512, &raw_full[rs]);
          bulk_load_1d(st + 2 * RAW_A + RAW_B + 512, P.argmax + g * P.m + m0, 128, &raw_full[rs]);
        }
#pragma unroll
        for (int g = 0; g < BM / 32; ++g) tma_load_3d(st + g * 4096, &maps.a, &raw_full[rs], m0 + g * 32, r0, 0);
        if (two_in) {
#pragma unroll
          for (int g = 0; g < BM / 32; ++g) tma_load_3d(st + RAW_A + g * 4096, &maps.a2, &raw_full[rs], m0 + g * 32, r0, 0);
        }
#pragma unroll
        for (int g = 0; g < BN / 32; ++g) tma_load_3d(st + 2 * RAW_A + g * 4096, &maps.b, &raw_full[rs], n0 + g * 32, r0, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = umma_idesc_f16(0, BM, BN, 1, 1);
    for (long long i = 0; i < nkb; ++i) {
      const int ps = (int)(i % PL_STAGES);
      mbar_wait(&pl_full[ps], (uint32_t)((i / PL_STAGES) & 1));
      tc_fence_after();
      if (elect_one_sync()) {
        unsigned char *st = pl_ring + (size_t)ps * PL_STAGE;
        // plane products: lo*hi, hi*lo, hi*hi (small terms first)
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const int pa = p == 0 ? 1 : 0, pb = p == 1 ? 1 : 0;
          const uint64_t ad = desc_mn_sw128(st + pa * PL_A, BKR * 128);
          const uint64_t bd = desc_mn_sw128(st + NS * PL_A + pb * PL_B, BKR * 128);
#pragma unroll
          for (int kk = 0; kk < BKR / 16; ++kk)
            umma_f16(tmem_acc, umma_desc_advance(ad, kk * 16 * 128), umma_desc_advance(bd, kk * 16 * 128), idesc,
                     (uint32_t)((i | p | kk) != 0));
        }
        umma_commit(&pl_empty[ps]);
        if (i == nkb - 1) umma_commit(&acc_full);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== transform: fp32 slabs -> prologue -> two bf16 planes, MN-major swizzled =====
    const int tw = warp - 4;                         // rows tw, tw + 16 of the slab
    const bool want_colsum = P.a_colsum != nullptr && n0 == 0;
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    // A: lane = 4-column chunk (128 columns = 32 chunks);  per-column coefficients live in registers
    const int ca = m0 + lane * 4;
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), ta = sa, al = sa, be = sa;
    if (P.a_mode != CODA_A32_PLAIN && ca < P.m) {
      sa = __ldg(reinterpret_cast<const float4 *>(P.a_scale + ca));
      ta = __ldg(reinterpret_cast<const float4 *>(P.a_shift + ca));
      al = __ldg(reinterpret_cast<const float4 *>(P.a_alpha + ca));
      be = __ldg(reinterpret_cast<const float4 *>(P.a_beta + ca));
    }
    // B: BN / 4 chunks per row; with BN = 64 a warp covers two rows per pass
    constexpr int B_CHUNKS = BN / 4, B_RPP = 32 / B_CHUNKS;        // rows per pass
    const int bch = lane % B_CHUNKS, brow_off = lane / B_CHUNKS;
    const int cb = n0 + bch * 4;
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), tb = sb;
    if (P.b_mode == CODA_A32_AFFINE_RELU && cb < P.n) {
      sb = __ldg(reinterpret_cast<const float4 *>(P.b_scale + cb));
      tb = __ldg(reinterpret_cast<const float4 *>(P.b_shift + cb));
    }
    const bool a_col_ok = ca < P.m;
    const bool pooled_mode = P.a_mode == CODA_A32_BN_BWD_POOLED || P.a_mode == CODA_A32_BN_BWD_POOLED_PRE;
    // the slab's first row within its group, carried from slab to slab (group >= 32 = BKR: one conditional subtract)
    int rem0 = pooled_mode ? (int)((kb0 * BKR) % P.group) : 0;
    int rs = 0, ps = 0;
    uint32_t raw_par = 0, pl_par = 1;
    for (long long i = 0; i < nkb; ++i) {
      mbar_wait_relaxed(&raw_full[rs], raw_par);
      mbar_wait_relaxed(&pl_empty[ps], pl_par);
      const unsigned char *raw = raw_ring + (size_t)rs * RAW_STAGE;
      unsigned char *pl = pl_ring + (size_t)ps * PL_STAGE;
      const long long r0 = (kb0 + i) * BKR;
      const bool tail = r0 + BKR > P.rows;         // only the very last slab can hold rows past the end
      // ---- A slab
#pragma unroll
      for (int j = 0; j < BKR / TW; ++j) {
        const int r = tw + j * TW;
        const long long grow = r0 + r;
        const uint32_t roff = (uint32_t)(lane >> 3) * 4096u + (uint32_t)r * 128u + (uint32_t)(((lane & 7) ^ (r & 7)) << 4);
        float4 y = *reinterpret_cast<const float4 *>(raw + roff);
        float4 o = y;
        if (P.a_mode == CODA_A32_BN_BWD) {
          const float4 d = *reinterpret_cast<const float4 *>(raw + RAW_A + roff);
          o.x = (fmaf(y.x, sa.x, ta.x) > 0.f ? sa.x * d.x : 0.f) + fmaf(y.x, al.x, be.x);
          o.y = (fmaf(y.y, sa.y, ta.y) > 0.f ? sa.y * d.y : 0.f) + fmaf(y.y, al.y, be.y);
          o.z = (fmaf(y.z, sa.z, ta.z) > 0.f ? sa.z * d.z : 0.f) + fmaf(y.z, al.z, be.z);
          o.w = (fmaf(y.w, sa.w, ta.w) > 0.f ? sa.w * d.w : 0.f) + fmaf(y.w, al.w, be.w);
        } else if (P.a_mode == CODA_A32_BN_BWD_POOLED_PRE) {
          const int gi = rem0 + r;
          const float4 d = *reinterpret_cast<const float4 *>(raw + 2 * RAW_A + RAW_B + lane * 16);
          const uchar4 id = *reinterpret_cast<const uchar4 *>(raw + 2 * RAW_A + RAW_B + 512 + lane * 4);
          o.x = (id.x == gi ? d.x : 0.f) + fmaf(y.x, al.x, be.x);
          o.y = (id.y == gi ? d.y : 0.f) + fmaf(y.y, al.y, be.y);
          o.z = (id.z == gi ? d.z : 0.f) + fmaf(y.z, al.z, be.z);
          o.w = (id.w == gi ? d.w : 0.f) + fmaf(y.w, al.w, be.w);
        } else if (P.a_mode == CODA_A32_BN_BWD_POOLED) {
          const int gi = rem0 + r;
          const float4 d = *reinterpret_cast<const float4 *>(raw + 2 * RAW_A + RAW_B + lane * 16);
          const uchar4 id = *reinterpret_cast<const uchar4 *>(raw + 2 * RAW_A + RAW_B + 512 + lane * 4);
          o.x = ((id.x == gi && fmaf(y.x, sa.x, ta.x) > 0.f) ? sa.x * d.x : 0.f) + fmaf(y.x, al.x, be.x);
          o.y = ((id.y == gi && fmaf(y.y, sa.y, ta.y) > 0.f) ? sa.y * d.y : 0.f) + fmaf(y.y, al.y, be.y);
          o.z = ((id.z == gi && fmaf(y.z, sa.z, ta.z) > 0.f) ? sa.z * d.z : 0.f) + fmaf(y.z, al.z, be.z);
          o.w = ((id.w == gi && fmaf(y.w, sa.w, ta.w) > 0.f) ? sa.w * d.w : 0.f) + fmaf(y.w, al.w, be.w);
        }
        if (tail && grow >= P.rows) o = make_float4(0.f, 0.f, 0.f, 0.f);     // padding rows of the last slab
        if (want_colsum) { csum.x += o.x; csum.y += o.y; csum.z += o.z; csum.w += o.w; }
        // destination: box = column / 64, 16-byte chunk = (column % 64) / 8 swizzled by the row, half = (column % 8) / 4
        const int col = lane * 4;
        store_planes4(o, pl + (col >> 6) * (BKR * 128) + r * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4) + ((col & 7) >> 2) * 8,
                      PL_A);
      }
      // ---- B slab
#pragma unroll
      for (int j = 0; j < BKR / (TW * B_RPP); ++j) {
        const int r = (tw + j * TW) * B_RPP + brow_off;
        const long long grow = r0 + r;
        const uint32_t roff = (uint32_t)(bch >> 3) * 4096u + (uint32_t)r * 128u + (uint32_t)(((bch & 7) ^ (r & 7)) << 4);
        float4 v = *reinterpret_cast<const float4 *>(raw + 2 * RAW_A + roff);
        if (P.b_mode == CODA_A32_AFFINE_RELU) {
          v.x = fmaxf(fmaf(v.x, sb.x, tb.x), 0.f); v.y = fmaxf(fmaf(v.y, sb.y, tb.y), 0.f);
          v.z = fmaxf(fmaf(v.z, sb.z, tb.z), 0.f); v.w = fmaxf(fmaf(v.w, sb.w, tb.w), 0.f);
        }
        if (tail && grow >= P.rows) v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int col = bch * 4;
        store_planes4(v, pl + NS * PL_A + (col >> 6) * (BKR * 128) + r * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4) +
                             ((col & 7) >> 2) * 8,
                      PL_B);
      }
      fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&raw_empty[rs]);
        mbar_arrive(&pl_full[ps]);
      }
      if (++rs == RAW_STAGES) { rs = 0; raw_par ^= 1u; }
      if (++ps == PL_STAGES) { ps = 0; pl_par ^= 1u; }
      if (pooled_mode) { rem0 += BKR; if (rem0 >= P.group) rem0 -= P.group; }
    }
    if (want_colsum && a_col_ok) {
      atomicAdd(&s_colsum[lane * 4], csum.x); atomicAdd(&s_colsum[lane * 4 + 1], csum.y);
      atomicAdd(&s_colsum[lane * 4 + 2], csum.z); atomicAdd(&s_colsum[lane * 4 + 3], csum.w);
    }
  }
  if (warp >= 4 && warp < 8) {
    // ===== epilogue: one accumulator per CTA, drained once (the raw ring is idle by then: staging lives there) =====
    const int q = warp - 4;
    unsigned char *stage = epi + (size_t)q * (32 * 128);
    unsigned char *srow = stage + lane * 128;
    const int sw = lane & 7;
    mbar_wait_relaxed(&acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      const int col0 = n0 + c0;
      if (col0 >= P.n) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4 *>(srow + ((j ^ sw) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (P.ksplit > 1) tma_reduce_add_3d(&maps.c, stage, col0, m0 + q * 32, 0);
        else tma_store_3d(&maps.c, stage, col0, m0 + q * 32, 0);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_read();
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (P.a_colsum && n0 == 0 && threadIdx.x < BM && m0 + (int)threadIdx.x < P.m)
    atomicAdd(P.a_colsum + m0 + threadIdx.x, s_colsum[threadIdx.x]);
  if (warp == 2) tmem_dealloc(tmem_acc, TMEM_COLS);
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

template <int BN>
int launch_tn32(const TN32Maps &maps, TN32Params P, float *c, long long ldc, cudaStream_t s) {
  constexpr size_t smem = (size_t)RAW_STAGES * (2 * BKR * BM * 4 + BKR * BN * 4 + 1024) +
                          (size_t)PL_STAGES * NS * (BKR * BM * 2 + BKR * BN * 2) + 1024;
  static_assert(smem <= 227 * 1024, "smem budget");
  auto kern = gemm_tn32_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int tiles = ((P.m + BM - 1) / BM) * ((P.n + BN - 1) / BN);
  const long long nkb_total = (P.rows + BKR - 1) / BKR;
  int ksplit = num_sms / tiles;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > nkb_total / 4) ksplit = (int)(nkb_total / 4 > 0 ? nkb_total / 4 : 1);   // >= 4 slabs per CTA
  P.ksplit = ksplit;
  if (ksplit > 1) {
    cudaError_t e = cudaMemset2DAsync(c, (size_t)ldc * 4, 0, (size_t)P.n * 4, (size_t)P.m, s);
    if (e != cudaSuccess) return (int)e;
  }
  if (P.a_colsum) {
    cudaError_t e = cudaMemsetAsync(P.a_colsum, 0, (size_t)P.m * 4, s);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<tiles * ksplit, 640, smem, s>>>(maps, P);
  return launch_status();
}

}  // namespace

extern "C" {

int coda_gemm_tn32(long long rows, int m, int n, const float *a, long long lda, int a_mode, const float *a_scale,
                   const float *a_shift, const float *a_alpha, const float *a_beta, const float *a2, long long lda2,
                   const unsigned char *a_argmax, int a_group, const float *b, long long ldb, int b_mode,
                   const float *b_scale, const float *b_shift, float *c, long long ldc, float *a_colsum, void *stream) {
  if (rows <= 0 || m <= 0 || n <= 0 || !a || !b || !c) return CODA_EINVAL;
  if ((lda & 3) || (ldb & 3) || (ldc & 3) || (m & 3) || (n & 3)) return CODA_EINVAL;
  if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) return CODA_EINVAL;
  if (a_mode != CODA_A32_PLAIN && a_mode != CODA_A32_BN_BWD && a_mode != CODA_A32_BN_BWD_POOLED &&
      a_mode != CODA_A32_BN_BWD_POOLED_PRE)
    return CODA_EINVAL;
  if (b_mode != CODA_A32_PLAIN && b_mode != CODA_A32_AFFINE_RELU) return CODA_EINVAL;
  if (a_mode != CODA_A32_PLAIN && (!a_scale || !a_shift || !a_alpha || !a_beta || !a2 || ((uintptr_t)a2 & 15))) return CODA_EINVAL;
  if (a_mode == CODA_A32_BN_BWD && (lda2 & 3)) return CODA_EINVAL;
  if ((a_mode == CODA_A32_BN_BWD_POOLED || a_mode == CODA_A32_BN_BWD_POOLED_PRE) &&
      (!a_argmax || a_group < 32 || a_group > 256 || a_group % 32 != 0 || rows % a_group != 0 || m % 128 != 0))
    return CODA_EINVAL;     // a 32-row slab must lie in one group; the group rows are staged 128 columns at a time
  if (b_mode == CODA_A32_AFFINE_RELU && (!b_scale || !b_shift)) return CODA_EINVAL;
  TN32Maps maps;
  int st = make_tmap_f32_box(&maps.a, a, m, rows, lda, 32, BKR);
  if (st != CODA_OK) return st;
  maps.a2 = maps.a;
  if (a_mode == CODA_A32_BN_BWD) {
    st = make_tmap_f32_box(&maps.a2, a2, m, rows, lda2, 32, BKR);
    if (st != CODA_OK) return st;
  }
  st = make_tmap_f32_box(&maps.b, b, n, rows, ldb, 32, BKR);
  if (st != CODA_OK) return st;
  st = make_tmap_rows_f32(&maps.c, c, n, m, 1, ldc, 0);
  if (st != CODA_OK) return st;
  TN32Params P;
  P.rows = rows; P.m = m; P.n = n; P.a_mode = a_mode; P.b_mode = b_mode;
  P.a_scale = a_scale; P.a_shift = a_shift; P.a_alpha = a_alpha; P.a_beta = a_beta;
  P.dpooled = a2; P.argmax = a_argmax; P.group = a_group; P.b_scale = b_scale; P.b_shift = b_shift; P.a_colsum = a_colsum; P.ksplit = 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (n <= 64) return launch_tn32<64>(maps, P, c, ldc, s);
  return launch_tn32<128>(maps, P, c, ldc, s);
}

}  // extern "C"
