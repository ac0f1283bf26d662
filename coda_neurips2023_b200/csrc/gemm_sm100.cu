// tcgen05 / TMA GEMM for B200:  C[b][m][n] = sum_k A[b][m][k] * B[b][n][k]  (+ bias[n]) (ReLU)
//
// Operands are 16-bit, K-major (K contiguous) planes; fp32 inputs are first split
// into NSPLIT bf16 planes (x = hi + lo (+ lo2)) by the pack kernel below and the
// GEMM accumulates the cross products hi*hi + hi*lo + lo*hi (+ ...) in fp32 in
// TMEM, which gives fp32-class accuracy on the bf16 tensor-core path.  NSPLIT = 1
// with fp16 operands is the CLIP ViT path.
//
// Structure (one CTA per 128 x BN output tile, 256 threads):
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem stages)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma.cta_group::1.kind::f16, accumulator in TMEM)
//   warp 2        : TMEM allocator / deallocator
//   warps 4..7    : epilogue       (tcgen05.ld 32x32b -> +bias/ReLU -> global)
// C-ABI in include/coda_gemm.h.
#include "../../include/coda_gemm.h"
#include "sm100_primitives.cuh"

using namespace coda;

namespace {

// ------------------------------------------------------------------ pack (fp32 -> bf16 planes)
template <int NSPLIT>
__device__ __forceinline__ void split_store(float x, __nv_bfloat16 *dst, size_t plane_stride) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  dst[0] = h;
  if (NSPLIT >= 2) {
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    dst[plane_stride] = m;
    if (NSPLIT >= 3) dst[2 * plane_stride] = __float2bfloat16_rn(r1 - __bfloat162float(m));
  }
}

// source is row-major along k (src_k_stride == 1): one thread per (row, k) element, k fastest
template <int NSPLIT>
__global__ void __launch_bounds__(256)
pack_rows_kernel(long long rows, int k, int kpad, long long src_row_stride, const float *__restrict__ src,
                 float scale, __nv_bfloat16 *__restrict__ planes, long long plane_stride) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * kpad) return;
  const long long r = idx / kpad;
  const int c = (int)(idx - r * kpad);
  const float x = c < k ? __ldg(src + r * src_row_stride + c) * scale : 0.f;
  split_store<NSPLIT>(x, planes + idx, (size_t)plane_stride);
}

// same, four consecutive k per thread (k % 4 == 0, 16-byte aligned rows): one 16-byte load, one 8-byte store per plane
template <int NSPLIT>
__global__ void __launch_bounds__(256)
pack_rows_vec4_kernel(long long rows, int k, int kpad, long long src_row_stride, const float *__restrict__ src,
                      float scale, __nv_bfloat16 *__restrict__ planes, long long plane_stride) {
  const int kq = kpad >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * kq) return;
  const long long r = idx / kq;
  const int c = (int)(idx - r * kq) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < k) v = __ldg(reinterpret_cast<const float4 *>(src + r * src_row_stride + c));
  float x[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
  __nv_bfloat16 *dst = planes + r * kpad + c;
#pragma unroll
  for (int p = 0; p < NSPLIT; ++p) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(x[0], x[1]), hi = __floats2bfloat162_rn(x[2], x[3]);
    uint2 w;
    w.x = *reinterpret_cast<const uint32_t *>(&lo);
    w.y = *reinterpret_cast<const uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(dst + (size_t)p * plane_stride) = w;
    if (p + 1 < NSPLIT) {
      x[0] -= __uint_as_float(w.x << 16); x[1] -= __uint_as_float(w.x & 0xFFFF0000u);
      x[2] -= __uint_as_float(w.y << 16); x[3] -= __uint_as_float(w.y & 0xFFFF0000u);
    }
  }
}

// source is contiguous along rows (src_row_stride == 1, "transposed" operand): 32x32 smem transpose
template <int NSPLIT>
__global__ void __launch_bounds__(256)
pack_transposed_kernel(long long rows, int k, int kpad, long long src_k_stride, const float *__restrict__ src,
                       float scale, __nv_bfloat16 *__restrict__ planes, long long plane_stride) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const long long r = r0 + tx;
    tile[i][tx] = (c < k && r < rows) ? __ldg(src + (long long)c * src_k_stride + r) * scale : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const long long r = r0 + i;
    const int c = c0 + tx;
    if (r < rows && c < kpad) split_store<NSPLIT>(tile[tx][i], planes + r * kpad + c, (size_t)plane_stride);
  }
}

// ------------------------------------------------------------------ GEMM
constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 2 B = one 128-byte swizzle span

struct GemmMaps {
  CUtensorMap a[3];
  CUtensorMap b[3];
  CUtensorMap c;   // output, for the TMA-store epilogue (valid when the kernel's tma_store flag is set)
};
constexpr int EPI_STAGE = 4 * 32 * 128;   // four epilogue warps x [32 rows x 128 B] staging tiles

// which (A plane, B plane) pairs are multiplied; small cross terms first
__host__ __device__ constexpr int n_products(int ns) { return ns == 1 ? 1 : (ns == 2 ? 3 : 6); }
__host__ __device__ constexpr int prod_a(int ns, int p) {
  return ns == 1 ? 0 : ns == 2 ? (p == 0 ? 1 : 0) : (p == 0 ? 1 : p == 1 ? 2 : p == 2 ? 0 : p == 3 ? 1 : 0);
}
__host__ __device__ constexpr int prod_b(int ns, int p) {
  return ns == 1 ? 0 : ns == 2 ? (p == 1 ? 1 : 0) : (p == 0 ? 1 : p == 1 ? 0 : p == 2 ? 2 : p == 3 ? 0 : p == 4 ? 1 : 0);
}

// MN = false: C = A B^T with K-contiguous operands ("NT").
// MN = true : both operands are stored with the contraction index as the ROW index (A planes
//             [kc][m], B planes [kc][n]) -- the weight-gradient form dW = dY^T X, which then needs no
//             transposed copies of dY and X.
template <int NSPLIT, int BN, int STAGES, bool FP16, bool MN>
__global__ void __launch_bounds__(256, 1)
gemm_nt_kernel(const __grid_constant__ GemmMaps maps, int m, int n, int kpad, int b_batched, int ksplit, int batch_n,
               const float *__restrict__ bias_in, int act, int out_half, void *__restrict__ c_void, long long ldc,
               long long c_batch_stride, int tma_store, const void *__restrict__ residual, long long ldr) {
  // PERSISTENT: each CTA walks work items w = blockIdx.x, blockIdx.x + gridDim.x, ...;  a work item is
  // (m-tile, n-tile, batch, k-split).  The accumulator is double-buffered in TMEM (2 x BN columns) so
  // the epilogue warps drain tile i while the MMA warp already accumulates tile i+1.
  float *__restrict__ c = reinterpret_cast<float *>(c_void);
  const int relu = act == 1;
  constexpr int A_TILE = BM * BK * 2;
  constexpr int B_TILE = BN * BK * 2;
  constexpr int STAGE = NSPLIT * (A_TILE + B_TILE);
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;
  // tensor-memory allocations are powers of two: two accumulators of 192 columns take the 512-column allocation
  constexpr uint32_t TMEM_ALLOC = 2 * ACC_COLS <= 32 ? 32 : 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128
                                  : 2 * ACC_COLS <= 256 ? 256 : 512;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  const long long nwork = (long long)tiles_m * tiles_n * batch_n * ksplit;
  const int nkb_total = kpad / BK;
  const int per = (nkb_total + ksplit - 1) / ksplit;

  // work item -> coordinates.  The CTAs that run concurrently should share operand tiles in L2: with few
  // n-tiles (skinny weights, the step's shapes) n runs fastest, so the n-tiles of one m-tile are in flight
  // together and the A rows are fetched from HBM once (ncu: the 1M x 256 x 128 SA layer read A twice when m
  // ran fastest); with more n-tiles than m-tiles the roles swap.
  const bool n_fast = tiles_n <= tiles_m;
  auto decode = [&](long long w, int &m0, int &n0, int &batch, int &ks) {
    int tm, tn;
    if (n_fast) {
      tn = (int)(w % tiles_n); w /= tiles_n;
      tm = (int)(w % tiles_m); w /= tiles_m;
    } else {
      tm = (int)(w % tiles_m); w /= tiles_m;
      tn = (int)(w % tiles_n); w /= tiles_n;
    }
    ks = (int)(w % ksplit);
    batch = (int)(w / ksplit);
    m0 = tm * BM; n0 = tn * BN;
  };

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int p = 0; p < NSPLIT; ++p) { prefetch_tmap(&maps.a[p]); prefetch_tmap(&maps.b[p]); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    mbar_fence_init_cluster();
  }
  if (warp == 2) tmem_alloc(&tmem_slot, TMEM_ALLOC);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (warp-uniform control flow; one elected lane issues the copies) =====
    uint32_t it = 0;  // running k-block counter across work items -> stage / phase
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
      int m0, n0, batch, ks;
      decode(w, m0, n0, batch, ks);
      const int kb0 = ks * per, nkb = max(0, min(per, nkb_total - kb0));
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)STAGE);
          unsigned char *st = smem + (size_t)s * STAGE;
#pragma unroll
          for (int p = 0; p < NSPLIT; ++p) {
            if (!MN) {
              tma_load_3d(st + p * A_TILE, &maps.a[p], &full_bar[s], (kb0 + kb) * BK, m0, batch);
              tma_load_3d(st + NSPLIT * A_TILE + p * B_TILE, &maps.b[p], &full_bar[s], (kb0 + kb) * BK, n0,
                          b_batched ? batch : 0);
            } else {
              // [64 contraction rows x 64 mn] boxes, one per 64-wide slab of the tile
#pragma unroll
              for (int g = 0; g < BM / 64; ++g)
                tma_load_3d(st + p * A_TILE + g * 8192, &maps.a[p], &full_bar[s], m0 + g * 64, (kb0 + kb) * BK, batch);
#pragma unroll
              for (int g = 0; g < BN / 64; ++g)
                tma_load_3d(st + NSPLIT * A_TILE + p * B_TILE + g * 8192, &maps.b[p], &full_bar[s], n0 + g * 64,
                            (kb0 + kb) * BK, b_batched ? batch : 0);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform control flow; one elected lane issues, see elect_one_sync) =====
    constexpr uint32_t idesc = umma_idesc_f16(FP16 ? 1 : 0, BM, BN, MN ? 1 : 0, MN ? 1 : 0);
    uint32_t it = 0, tile_i = 0;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x, ++tile_i) {
      int m0, n0, batch, ks;
      decode(w, m0, n0, batch, ks);
      const int kb0 = ks * per, nkb = max(0, min(per, nkb_total - kb0));
      if (nkb == 0) continue;  // (the epilogue skips it as well)
      const uint32_t buf = tile_i & 1u, use = tile_i >> 1;
      mbar_wait(&acc_empty[buf], (use & 1u) ^ 1u);  // epilogue drained this accumulator buffer
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + buf * ACC_COLS;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          unsigned char *st = smem + (size_t)s * STAGE;
#pragma unroll
          for (int p = 0; p < n_products(NSPLIT); ++p) {
            const void *at = st + prod_a(NSPLIT, p) * A_TILE, *bt = st + NSPLIT * A_TILE + prod_b(NSPLIT, p) * B_TILE;
            const uint64_t ad = MN ? umma_smem_desc_mn_sw128(at) : umma_smem_desc_k_sw128(at);
            const uint64_t bd = MN ? umma_smem_desc_mn_sw128(bt) : umma_smem_desc_k_sw128(bt);
            constexpr uint32_t KSTEP = MN ? 16 * 128 : 32;  // bytes per 16-deep k-step
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk)
              umma_f16(tmem_acc, umma_desc_advance(ad, kk * KSTEP), umma_desc_advance(bd, kk * KSTEP), idesc,
                       (uint32_t)((kb | p | kk) != 0));
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
          if (kb == nkb - 1) umma_commit(&acc_full[buf]);   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp - 4;
    uint32_t tile_i = 0;
    for (long long w = blockIdx.x; w < nwork; w += gridDim.x, ++tile_i) {
      int m0, n0, batch, ks;
      decode(w, m0, n0, batch, ks);
      const int kb0 = ks * per, nkb = max(0, min(per, nkb_total - kb0));
      if (nkb == 0) continue;
      const uint32_t buf = tile_i & 1u, use = tile_i >> 1;
      const int row = m0 + q * 32 + lane;
      const float *bias = ks == 0 ? bias_in : nullptr;
      mbar_wait(&acc_full[buf], use & 1u);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + buf * ACC_COLS;
      float *crow = c + (size_t)batch * c_batch_stride + (size_t)row * ldc;
      if (tma_store) {
        // coalesced epilogue: the warp's [32 rows x 128 B] chunk goes through a 128B-swizzled staging tile and
        // leaves as ONE bulk tensor store (full lines, clipped at the matrix edge, asynchronous)
        unsigned char *stage = smem + (size_t)STAGES * STAGE + (size_t)q * (32 * 128);
        unsigned char *srow = stage + lane * 128;
        const int sw = lane & 7;
        constexpr int CHUNK = FP16 ? 64 : 32;   // output columns per staging tile (fp16 out only with fp16 operands)
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CHUNK) {
          const int col0 = n0 + c0;
          if (col0 >= n) break;                 // warp-uniform
          float v[CHUNK];
          {
            uint32_t r[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
            if (CHUNK == 64) {
              tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + 32), r);
              tmem_ld_wait();
#pragma unroll
              for (int t = 0; t < 32; ++t) v[(CHUNK == 64 ? 32 : 0) + t] = __uint_as_float(r[t]);
            }
          }
          if (bias) {
            if (col0 + CHUNK <= n) {
#pragma unroll
              for (int t = 0; t < CHUNK; t += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + col0 + t));
                v[t] += bb.x; v[t + 1] += bb.y; v[t + 2] += bb.z; v[t + 3] += bb.w;
              }
            } else {
#pragma unroll
              for (int t = 0; t < CHUNK; ++t)
                if (col0 + t < n) v[t] += __ldg(bias + col0 + t);
            }
          }
          if (act == 1) {
#pragma unroll
            for (int t = 0; t < CHUNK; ++t) v[t] = fmaxf(v[t], 0.f);
          } else if (act == 2) {
#pragma unroll
            for (int t = 0; t < CHUNK; ++t) v[t] = __fdividef(v[t], 1.0f + __expf(-1.702f * v[t]));
          }
          if (FP16 && residual != nullptr && row < m) {
            // fused residual connection (CLIP blocks: x + proj(...)): this thread's row, one full 128-byte line
            const __half *rr = reinterpret_cast<const __half *>(residual) + (size_t)row * ldr + col0;
            if (col0 + CHUNK <= n) {
#pragma unroll
              for (int t = 0; t < CHUNK; t += 8) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(rr + t));
                const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float2 f = __half22float2(h2[u]);
                  v[t + 2 * u] += f.x;
                  v[t + 2 * u + 1] += f.y;
                }
              }
            } else {
#pragma unroll
              for (int t = 0; t < CHUNK; ++t)
                if (col0 + t < n) v[t] += __half2float(rr[t]);
            }
          }
          // the previous chunk's store must have read the staging tile before it is rewritten
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {   // eight 16-byte pieces per 128-byte row, XOR-swizzled by the row
            uint4 pk;
            if (FP16) {
              const __half2 h0 = __floats2half2_rn(v[(FP16 ? 8 : 0) * j + 0], v[(FP16 ? 8 : 0) * j + 1]);
              const __half2 h1 = __floats2half2_rn(v[(FP16 ? 8 : 0) * j + 2], v[(FP16 ? 8 : 0) * j + 3]);
              const __half2 h2 = __floats2half2_rn(v[(FP16 ? 8 : 0) * j + 4], v[(FP16 ? 8 : 0) * j + 5]);
              const __half2 h3 = __floats2half2_rn(v[(FP16 ? 8 : 0) * j + 6], v[(FP16 ? 8 : 0) * j + 7]);
              pk.x = *reinterpret_cast<const uint32_t *>(&h0); pk.y = *reinterpret_cast<const uint32_t *>(&h1);
              pk.z = *reinterpret_cast<const uint32_t *>(&h2); pk.w = *reinterpret_cast<const uint32_t *>(&h3);
            } else {
              pk = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                              __float_as_uint(v[4 * j + 3]));
            }
            *reinterpret_cast<uint4 *>(srow + ((j ^ sw) << 4)) = pk;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (tma_store == 2) tma_reduce_add_3d(&maps.c, stage, col0, m0 + q * 32, batch);   // split-K partial
            else tma_store_3d(&maps.c, stage, col0, m0 + q * 32, batch);
            tma_store_commit();
          }
        }
        if (lane == 0) tma_store_wait_read();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (row < m) {
          const int col0 = n0 + c0;
          if (out_half) {
            // fp16 output (CLIP ViT path), optional QuickGELU x * sigmoid(1.702 x)
            __half *hrow = reinterpret_cast<__half *>(c_void) + (size_t)batch * c_batch_stride + (size_t)row * ldc;
            const bool full = col0 + 32 <= n && (ldc & 7) == 0 && ((reinterpret_cast<uintptr_t>(hrow + col0) & 15) == 0);
            if (full) {
#pragma unroll
              for (int j0 = 0; j0 < 32; j0 += 8) {
                float v[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[j0 + t]);
                if (bias) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + col0 + j0));
                  const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + col0 + j0 + 4));
                  v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                  v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                }
                if (act == 1) {
#pragma unroll
                  for (int t = 0; t < 8; ++t) v[t] = fmaxf(v[t], 0.f);
                } else if (act == 2) {
#pragma unroll
                  for (int t = 0; t < 8; ++t) v[t] = __fdividef(v[t], 1.0f + __expf(-1.702f * v[t]));
                }
                uint4 pk;
                __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
                __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
                pk.x = *reinterpret_cast<uint32_t *>(&h0); pk.y = *reinterpret_cast<uint32_t *>(&h1);
                pk.z = *reinterpret_cast<uint32_t *>(&h2); pk.w = *reinterpret_cast<uint32_t *>(&h3);
                *reinterpret_cast<uint4 *>(hrow + col0 + j0) = pk;   // 8 halves = one 16-byte store
              }
            } else {
#pragma unroll 4
              for (int t = 0; t < 32; ++t) {
                const int col = col0 + t;
                if (col < n) {
                  float x = __uint_as_float(r[t]);
                  if (bias) x += __ldg(bias + col);
                  if (act == 1) x = fmaxf(x, 0.f);
                  if (act == 2) x = __fdividef(x, 1.0f + __expf(-1.702f * x));
                  hrow[col] = __float2half_rn(x);
                }
              }
            }
          } else if (ksplit > 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              if (col < n) {
                float v = __uint_as_float(r[j]);
                if (bias) v += __ldg(bias + col);
                atomicAdd(crow + col, v);
              }
            }
          } else if (col0 + 32 <= n && (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(crow + col0) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                     __uint_as_float(r[j + 3]));
              if (bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + col0 + j));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
              }
              if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
              *reinterpret_cast<float4 *>(crow + col0 + j) = v;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              if (col < n) {
                float v = __uint_as_float(r[j]);
                if (bias) v += __ldg(bias + col);
                if (relu) v = fmaxf(v, 0.f);
                crow[col] = v;
              }
            }
          }
        }
      }
      // this warp is done reading the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_ALLOC);
}

template <int NSPLIT, int BN, int STAGES, bool FP16, bool MN = false>
int launch_gemm(const GemmMaps &maps, int batch, int m, int n, int kpad, int b_batched, const float *bias, int relu,
                void *c_out, long long ldc, long long c_batch_stride, cudaStream_t s, int out_half = 0,
                const void *residual = nullptr, long long ldr = 0) {
  float *c = reinterpret_cast<float *>(c_out);
  // split-K when the output has few tiles but the contraction is long (weight gradients)
  const long long tiles = (long long)((m + BM - 1) / BM) * ((n + BN - 1) / BN) * batch;
  const int nkb_total = kpad / BK;
  int ksplit = 1;
  if (!relu && !out_half && tiles < 148 && nkb_total >= 32) {
    ksplit = (int)((2 * 148 + tiles - 1) / tiles);
    if (ksplit > nkb_total / 8) ksplit = nkb_total / 8;
    if (ksplit < 1) ksplit = 1;
  }
  if (ksplit > 1) {
    cudaError_t ez = cudaSuccess;
    for (int bi = 0; bi < batch && ez == cudaSuccess; ++bi)
      ez = cudaMemset2DAsync(c + (size_t)bi * c_batch_stride, (size_t)ldc * 4, 0, (size_t)n * 4, (size_t)m, s);
    if (ez != cudaSuccess) return (int)ez;
  }
  constexpr size_t smem = (size_t)STAGES * NSPLIT * (BM * BK * 2 + BN * BK * 2) + EPI_STAGE + 1024;
  auto kern = gemm_nt_kernel<NSPLIT, BN, STAGES, FP16, MN>;
  // TMA-store epilogue whenever the output satisfies the tensor-map rules (16-byte aligned rows); fp16 output
  // exists only on the fp16-operand instances
  GemmMaps lmaps = maps;
  int tma_store = 0;
  if ((ksplit == 1 || !out_half) && (!out_half || FP16) && (out_half != 0) == FP16) {
    const long long align = out_half ? 8 : 4;
    if (ldc % align == 0 && c_batch_stride % align == 0 && ((uintptr_t)c_out & 15) == 0) {
      const int st = out_half ? make_tmap_k_major_16b(&lmaps.c, c_out, 1, n, m, batch, ldc, c_batch_stride, 32)
                              : make_tmap_rows_f32(&lmaps.c, c_out, n, m, batch, ldc, c_batch_stride);
      tma_store = st == CODA_OK ? (ksplit > 1 ? 2 : 1) : 0;   // split-K: bulk reduce-add into the zeroed output
    }
  }
  static bool configured = false;  // once per template instance
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const long long nwork = tiles * ksplit;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const unsigned grid = (unsigned)(nwork < num_sms ? nwork : num_sms);
  if (residual && (tma_store != 1 || !FP16 || batch != 1 || (ldr & 7) != 0 || ((uintptr_t)residual & 15) != 0))
    return CODA_EINVAL;   // the fused residual exists on the fp16 TMA-store epilogue only
  kern<<<grid, 256, smem, s>>>(lmaps, m, n, kpad, b_batched, ksplit, batch, bias, relu, out_half, c_out, ldc,
                               c_batch_stride, tma_store, residual, ldr);
  return launch_status();
}

}  // namespace

extern "C" {

int coda_pack_split_bf16_strided(long long rows, int k, int kpad, long long src_row_stride,
                                 long long src_k_stride, const float *src, float scale, int nsplit, void *planes,
                                 long long plane_stride, void *stream) {
  if (rows < 0 || k < 0 || kpad < k || kpad % 64 != 0 || nsplit < 1 || nsplit > 3) return CODA_EINVAL;
  if (rows == 0 || kpad == 0) return CODA_OK;
  if (!src || !planes) return CODA_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  __nv_bfloat16 *out = (__nv_bfloat16 *)planes;
  if ((src_k_stride == 1 || k <= 1) && k % 4 == 0 && src_row_stride % 4 == 0 && ((uintptr_t)src & 15) == 0 &&
      plane_stride % 4 == 0 && ((uintptr_t)out & 7) == 0) {
    const long long total = rows * (kpad / 4);
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (nsplit == 1) pack_rows_vec4_kernel<1><<<grid, 256, 0, s>>>(rows, k, kpad, src_row_stride, src, scale, out, plane_stride);
    else if (nsplit == 2) pack_rows_vec4_kernel<2><<<grid, 256, 0, s>>>(rows, k, kpad, src_row_stride, src, scale, out, plane_stride);
    else pack_rows_vec4_kernel<3><<<grid, 256, 0, s>>>(rows, k, kpad, src_row_stride, src, scale, out, plane_stride);
  } else if (src_k_stride == 1 || k <= 1) {
    const long long total = rows * kpad;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (nsplit == 1) pack_rows_kernel<1><<<grid, 256, 0, s>>>(rows, k, kpad, src_row_stride, src, scale, out, plane_stride);
    else if (nsplit == 2) pack_rows_kernel<2><<<grid, 256, 0, s>>>(rows, k, kpad, src_row_stride, src, scale, out, plane_stride);
    else pack_rows_kernel<3><<<grid, 256, 0, s>>>(rows, k, kpad, src_row_stride, src, scale, out, plane_stride);
  } else if (src_row_stride == 1) {
    const dim3 grid((unsigned)((rows + 31) / 32), (kpad + 31) / 32);
    if (grid.y > 65535) return CODA_ETOOLARGE;
    if (nsplit == 1) pack_transposed_kernel<1><<<grid, 256, 0, s>>>(rows, k, kpad, src_k_stride, src, scale, out, plane_stride);
    else if (nsplit == 2) pack_transposed_kernel<2><<<grid, 256, 0, s>>>(rows, k, kpad, src_k_stride, src, scale, out, plane_stride);
    else pack_transposed_kernel<3><<<grid, 256, 0, s>>>(rows, k, kpad, src_k_stride, src, scale, out, plane_stride);
  } else {
    return CODA_EINVAL;  // one of the two strides must be 1
  }
  return launch_status();
}

int coda_pack_split_bf16(long long rows, int k, int kpad, long long src_row_stride, long long src_k_stride,
                         const float *src, float scale, int nsplit, void *planes, void *stream) {
  return coda_pack_split_bf16_strided(rows, k, kpad, src_row_stride, src_k_stride, src, scale, nsplit, planes,
                                      rows * kpad, stream);
}

int coda_gemm_nt(int nsplit, int is_fp16, int batch, int m, int n, int kpad, const void *a,
                 long long a_plane_stride, long long a_batch_stride, const void *b, long long b_plane_stride,
                 long long b_batch_stride, const float *bias, int relu, float *c, long long ldc,
                 long long c_batch_stride, void *stream) {
  return coda_gemm_nt_ex(nsplit, is_fp16, batch, m, n, kpad, a, a_plane_stride, a_batch_stride, b, b_plane_stride,
                         b_batch_stride, bias, relu, 0, c, ldc, c_batch_stride, stream);
}

int coda_gemm_nt_ex(int nsplit, int is_fp16, int batch, int m, int n, int kpad, const void *a,
                    long long a_plane_stride, long long a_batch_stride, const void *b, long long b_plane_stride,
                    long long b_batch_stride, const float *bias, int relu, int out_half, void *c, long long ldc,
                    long long c_batch_stride, void *stream) {
  return coda_gemm_nt_res(nsplit, is_fp16, batch, m, n, kpad, a, a_plane_stride, a_batch_stride, b, b_plane_stride,
                          b_batch_stride, bias, relu, out_half, nullptr, 0, c, ldc, c_batch_stride, stream);
}

int coda_gemm_nt_res(int nsplit, int is_fp16, int batch, int m, int n, int kpad, const void *a,
                     long long a_plane_stride, long long a_batch_stride, const void *b, long long b_plane_stride,
                     long long b_batch_stride, const float *bias, int relu, int out_half, const void *residual,
                     long long ldr, void *c, long long ldc, long long c_batch_stride, void *stream) {
  if (nsplit < 1 || nsplit > 3 || batch < 0 || m < 0 || n < 0 || kpad < 0 || kpad % 64 != 0) return CODA_EINVAL;
  if (is_fp16 && nsplit != 1) return CODA_EINVAL;
  if (batch == 0 || m == 0 || n == 0) return CODA_OK;
  if (!a || !b || !c || kpad == 0 || batch > 65535) return CODA_EINVAL;
  // fp16 tower GEMMs with wide outputs: a 128 x 256 tile moves 48 KB per k-block for twice the flops of a 128 x 128
  // tile (32 KB) -- these shapes are L2 -> SM bandwidth bound, not tensor bound.  Narrow outputs (N = 768: 300 tiles
  // on 148 SMs) keep the 128-wide tile for its finer wave quantisation.
  // N = 768 (attention / MLP output projections of the ViT): 192-wide tiles give 400 work items (2.7 waves) at
  // 77 flop/B instead of 600 (4.05 waves -> 5 rounds) at 64 flop/B.
  const int bn = n <= 64 ? 64
                 : (is_fp16 && n % 256 == 0 && n >= 1536) ? 256
                 : (is_fp16 && n % 192 == 0 && n >= 384)  ? 192
                                                          : 128;
  GemmMaps maps;
  const char *ap = (const char *)a, *bp = (const char *)b;
  for (int p = 0; p < nsplit; ++p) {
    int st = make_tmap_k_major_16b(&maps.a[p], ap + (size_t)p * a_plane_stride * 2, is_fp16, kpad, m, batch, kpad,
                                   a_batch_stride, BM);
    if (st != CODA_OK) return st;
    st = make_tmap_k_major_16b(&maps.b[p], bp + (size_t)p * b_plane_stride * 2, is_fp16, kpad, n,
                               b_batch_stride ? batch : 1, kpad, b_batch_stride, bn);
    if (st != CODA_OK) return st;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int bb = b_batch_stride ? 1 : 0;
#define CODA_GEMM(NS, BN_, ST, F16) \
  return launch_gemm<NS, BN_, ST, F16>(maps, batch, m, n, kpad, bb, bias, relu, c, ldc, c_batch_stride, s, out_half, \
                                       residual, ldr)
  if (is_fp16) {
    if (bn == 64) CODA_GEMM(1, 64, 6, true);
    if (bn == 256) CODA_GEMM(1, 256, 4, true);
    if (bn == 192) CODA_GEMM(1, 192, 5, true);
    CODA_GEMM(1, 128, 6, true);
  }
  if (nsplit == 1) {
    if (bn == 64) CODA_GEMM(1, 64, 6, false);
    CODA_GEMM(1, 128, 6, false);
  }
  if (nsplit == 2) {
    if (bn == 64) CODA_GEMM(2, 64, 4, false);
    CODA_GEMM(2, 128, 3, false);
  }
  if (bn == 64) CODA_GEMM(3, 64, 2, false);
  CODA_GEMM(3, 128, 2, false);
#undef CODA_GEMM
}


int coda_gemm_tn(int nsplit, int mc, int m, int n, const void *a, long long a_plane_stride, int lda,
                 const void *b, long long b_plane_stride, int ldb, float *c, long long ldc, void *stream) {
  // C[m][n] = sum_r A[r][m] * B[r][n], r < mc;  A planes [nsplit][mc][lda], B planes [nsplit][mc][ldb]
  if (nsplit < 1 || nsplit > 3 || mc < 0 || m < 0 || n < 0 || lda % 64 != 0 || ldb % 64 != 0) return CODA_EINVAL;
  if (m == 0 || n == 0) return CODA_OK;
  if (!a || !b || !c || mc == 0 || lda < m || ldb < n) return CODA_EINVAL;
  const int bn = n <= 64 ? 64 : 128;
  GemmMaps maps;
  const char *ap = (const char *)a, *bp = (const char *)b;
  for (int p = 0; p < nsplit; ++p) {
    // tensors [1][mc rows][lda cols]; box = [64 rows][64 cols]
    int st = make_tmap_k_major_16b(&maps.a[p], ap + (size_t)p * a_plane_stride * 2, 0, lda, mc, 1, lda, 0, 64);
    if (st != CODA_OK) return st;
    st = make_tmap_k_major_16b(&maps.b[p], bp + (size_t)p * b_plane_stride * 2, 0, ldb, mc, 1, ldb, 0, 64);
    if (st != CODA_OK) return st;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int kpad = (mc + 63) / 64 * 64;  // contraction length in 64-row blocks (rows >= mc read as zero)
#define CODA_GEMM_TN(NS, BN_, ST) \
  return launch_gemm<NS, BN_, ST, false, true>(maps, 1, m, n, kpad, 0, nullptr, 0, c, ldc, 0, s)
  if (nsplit == 1) { if (bn == 64) CODA_GEMM_TN(1, 64, 6); CODA_GEMM_TN(1, 128, 6); }
  if (nsplit == 2) { if (bn == 64) CODA_GEMM_TN(2, 64, 4); CODA_GEMM_TN(2, 128, 3); }
  if (bn == 64) CODA_GEMM_TN(3, 64, 2);
  CODA_GEMM_TN(3, 128, 2);
#undef CODA_GEMM_TN
}

}  // extern "C"
