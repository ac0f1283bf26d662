// Blackwell (sm_100a) building blocks shared by the tensor-core kernels:
// TMA tensor-map encoding (host), TMA loads, tcgen05 MMA / commit / ld, TMEM
// allocation, UMMA shared-memory and instruction descriptors.  Inline PTX only.
#pragma once
#include <cuda.h>  // CUtensorMap types (header only; the driver entry point is fetched at run time)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "coda_common.cuh"

namespace coda {

// --------------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  // cuTensorMapEncodeTiled validates the global address against the CURRENT context.  An entry point that encodes a
  // map before its first runtime call can be the first CUDA call of its thread (autograd's backward worker): bind the
  // primary context to the thread once.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 16-bit matrix tiles, K contiguous: tensor [batch][rows][k] with strides (elements),
// box = [1][box_rows][64] (64 x 2 B = one 128-byte swizzle span).
inline int make_tmap_k_major_16b(CUtensorMap *map, const void *base, int is_fp16, long long k, long long rows,
                                 long long batch, long long row_stride, long long batch_stride, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return CODA_EINVAL;
  cuuint64_t gdim[3] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)(batch > 1 ? batch_stride : row_stride * rows) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, is_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                  const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CODA_OK : CODA_EINVAL;
}

// fp32 output tiles for TMA stores: tensor [batch][rows][cols], box = [1][32][32] (32 x 4 B = one swizzle span)
inline int make_tmap_rows_f32(CUtensorMap *map, const void *base, long long cols, long long rows, long long batch,
                              long long row_stride, long long batch_stride) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return CODA_EINVAL;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)(batch > 1 ? batch_stride : row_stride * rows) * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CODA_OK : CODA_EINVAL;
}

// --------------------------------------------------------------------- device: TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 1-D bulk copy global -> shared (size and both addresses multiples of 16 bytes), completing on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shared -> global tile store (bulk async group of the issuing thread); out-of-range rows / columns are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// same, but the tile is ADDED to global memory (fp32): split-K partial sums without per-element atomics
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the shared-memory source of every committed store of this thread has been read (it may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// --------------------------------------------------------------------- device: tcgen05
// One lane of a converged warp.  The MMA-issuing warp runs its control flow warp-uniformly and predicates
// only the tcgen05 instructions on this: operands then stay in uniform registers.  (Under `if (lane == 0)`
// the compiler treats the region as divergent and wraps every UTCHMMA in an ELECT / BRA.U.ANY loop.)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp; writes the TMEM base address into *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16 and bf16 operands
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (e.g. softmax probabilities written back with tcgen05.st)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// --------------------------------------------------------------------- descriptors
// K-major operand tile in shared memory, 128-byte swizzle: rows of 64 16-bit
// elements (128 B) stored densely, 8-row groups 1024 B apart (SBO), tile base
// 1024-byte aligned (cute UMMA::SmemDescriptor, mma_sm100_desc.hpp).
__device__ __forceinline__ uint64_t umma_smem_desc_k_sw128(const void *tile) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(tile) & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                             // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B
  return d;
}
// MN-major operand tile (the contraction index is the ROW index of the stored tensor): the tile is
// a sequence of [64 k-rows x 64 mn] boxes (8 KB each, 128B-swizzled rows of 64 contiguous MN
// elements).  Canonical layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in elements (cute
// mma_traits_sm100.hpp): 8-row K groups SBO = 1024 B apart, 64-wide MN groups LBO = 8192 B apart.
__device__ __forceinline__ uint64_t umma_smem_desc_mn_sw128(const void *tile) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(tile) & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;                   // leading byte offset: next 64 MN elements
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: next 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

// advance the start address by `bytes` inside the swizzle atom (k-steps of 32 B)
__device__ __forceinline__ uint64_t umma_desc_advance(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

// instruction descriptor for kind::f16: fp32 accumulate, K-major A and B
__host__ __device__ constexpr uint32_t umma_idesc_f16(int is_fp16, int m, int n, int a_mn = 0, int b_mn = 0) {
  return ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16)  // operand majors: 0 = K-major, 1 = MN-major
         | (1u << 4)                                 // c_format = F32
         | ((is_fp16 ? 0u : 1u) << 7)              // a_format (0 = F16, 1 = BF16)
         | ((is_fp16 ? 0u : 1u) << 10)             // b_format
         | ((uint32_t)(n >> 3) << 17)              // n_dim
         | ((uint32_t)(m >> 4) << 24);             // m_dim
}

// byte offset of element (row, col16B chunk) inside a K-major SW128 tile: the 16-byte
// chunk index is XOR-ed with (row mod 8)
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ (row & 7u)) << 4);
}

}  // namespace coda
