/*
 * coda_gemm.h -- C-ABI of the tcgen05 / TMA GEMM used for every dense layer on
 * the path (Linear / 1x1-Conv of the 3DETR encoder, decoder and heads, the SA
 * shared MLP, and the CLIP ViT blocks).  In the reference these are
 * nn.Linear / nn.Conv1d / nn.Conv2d(1x1) calls into cuBLAS / cuDNN
 * (models/helpers.py:45-112, models/transformer.py, pytorch_utils.py:8-33,
 * CLIP/clip/model.py:295-316).
 *
 * The GEMM is "NT": C[b][m][n] = sum_k A[b][m][k] * B[b][n][k] (+ bias[n]) (ReLU),
 * both operands K-contiguous 16-bit planes whose K extent is padded to a multiple
 * of 64.  fp32 tensors are first split into nsplit bf16 planes by
 * coda_pack_split_bf16 (x = p0 + p1 (+ p2), each the bf16 rounding of what the
 * previous planes left over); the GEMM accumulates the 1 / 3 / 6 significant
 * cross products in fp32 in tensor memory.  nsplit = 2 carries ~16 mantissa
 * bits per operand, nsplit = 3 the full 24.
 */
#ifndef CODA_GEMM_H
#define CODA_GEMM_H

#ifdef __cplusplus
extern "C" {
#endif

/*
 * src: fp32 matrix of `rows` x `k` addressed as src[r * src_row_stride + c * src_k_stride]
 *      (one of the two strides must be 1: plain or transposed source); every element is
 *      multiplied by `scale` before splitting.
 * planes: bf16 [nsplit][rows][kpad], zero-padded for k <= c < kpad.  kpad % 64 == 0.
 */
int coda_pack_split_bf16(long long rows, int k, int kpad, long long src_row_stride,
                         long long src_k_stride, const float *src, float scale, int nsplit,
                         void *planes, void *stream);

/* Same, with an explicit distance (elements) between consecutive planes, so that several
 * batch entries can be packed into one [nsplit][batch][rows][kpad] buffer. */
int coda_pack_split_bf16_strided(long long rows, int k, int kpad, long long src_row_stride,
                                 long long src_k_stride, const float *src, float scale, int nsplit,
                                 void *planes, long long plane_stride, void *stream);

/*
 * a: [nsplit planes][batch][m][kpad], plane / batch strides in ELEMENTS; b likewise with n rows;
 * b_batch_stride == 0 means one B shared by every batch entry (weights).
 * is_fp16: operands are IEEE fp16 instead of bf16 (requires nsplit == 1).
 * c: fp32, row stride ldc, batch stride c_batch_stride (elements).  bias may be NULL.
 */
int coda_gemm_nt(int nsplit, int is_fp16, int batch, int m, int n, int kpad, const void *a,
                 long long a_plane_stride, long long a_batch_stride, const void *b,
                 long long b_plane_stride, long long b_batch_stride, const float *bias, int relu,
                 float *c, long long ldc, long long c_batch_stride, void *stream);

/* Extended form: `act` 0 none / 1 ReLU / 2 QuickGELU (x * sigmoid(1.702 x), CLIP/clip/model.py:263-265);
 * out_half != 0 writes C as IEEE fp16 (the CLIP ViT path keeps activations in fp16). */
int coda_gemm_nt_ex(int nsplit, int is_fp16, int batch, int m, int n, int kpad, const void *a,
                    long long a_plane_stride, long long a_batch_stride, const void *b,
                    long long b_plane_stride, long long b_batch_stride, const float *bias, int act,
                    int out_half, void *c, long long ldc, long long c_batch_stride, void *stream);

/* As coda_gemm_nt_ex, plus a fused residual connection on the fp16 path: C = act(A B^T + bias) + residual,
 * residual (m, n) IEEE half with row stride ldr (multiple of 8), batch == 1, fp16 operands and output
 * (the `x + proj(...)` of the CLIP residual blocks, CLIP/clip/model.py:283-288).  NULL = no residual. */
int coda_gemm_nt_res(int nsplit, int is_fp16, int batch, int m, int n, int kpad, const void *a,
                     long long a_plane_stride, long long a_batch_stride, const void *b,
                     long long b_plane_stride, long long b_batch_stride, const float *bias, int act,
                     int out_half, const void *residual, long long ldr, void *c, long long ldc,
                     long long c_batch_stride, void *stream);

/*
 * "TN" form for weight gradients: C[m][n] = sum_{r < mc} A[r][m] * B[r][n], with A planes
 * [nsplit][mc][lda] and B planes [nsplit][mc][ldb] (row-major, lda / ldb multiples of 64), i.e. the
 * contraction runs over ROWS of both stored operands (MN-major tensor-core operands): dW = dY^T X
 * straight from the row-packed dY and X, no transposed copies.  Split-K over mc when m x n is small.
 */
int coda_gemm_tn(int nsplit, int mc, int m, int n, const void *a, long long a_plane_stride, int lda,
                 const void *b, long long b_plane_stride, int ldb, float *c, long long ldc, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CODA_GEMM_H */
