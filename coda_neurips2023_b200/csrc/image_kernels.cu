// Batched crop + white-pad + antialiased bicubic resize + normalise for the CLIP
// crops (include/coda_image.h).  One thread per output pixel (all 3 channels);
// the source canvas is virtual: pixels outside the pasted crop read as 255.
#include <cuda_fp16.h>
#include <math.h>

#include "../../include/coda_image.h"
#include "coda_common.cuh"

using namespace coda;

namespace {

constexpr int MAX_TAPS = 40;

// ATen upsample_antialias: cubic convolution filter with a = -0.5
__device__ __forceinline__ float cubic_aa(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
  if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
  return 0.0f;
}

// _compute_indices_span + _compute_weights of ATen/native/cuda/UpSample.cuh
__device__ __forceinline__ void aa_weights(float scale, int out_idx, int in_size, int &lo, int &n, float *w) {
  const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
  const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  const float center = scale * (out_idx + 0.5f);
  lo = max((int)(center - support + 0.5f), 0);
  n = min((int)(center + support + 0.5f), in_size) - lo;
  n = min(n, MAX_TAPS);
  float total = 0.f;
  for (int j = 0; j < n; ++j) {
    const float v = cubic_aa((j + lo - center + 0.5f) * invscale);
    w[j] = v;
    total += v;
  }
  if (total != 0.f)
    for (int j = 0; j < n; ++j) w[j] /= total;
}

template <typename OutT>
__global__ void __launch_bounds__(256)
crop_resize_kernel(int h, int w, int res, const unsigned char *__restrict__ images,
                   const int *__restrict__ scene, const int *__restrict__ boxes,
                   const unsigned char *__restrict__ valid, float m0, float m1, float m2, float s0,
                   float s1, float s2, OutT *__restrict__ out) {
  const int crop = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= res * res) return;
  OutT *o = out + (size_t)crop * 3 * res * res + pix;
  const size_t plane = (size_t)res * res;
  if (!valid[crop]) {
    o[0] = OutT(0.f); o[plane] = OutT(0.f); o[2 * plane] = OutT(0.f);
    return;
  }
  const int xmin = boxes[crop * 4 + 0], ymin = boxes[crop * 4 + 1];
  const int wc = boxes[crop * 4 + 2] - xmin, hc = boxes[crop * 4 + 3] - ymin;
  const int edge = max(wc, hc);
  const int y_begin = (edge - hc) / 2, x_begin = (edge - wc) / 2;
  const unsigned char *img = images + (size_t)scene[crop] * h * w * 3;
  const float scale = (float)edge / (float)res;
  const int oy = pix / res, ox = pix - oy * res;
  float wy[MAX_TAPS], wx[MAX_TAPS];
  int ylo, yn, xlo, xn;
  aa_weights(scale, oy, edge, ylo, yn, wy);
  aa_weights(scale, ox, edge, xlo, xn, wx);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  for (int iy = 0; iy < yn; ++iy) {
    const int cy = ylo + iy - y_begin;            // row inside the crop
    const bool yin = cy >= 0 && cy < hc;
    const unsigned char *row = img + ((size_t)(ymin + cy) * w + xmin) * 3;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    for (int ix = 0; ix < xn; ++ix) {
      const int cx = xlo + ix - x_begin;
      float p0 = 255.f, p1 = 255.f, p2 = 255.f;   // white canvas
      if (yin && cx >= 0 && cx < wc) {
        const unsigned char *p = row + (size_t)cx * 3;
        p0 = (float)p[0]; p1 = (float)p[1]; p2 = (float)p[2];
      }
      r0 += wx[ix] * p0; r1 += wx[ix] * p1; r2 += wx[ix] * p2;
    }
    acc0 += wy[iy] * r0; acc1 += wy[iy] * r1; acc2 += wy[iy] * r2;
  }
  // back to what the uint8 tensor would hold, then /255 and Normalize
  acc0 = rintf(fminf(fmaxf(acc0, 0.f), 255.f));
  acc1 = rintf(fminf(fmaxf(acc1, 0.f), 255.f));
  acc2 = rintf(fminf(fmaxf(acc2, 0.f), 255.f));
  o[0] = OutT((acc0 / 255.0f - m0) / s0);
  o[plane] = OutT((acc1 / 255.0f - m1) / s1);
  o[2 * plane] = OutT((acc2 / 255.0f - m2) / s2);
}

}  // namespace

extern "C" int coda_crop_resize_normalize(int nimg, int h, int w, int ncrops, int res,
                                          const unsigned char *images, const int *scene,
                                          const int *boxes, const unsigned char *valid,
                                          const float *mean, const float *std, int out_half,
                                          void *out, void *stream) {
  if (nimg < 0 || h <= 0 || w <= 0 || ncrops < 0 || res <= 0) return CODA_EINVAL;
  if (ncrops == 0) return CODA_OK;
  if (!images || !scene || !boxes || !valid || !mean || !std || !out || ncrops > 65535) return CODA_EINVAL;
  // taps per axis = 2 * ceil(2 * scale) + 1 with scale <= max(h, w) / res
  const float max_scale = (float)(h > w ? h : w) / (float)res;
  if (2 * (int)ceilf(2.0f * (max_scale > 1.f ? max_scale : 1.f)) + 1 > MAX_TAPS) return CODA_ETOOLARGE;
  const dim3 grid((res * res + 255) / 256, ncrops);
  cudaStream_t s = (cudaStream_t)stream;
  if (out_half)
    crop_resize_kernel<__half><<<grid, 256, 0, s>>>(h, w, res, images, scene, boxes, valid, mean[0], mean[1],
                                                   mean[2], std[0], std[1], std[2], (__half *)out);
  else
    crop_resize_kernel<float><<<grid, 256, 0, s>>>(h, w, res, images, scene, boxes, valid, mean[0], mean[1],
                                                  mean[2], std[0], std[1], std[2], (float *)out);
  return launch_status();
}
