/*
 * coda_data.h -- C-ABI of the device-side data layer: the per-scene numpy pipeline of
 * the reference's dataset __getitem__ (datasets/sunrgbd_anonymous_aligned_image.py:618-795 and its ScanNet twin,
 * utils/random_cuboid.py, utils/pc_util.py:24-32) for a batch of raw scenes resident in HBM.  Conventions as in
 * coda_pointnet2.h.  Scenes are padded: points (b, nmax, stride) fp32 with npts (b) valid rows (columns 0-2 = xyz,
 * the others -- colour, height -- travel along); boxes (b, gmax, box_stride) fp32 rows [cx, cy, cz, ...] with
 * nbox (b) valid rows.  All randomness comes in as small device arrays drawn by the caller.
 */
#ifndef CODA_DATA_H
#define CODA_DATA_H

#include "coda_pointnet2.h" /* status codes */

#ifdef __cplusplus
extern "C" {
#endif

/*
 * In place: xyz <- ((flip * x, y, z) @ rot^T) * scale with flip (b) = +-1, rot (b, 3, 3), scale (b)
 *   replaces :663-700 (flip about the YZ plane, pc_util.rotz rotation, 0.85-1.15 scaling) on the points; float32
 *   products / sums rounded separately in numpy's order.
 */
int coda_scene_transform(int b, int nmax, int stride, const int *npts, const float *flip, const float *rot,
                         const float *scale, float *points, void *stream);

/* dims (b, 6) = [min xyz | max xyz] of the valid points (npts may be NULL: all nmax rows). */
int coda_points_extent(int b, int nmax, int stride, const int *npts, const float *points, float *dims, void *stream);

/*
 * RandomCuboid (utils/random_cuboid.py:39-116) with all `ncand` attempts of a scene evaluated at once.
 *   range_xyz (b, 3) = extent of the cloud; crop_range (b, ncand, 3) fp64 in [min_crop, max_crop] (the reference's
 *   random numbers are doubles and so are the bounds they produce); center_u (b, ncand) in [0, 1) selects the centre
 *   point floor(u * npts).  Attempt c crops to centre +- range_xyz * crop_range / 2 (inclusive, evaluated in fp64).  chosen (b) = first attempt with (i) an aspect ratio >= aspect_min in some plane, (ii) >= min_points
 *   points inside, (iii) -- if the scene has ground truth -- at least one box centre within the extent of the points
 *   inside; -1 = none (the scene is kept whole).  crop (b, 6) = the chosen bounds (+-inf when -1);
 *   box_keep (b, gmax) = boxes that stay (centre inside that extent).  stats_scratch: b * ncand * 8 floats.
 */
int coda_random_cuboid(int b, int nmax, int stride, int ncand, int gmax, int box_stride, int min_points,
                       float aspect_min, const int *npts, const float *points, const float *range_xyz,
                       const double *crop_range, const float *center_u, const float *boxes, const int *nbox,
                       float *stats_scratch, int *chosen, double *crop, unsigned char *box_keep, void *stream);

/*
 * pc_util.random_sampling (:24-32) of the points inside crop (b, 6): out (b, nsample, stride), choice (b, nsample)
 *   = row of the raw scene each sample came from, count (b) = points inside, dims (b, 6) = extent of the sample
 *   (point_cloud_dims_min / max, :748-749).  Without replacement when count >= nsample (a keyed Feistel permutation
 *   with cycle walking instead of a sort), hashed draws with replacement otherwise.  list_scratch: b * nmax ints.
 */
int coda_sample_points(int b, int nmax, int stride, int nsample, const int *npts, const float *points,
                       const double *crop, const unsigned int *seed, int *list_scratch, int *count, float *out,
                       int *choice, float *dims, void *stream);

/*
 * Image augmentation of :624-655 on uint8 HWC images: horizontal flip (flip (b) != 0), per-channel gain (b, 3) and
 * shift (b, 3), per-pixel jitter in [-0.025, 0.025) from a counter hash of seed (b), clip to [0, 1], truncation to
 * uint8.  `out` must not alias `in`.
 */
int coda_image_augment(int b, int h, int w, const unsigned char *in, const unsigned char *flip, const float *gain,
                       const float *shift, const unsigned int *seed, unsigned char *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CODA_DATA_H */
