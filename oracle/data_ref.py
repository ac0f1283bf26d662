"""TEST INFRASTRUCTURE (checker only; nothing in the package imports this).

CPU restatement, in numpy, of the device data layer (coda_neurips2023_b200/csrc/data_kernels.cu, include/coda_data.h)
following the reference's per-scene pipeline line by line:

  * scene_transform   datasets/sunrgbd_anonymous_aligned_image.py:660-705 (flip about YZ, rotz, scale) on float32
  * random_cuboid     utils/random_cuboid.py:39-116 with the random draws taken from a TABLE (crop_range per attempt,
                      centre-point selector u per attempt) instead of np.random -- pinned by
                      tests/test_data_cpu.py against the reference's own RandomCuboid fed the same draws
  * sample_points     utils/pc_util.py:24-32 random_sampling; the permutation is the keyed Feistel bijection of the
                      kernel (a different -- but equally uniform -- random stream than np.random.choice)
  * image_augment     datasets/...:624-655
"""
from __future__ import annotations

import numpy as np

M32 = 0xFFFFFFFF


def mix32d(h):
    h = np.asarray(h, dtype=np.uint64) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x7FEB352D)) & M32
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x846CA68B)) & M32
    h ^= h >> np.uint64(16)
    return h


def scene_transform(points: np.ndarray, flip: float, rot: np.ndarray, scale: float) -> np.ndarray:
    """points (n, stride) float32 -> transformed copy; float32 products / sums rounded one by one, left to right"""
    p = points.astype(np.float32).copy()
    R = rot.astype(np.float32)
    x = p[:, 0] * np.float32(flip)
    y, z = p[:, 1].copy(), p[:, 2].copy()
    for j in range(3):
        v = (x * R[j, 0] + y * R[j, 1]) + z * R[j, 2]
        p[:, j] = v * np.float32(scale)
    return p


def check_aspect(crop_range, aspect_min):          # utils/random_cuboid.py:5-13
    xy = min(crop_range[0], crop_range[1]) / max(crop_range[0], crop_range[1])
    xz = min(crop_range[0], crop_range[2]) / max(crop_range[0], crop_range[2])
    yz = min(crop_range[1], crop_range[2]) / max(crop_range[1], crop_range[2])
    return xy >= aspect_min or xz >= aspect_min or yz >= aspect_min


def center_index(u: float, n: int) -> int:
    i = int(np.float32(u) * np.float32(n))
    return min(max(i, 0), n - 1)


def random_cuboid(points, boxes, crop_range, center_u, min_points, aspect=0.8):
    """points (n, >=3) float32, boxes (g, >=3) float32 rows [cx, cy, cz, ...], crop_range (ncand, 3) float64,
    center_u (ncand,) float32 -> (chosen attempt or -1, crop bounds (6,) float64, box_keep (g,) bool)"""
    xyz = points[:, 0:3]
    n = len(xyz)
    range_xyz = np.max(xyz, axis=0) - np.min(xyz, axis=0)          # float32
    has_boxes = boxes.size > 0 and np.float32(boxes.astype(np.float32).sum()) > 0
    for c in range(len(crop_range)):
        cr = crop_range[c]
        if not check_aspect(cr, np.float64(np.float32(aspect))):
            continue
        center = xyz[center_index(center_u[c], n)]
        new_range = range_xyz * cr / 2.0                               # float64
        max_xyz, min_xyz = center + new_range, center - new_range
        inside = (np.sum((xyz <= max_xyz).astype(np.int32), 1) == 3) & (np.sum((xyz >= min_xyz).astype(np.int32), 1) == 3)
        if np.sum(inside) < min_points:
            continue
        keep = np.ones(len(boxes), bool)
        if has_boxes:
            kept = xyz[inside]
            lo, hi = np.min(kept, axis=0), np.max(kept, axis=0)
            keep = np.logical_and(np.all(boxes[:, 0:3] >= lo, axis=1), np.all(boxes[:, 0:3] <= hi, axis=1))
            if keep.sum() == 0:
                continue
        return c, np.concatenate((min_xyz, max_xyz)).astype(np.float64), keep
    return -1, np.array([-np.inf] * 3 + [np.inf] * 3), np.ones(len(boxes), bool)


def feistel(x, half_bits, key):
    mask = np.uint64((1 << half_bits) - 1)
    x = np.asarray(x, dtype=np.uint64)
    left, right = x >> np.uint64(half_bits), x & mask
    for rnd in range(4):
        f = mix32d((right * np.uint64(0x9E3779B1) + np.uint64(key) + np.uint64(rnd) * np.uint64(0x85EBCA6B)) & M32) & mask
        left, right = right, left ^ f
    return (left << np.uint64(half_bits)) | right


def sample_points(points, crop, seed, nsample):
    """-> (sampled (nsample, stride), choice (nsample,) rows of `points`, count inside, dims (6,))"""
    xyz = points[:, 0:3].astype(np.float64)
    inside = np.all(xyz <= crop[3:], axis=1) & np.all(xyz >= crop[:3], axis=1)
    lst = np.nonzero(inside)[0]
    m = len(lst)
    if m == 0:
        return np.zeros((nsample, points.shape[1]), np.float32), -np.ones(nsample, np.int64), 0, None
    key = int(mix32d(np.uint64(seed) ^ np.uint64(0xA511E9B3)))
    i = np.arange(nsample, dtype=np.uint64)
    if m >= nsample:
        half_bits = 1
        while (1 << (2 * half_bits)) < m:
            half_bits += 1
        j = feistel(i, half_bits, key)
        while True:
            out = j >= m
            if not out.any():
                break
            j[out] = feistel(j[out], half_bits, key)
    else:
        j = mix32d((i * np.uint64(0x9E3779B1) + np.uint64(key)) & M32) % np.uint64(m)
    choice = lst[j.astype(np.int64)]
    out = points[choice].astype(np.float32)
    dims = np.concatenate((out[:, :3].min(axis=0), out[:, :3].max(axis=0)))
    return out, choice, m, dims


def image_augment(img, flip, gain, shift, seed):
    """img (h, w, 3) uint8 -> uint8, float32 arithmetic in the kernel's order"""
    h, w, _ = img.shape
    src = img[:, ::-1, :] if flip else img
    t = np.arange(h * w, dtype=np.uint64)
    r = mix32d((np.uint64(seed) + t * np.uint64(0x9E3779B1)) & M32)
    u = (r >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    jit = (np.float32(0.05) * u - np.float32(0.025)).reshape(h, w, 1)
    v = src.astype(np.float32) / np.float32(255.0)
    v = v * gain.astype(np.float32).reshape(1, 1, 3)
    v = v + shift.astype(np.float32).reshape(1, 1, 3)
    v = v + jit
    v = np.clip(v, np.float32(0), np.float32(1))
    return (v * np.float32(255.0)).astype(np.uint8)
