"""Seeded synthetic SUN RGB-D / ScanNet-shaped inputs.

There is no dataset offline; every test and the benchmark draw their scenes
from here.  Shapes and dtypes follow what the reference's dataloader collates
(datasets/sunrgbd_anonymous_aligned_image.py:813-899).
"""
from __future__ import annotations

import numpy as np

ROOM_MIN = np.array([-3.0, 0.5, -1.2], dtype=np.float32)
ROOM_MAX = np.array([3.0, 6.0, 1.5], dtype=np.float32)


def point_clouds(batch: int, npoints: int, seed: int = 0, dup_frac: float = 0.01,
                 near_origin: int = 3) -> np.ndarray:
    """(batch, npoints, 3) fp32 points uniform in a 6 x 5.5 x 2.7 m room, with
    `dup_frac` exact duplicates (the reference resamples WITH replacement when a
    scan has fewer than 20 000 points, utils/pc_util.py:24-32) and a few points
    with |p|^2 <= 1e-3, which FPS must skip (sampling_gpu.cu:104)."""
    rng = np.random.default_rng(seed)
    pc = rng.uniform(ROOM_MIN, ROOM_MAX, size=(batch, npoints, 3)).astype(np.float32)
    ndup = int(npoints * dup_frac)
    for b in range(batch):
        if ndup > 0 and npoints > 1:
            dst = rng.choice(npoints, size=ndup, replace=False)
            src = rng.integers(0, npoints, size=ndup)
            pc[b, dst] = pc[b, src]
        k = min(near_origin, max(npoints - 1, 0))
        if k > 0:
            where = rng.choice(np.arange(1, npoints), size=k, replace=False)
            pc[b, where] = rng.uniform(-0.015, 0.015, size=(k, 3)).astype(np.float32)
    return pc
