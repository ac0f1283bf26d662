"""Point manipulations on the hot path (reference utils/pc_util.py:38-73)."""
from __future__ import annotations

import torch


def shift_scale_points(pred_xyz, src_range, dst_range=None):
    """Affine map of `pred_xyz` (B, N, 3) from `src_range` = [min (B, 3), max (B, 3)]
    to `dst_range` (default the unit cube).  Evaluated exactly as the reference
    does, ``((x - src_min) * dst_diff) / src_diff + dst_min``, so fp32 results agree."""
    if dst_range is None:
        dst_range = [
            torch.zeros((src_range[0].shape[0], 3), device=src_range[0].device),
            torch.ones((src_range[0].shape[0], 3), device=src_range[0].device),
        ]
    if pred_xyz.ndim == 4:
        src_range = [x[:, None] for x in src_range]
        dst_range = [x[:, None] for x in dst_range]
    assert src_range[0].shape[0] == pred_xyz.shape[0]
    assert dst_range[0].shape[0] == pred_xyz.shape[0]
    assert src_range[0].shape[-1] == pred_xyz.shape[-1]
    assert src_range[0].shape == src_range[1].shape
    assert dst_range[0].shape == dst_range[1].shape
    assert src_range[0].shape == dst_range[1].shape
    src_diff = src_range[1][:, None, :] - src_range[0][:, None, :]
    dst_diff = dst_range[1][:, None, :] - dst_range[0][:, None, :]
    return (((pred_xyz - src_range[0][:, None, :]) * dst_diff) / src_diff) + dst_range[0][:, None, :]


def scale_points(pred_xyz, mult_factor):
    """pred_xyz (B, N, 3) * mult_factor (B, 3)."""
    if pred_xyz.ndim == 4:
        mult_factor = mult_factor[:, None]
    return pred_xyz * mult_factor[:, None, :]
