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


# ---------------------------------------------------------------------------
#  Synthetic training batches, argument namespace and dataset config
# ---------------------------------------------------------------------------
import argparse  # noqa: E402


class SyntheticDatasetConfig:
    """The attributes / methods of the reference's dataset config objects that the model
    and the criterion touch (datasets/sunrgbd_anonymous_aligned_image.py:86-300)."""

    def __init__(self, args=None, num_semcls: int = 1, num_angle_bin: int = 12, max_num_obj: int = 64):
        self.num_semcls = num_semcls
        self.num_angle_bin = num_angle_bin
        self.max_num_obj = max_num_obj
        self.image_size = [getattr(args, "image_size_width", 730), getattr(args, "image_size_height", 531)]

    def box_parametrization_to_corners(self, box_center_unnorm, box_size, box_angle):
        from .utils.box_util import flip_axis_to_camera_tensor, get_3d_box_batch_tensor

        return get_3d_box_batch_tensor(box_size, box_angle, flip_axis_to_camera_tensor(box_center_unnorm))

    def box_parametrization_to_corners_xyz(self, box_center_unnorm, box_size, box_angle):
        from .utils.box_util import get_3d_box_batch_tensor_xyz

        return get_3d_box_batch_tensor_xyz(box_size, box_angle, box_center_unnorm)


def make_args(**overrides) -> argparse.Namespace:
    """Defaults of the reference's main.py argument parser (main.py:37-304) for every field
    the model / criterion read, overlaid with scripts/coda_sunrgbd_stage1.sh, overlaid with
    `overrides`.  nqueries defaults to the BASELINE metric's 256."""
    a = dict(
        # model (main.py:69-74, :127-143; stage1 script)
        model_name="3detr_predictedbox_distillation", dataset_name="sunrgbd_anonymous_aligned_image",
        enc_type="vanilla", enc_nlayers=3, enc_dim=256, enc_ffn_dim=128, enc_dropout=0.1, enc_nhead=4,
        enc_activation="relu", dec_nlayers=8, dec_dim=512, dec_ffn_dim=256, dec_dropout=0.1, dec_nhead=4,
        mlp_dropout=0.3, nqueries=256, preenc_npoints=2048, use_color=False,
        if_use_v1=True, if_clip_more_prompts=True, if_clip_superset=False, if_clip_weak_labels=False,
        train_range_max=10, test_range_max=46, distillation_box_num=32, keep_objectness=0.5,
        image_size_width=730, image_size_height=531, clip_arch="ViT-B/32",
        clip_checkpoint="./CLIP/pretrain_models/ViT-B-16.pt",
        # matcher / losses (stage1 script)
        matcher_giou_cost=3.0, matcher_cls_cost=1.0, matcher_center_cost=5.0, matcher_objectness_cost=5.0,
        loss_giou_weight=0.0, loss_sem_cls_weight=0.0, loss_sem_cls_softmax_weight=0.0,
        loss_sem_cls_softmax_skip_none_gt_sample_weight=1.0, loss_no_object_weight=0.05,
        loss_angle_cls_weight=0.1, loss_angle_reg_weight=0.5, loss_center_weight=5.0, loss_size_weight=1.0,
        loss_no_object_contrast_weight=0.05, loss_predicted_region_embed_l1_weight=1.0,
        confidence_type="non-confidence",
        # optimiser (main.py:41-52; stage1 script)
        base_lr=1.97e-4, warm_lr=1e-6, warm_lr_epochs=18, final_lr=1e-6, lr_scheduler="cosine",
        weight_decay=0.1, filter_biases_wd=False, clip_gradient=0.1, max_epoch=1080,
        batchsize_per_gpu=8, ngpus=1,
    )
    a.update(overrides)
    return argparse.Namespace(**a)


def _corners_np(size, angle, center_cam):
    """(..., 8, 3) camera-frame corners in numpy (same closed form as utils/box_util.py)."""
    sx = np.array((1, 1, -1, -1, 1, 1, -1, -1), np.float32)
    sy = np.array((1, 1, 1, 1, -1, -1, -1, -1), np.float32)
    sz = np.array((1, -1, -1, 1, 1, -1, -1, 1), np.float32)
    x = size[..., 0:1] / 2 * sx
    y = size[..., 2:3] / 2 * sy
    z = size[..., 1:2] / 2 * sz
    c, s = np.cos(angle)[..., None], np.sin(angle)[..., None]
    out = np.stack((x * c + z * s, y, -x * s + z * c), -1)
    return (out + center_cam[..., None, :]).astype(np.float32)


def make_batch(batch: int, npoints: int = 20000, seed: int = 0, max_gt: int = 64, image_hw=(531, 730),
               num_angle_bin: int = 12, ncls_seen: int = 10, min_gt: int = 3, max_real_gt: int = 20):
    """One SUN RGB-D-shaped training batch as numpy arrays / what the dataloader collates
    (datasets/sunrgbd_anonymous_aligned_image.py:813-899).  fp64 where
    the reference's numpy arrays are fp64 (K, Rtilt, rot_array, scale_array, flip arrays)."""
    rng = np.random.default_rng(seed + 7919)
    pc = point_clouds(batch, npoints, seed=seed)
    h, w = image_hw
    d = {
        "point_clouds": pc,
        "point_cloud_dims_min": pc.min(axis=1),
        "point_cloud_dims_max": pc.max(axis=1),
    }
    present = np.zeros((batch, max_gt), np.float32)
    centers = np.zeros((batch, max_gt, 3), np.float32)
    sizes = np.zeros((batch, max_gt, 3), np.float32)
    angles = np.zeros((batch, max_gt), np.float32)
    for b in range(batch):
        n = int(rng.integers(min_gt, max_real_gt + 1))
        present[b, :n] = 1
        centers[b, :n] = rng.uniform(ROOM_MIN + 0.3, ROOM_MAX - 0.3, size=(n, 3))
        sizes[b, :n] = rng.uniform(0.3, 2.0, size=(n, 3))
        angles[b, :n] = rng.uniform(-np.pi, np.pi, size=n)
    per_cls = 2 * np.pi / num_angle_bin
    shifted = (angles % (2 * np.pi) + per_cls / 2) % (2 * np.pi)
    cls = (shifted / per_cls).astype(np.int64)
    res = (shifted - (cls * per_cls + per_cls / 2)).astype(np.float32)
    span = d["point_cloud_dims_max"] - d["point_cloud_dims_min"]
    cam = np.stack((centers[..., 0], -centers[..., 2], centers[..., 1]), -1)
    d.update({
        "gt_box_present": present,
        "gt_box_centers": centers,
        "gt_box_centers_normalized": ((centers - d["point_cloud_dims_min"][:, None]) / span[:, None]).astype(np.float32),
        "gt_box_sizes": sizes,
        "gt_box_sizes_normalized": (sizes / np.clip(span, 0.1, None)[:, None]).astype(np.float32),
        "gt_box_angles": angles * present,
        "gt_angle_class_label": cls * present.astype(np.int64),
        "gt_angle_residual_label": res * present,
        "gt_box_corners": _corners_np(sizes, angles, cam) * present[..., None, None],
        "gt_box_sem_cls_label": np.zeros((batch, max_gt), np.int64),
        "gt_box_seen_sem_cls_label": rng.integers(0, ncls_seen, size=(batch, max_gt)).astype(np.int64),
        "gt_box_seen_sem_cls_confi": present.copy(),
    })
    # image side: SUN RGB-D-like intrinsics, small tilt, augmentation bookkeeping
    # SUN RGB-D intrinsics at 730 x 531; other image shapes (ScanNet: 1296 x 968) scale the focal length with
    # the width and keep the principal point at the same relative position
    f = 529.5 * w / 730.0
    K = np.tile(np.array([[f, 0, 365.0 * w / 730.0], [0, f, 265.0 * h / 531.0], [0, 0, 1.0]]), (batch, 1, 1))
    tilt = rng.uniform(-0.05, 0.05, size=batch)
    Rtilt = np.stack([np.array([[1, 0, 0], [0, np.cos(t), -np.sin(t)], [0, np.sin(t), np.cos(t)]]) for t in tilt])
    rot = rng.uniform(-np.pi / 18, np.pi / 18, size=batch)
    rot_array = np.stack([np.array([[np.cos(t), -np.sin(t), 0], [np.sin(t), np.cos(t), 0], [0, 0, 1.0]]) for t in rot])
    d.update({
        "input_image": rng.integers(0, 256, size=(batch, h, w, 3), dtype=np.uint8),
        "K": K.astype(np.float64), "Rtilt": Rtilt.astype(np.float64), "rot_array": rot_array.astype(np.float64),
        "flip_array": rng.choice([-1.0, 1.0], size=(batch, 1)).astype(np.float64),
        "scale_array": rng.uniform(0.9, 1.1, size=(batch, 1, 1)).repeat(3, axis=2).astype(np.float64),
        "image_flip_array": rng.choice([0.0, 1.0], size=(batch, 1)).astype(np.float64),
        "flip_length": np.full((batch,), float(w), np.float64),
        "ori_width": np.full((batch,), w, np.int64), "ori_height": np.full((batch,), h, np.int64),
        "x_offset": np.zeros((batch,), np.int64), "y_offset": np.zeros((batch,), np.int64),
        # stage 2 (novel-box discovery) reads these as well (model_3detr.py:1228, :1515)
        "rot_angle": rot.astype(np.float64),
        "gt_ori_box_num": present.sum(axis=1).astype(np.int64),
    })
    return d


def to_device(batch_np: dict, device, pinned: bool = False) -> dict:
    """What engine.py:125-129 does with a collated batch: every array -> tensor on `device`."""
    import torch

    out = {}
    for k, v in batch_np.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        if pinned:
            t = t.pin_memory()
        out[k] = t.to(device, non_blocking=pinned)
    return out
