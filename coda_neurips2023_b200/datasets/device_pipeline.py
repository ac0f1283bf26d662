"""The per-scene numpy pipeline of the reference's dataset `__getitem__`
(datasets/sunrgbd_anonymous_aligned_image.py:618-795; the ScanNet twin has the same steps) for a whole BATCH of raw
scenes that already live in HBM: point-cloud / box augmentation (flip about YZ, rotation about the up axis, scale),
RandomCuboid (utils/random_cuboid.py), random sampling to `num_points` (utils/pc_util.py:24-32), the image
augmentation (:624-655), and the label tensors the model and the criterion read.

Why on the device: the reference runs this in DataLoader workers, one scene at a time, on the host; at 275 scenes/s
per GPU that is ~14 M points/s of numpy work per GPU plus an 11 MB host-to-device copy per step.  Raw scenes are a
few GB for SUN RGB-D (10 335 scenes x ~50 k points x 12 B): they fit in HBM many times over, so the epoch loop
becomes index selection + five kernel launches (include/coda_data.h), no workers, no copies.

Randomness stays with the caller, as small arrays (`draw_augmentation`): the same role np.random plays in the
reference, which makes the device path a deterministic function that oracle/data_ref.py restates on the CPU.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from .._lib import check, lib, ptr, stream_of

_i = lambda v: ctypes.c_int(int(v))  # noqa: E731
_f = lambda v: ctypes.c_float(float(v))  # noqa: E731


def draw_augmentation(rng: np.random.Generator, batch: int, ncand: int = 100, augment: bool = True,
                      min_crop: float = 0.75, max_crop: float = 1.0, image_augment: bool = True) -> dict:
    """One step's random numbers (host, numpy) -- what the reference draws inside __getitem__ / RandomCuboid:
    flip (:663), rotation angle in +-30 degrees (:672), scale 0.85-1.15 (:700), per attempt a crop range in
    [min_crop, max_crop]^3 and a centre point (random_cuboid.py:43-50), a sampling seed (pc_util.py:28), and for the
    image a flip, per-channel gain 0.8-1.2 and shift +-0.05, a jitter seed (:630-645)."""
    p = {}
    if augment:
        p["flip"] = np.where(rng.random(batch) > 0.5, -1.0, 1.0).astype(np.float32)
        p["rot_angle"] = (rng.random(batch) * np.pi / 3 - np.pi / 6)
        p["scale"] = (rng.random(batch) * 0.3 + 0.85).astype(np.float32)
    else:
        p["flip"] = np.ones(batch, np.float32)
        p["rot_angle"] = np.zeros(batch)
        p["scale"] = np.ones(batch, np.float32)
    p["crop_range"] = min_crop + rng.random((batch, ncand, 3)) * (max_crop - min_crop)
    p["center_u"] = rng.random((batch, ncand)).astype(np.float32)
    p["seed"] = rng.integers(0, 2 ** 32, size=batch, dtype=np.uint32)
    if image_augment:
        p["image_flip"] = (rng.random(batch) > 0.5).astype(np.uint8)
        p["image_gain"] = (1 + 0.4 * rng.random((batch, 3)) - 0.2).astype(np.float32)
        p["image_shift"] = (0.1 * rng.random((batch, 3)) - 0.05).astype(np.float32)
        p["image_seed"] = rng.integers(0, 2 ** 32, size=batch, dtype=np.uint32)
    return p


def rotz(t: np.ndarray) -> np.ndarray:
    """(B,) angles -> (B, 3, 3) rotation about the up axis (utils/pc_util.py:125-129)"""
    c, s = np.cos(t), np.sin(t)
    z, o = np.zeros_like(c), np.ones_like(c)
    return np.stack((np.stack((c, -s, z), -1), np.stack((s, c, z), -1), np.stack((z, z, o), -1)), -2)


class DeviceSceneAugmentor:
    """raw scenes on the device -> the collated training batch.

    raw_points (B, Nmax, stride) fp32 with npts (B,) valid rows (xyz [+ colour]); raw_boxes (B, Gmax, 8) fp32 rows
    [cx, cy, cz, l/2, w/2, h/2, heading, class] in the upright depth frame with nbox (B,) valid rows
    (datasets/...:441-442 `_bbox.npy`)."""

    def __init__(self, num_points: int = 20000, max_num_obj: int = 64, num_angle_bin: int = 12, augment: bool = True,
                 use_random_cuboid: bool = True, random_cuboid_min_points: int = 30000, aspect: float = 0.75,
                 ncand: int = 100):
        self.num_points, self.max_num_obj, self.num_angle_bin = num_points, max_num_obj, num_angle_bin
        self.augment, self.use_random_cuboid = augment, use_random_cuboid
        self.min_points, self.aspect, self.ncand = random_cuboid_min_points, aspect, ncand

    # ------------------------------------------------------------------ kernels
    @torch.no_grad()
    def points(self, raw_points: torch.Tensor, npts: torch.Tensor, raw_boxes: torch.Tensor, nbox: torch.Tensor,
               params: dict):
        """-> dict(point_clouds (B, num_points, stride), choice, dims (B, 6), boxes (B, Gmax, 8) transformed,
        box_keep (B, Gmax) bool, chosen (B,) int32)"""
        if not raw_points.is_cuda:
            raise RuntimeError("DeviceSceneAugmentor: CPU not supported (oracle/data_ref.py is the CPU restatement)")
        dev = raw_points.device
        b, nmax, stride = raw_points.shape
        gmax = raw_boxes.shape[1]
        pts = raw_points.detach().float().clone().contiguous()       # transformed in place
        npts_i = npts.to(device=dev, dtype=torch.int32).contiguous()
        nbox_i = nbox.to(device=dev, dtype=torch.int32).contiguous()
        up = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)  # noqa: E731
        flip, scale = up(params["flip"], torch.float32), up(params["scale"], torch.float32)
        rot64 = rotz(np.asarray(params["rot_angle"], np.float64))
        rot = up(rot64.astype(np.float32), torch.float32)
        L = lib()
        st = stream_of(pts)
        with torch.cuda.device(dev):
            check(L.coda_scene_transform(_i(b), _i(nmax), _i(stride), ptr(npts_i), ptr(flip), ptr(rot), ptr(scale),
                                         ptr(pts), st), "scene_transform")
            # boxes: a handful of rows per scene (datasets/...:664-666, 675-677, 701-703), same float32 arithmetic
            boxes = raw_boxes.detach().float().clone()
            f, s = flip.view(b, 1), scale.view(b, 1)
            ang = up(np.asarray(params["rot_angle"], np.float64).astype(np.float32), torch.float32).view(b, 1)
            boxes[..., 0] = boxes[..., 0] * f
            boxes[..., 6] = torch.where(f < 0, math.pi - boxes[..., 6], boxes[..., 6]) - ang
            ctr = boxes[..., 0:3]
            boxes[..., 0:3] = torch.stack([(ctr[..., 0] * rot[:, j, 0:1] + ctr[..., 1] * rot[:, j, 1:2])
                                           + ctr[..., 2] * rot[:, j, 2:3] for j in range(3)], dim=-1) * s.unsqueeze(-1)
            boxes[..., 3:6] = boxes[..., 3:6] * s.unsqueeze(-1)
            boxes = boxes.contiguous()
            extent = torch.empty((b, 6), dtype=torch.float32, device=dev)
            check(L.coda_points_extent(_i(b), _i(nmax), _i(stride), ptr(npts_i), ptr(pts), ptr(extent), st),
                  "points_extent")
            range_xyz = (extent[:, 3:] - extent[:, :3]).contiguous()
            crop = torch.empty((b, 6), dtype=torch.float64, device=dev)
            chosen = torch.full((b,), -1, dtype=torch.int32, device=dev)
            keep = torch.ones((b, max(gmax, 1)), dtype=torch.uint8, device=dev)
            if self.augment and self.use_random_cuboid:
                cr = up(params["crop_range"], torch.float64)
                cu = up(params["center_u"], torch.float32)
                ncand = cr.shape[1]
                scratch = torch.empty((b, ncand, 8), dtype=torch.float32, device=dev)
                check(L.coda_random_cuboid(_i(b), _i(nmax), _i(stride), _i(ncand), _i(gmax), _i(boxes.shape[2]),
                                           _i(self.min_points), _f(self.aspect), ptr(npts_i), ptr(pts), ptr(range_xyz),
                                           ptr(cr), ptr(cu), ptr(boxes), ptr(nbox_i), ptr(scratch), ptr(chosen),
                                           ptr(crop), ptr(keep), st), "random_cuboid")
            else:
                crop[:, :3] = float("-inf")
                crop[:, 3:] = float("inf")
                keep = (torch.arange(max(gmax, 1), device=dev).view(1, -1) < nbox_i.view(b, 1)).to(torch.uint8)
            seed = up(np.asarray(params["seed"]).astype(np.int64), torch.int64).to(torch.int32).contiguous()   # bit pattern
            lst = torch.empty((b, nmax), dtype=torch.int32, device=dev)
            count = torch.empty((b,), dtype=torch.int32, device=dev)
            out = torch.empty((b, self.num_points, stride), dtype=torch.float32, device=dev)
            choice = torch.empty((b, self.num_points), dtype=torch.int32, device=dev)
            dims = torch.empty((b, 6), dtype=torch.float32, device=dev)
            check(L.coda_sample_points(_i(b), _i(nmax), _i(stride), _i(self.num_points), ptr(npts_i), ptr(pts), ptr(crop),
                                       ptr(seed), ptr(lst), ptr(count), ptr(out), ptr(choice), ptr(dims), st),
                  "sample_points")
        return dict(point_clouds=out, choice=choice, dims=dims, boxes=boxes, box_keep=keep[:, :gmax].bool(),
                    chosen=chosen, count=count, crop=crop, rot=rot64)

    @torch.no_grad()
    def images(self, images: torch.Tensor, params: dict) -> torch.Tensor:
        """(B, H, W, 3) uint8 -> augmented uint8 (datasets/...:624-655)"""
        if not images.is_cuda or images.dtype != torch.uint8:
            raise RuntimeError("images must be uint8 CUDA tensors (B, H, W, 3)")
        img = images.contiguous()
        b, h, w, _ = img.shape
        dev = img.device
        up = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)  # noqa: E731
        flip = up(params["image_flip"], torch.uint8)
        gain, shift = up(params["image_gain"], torch.float32), up(params["image_shift"], torch.float32)
        seed = up(np.asarray(params["image_seed"]).astype(np.int64), torch.int64).to(torch.int32).contiguous()
        out = torch.empty_like(img)
        with torch.cuda.device(dev):
            check(lib().coda_image_augment(_i(b), _i(h), _i(w), ptr(img), ptr(flip), ptr(gain), ptr(shift), ptr(seed),
                                           ptr(out), stream_of(img)), "image_augment")
        return out

    # ------------------------------------------------------------------ labels (tiny tensors: plain tensor ops)
    @torch.no_grad()
    def labels(self, boxes: torch.Tensor, box_keep: torch.Tensor, dims: torch.Tensor, dataset_config) -> dict:
        """The ground-truth tensors of datasets/...:707-795 from the augmented boxes: kept boxes are packed to the
        front (RandomCuboid drops the others), padded to max_num_obj."""
        b, gmax, _ = boxes.shape
        g = self.max_num_obj
        dev = boxes.device
        order = torch.argsort((~box_keep).to(torch.int8), dim=1, stable=True)           # kept boxes first, in order
        bx = torch.gather(boxes.double(), 1, order.unsqueeze(-1).expand(-1, -1, boxes.shape[2]))
        present = torch.gather(box_keep, 1, order)
        pad = g - gmax
        if pad > 0:
            bx = torch.cat((bx, bx.new_zeros(b, pad, bx.shape[2])), 1)
            present = torch.cat((present, present.new_zeros(b, pad)), 1)
        bx, present = bx[:, :g], present[:, :g]
        mask = present.double()
        bx = bx * mask.unsqueeze(-1)
        two_pi = 2 * math.pi
        per = two_pi / self.num_angle_bin
        ang = bx[..., 6] % two_pi
        shifted = (ang + per / 2) % two_pi
        cls = torch.floor(shifted / per).long()
        res = shifted - (cls.double() * per + per / 2)
        raw_sizes = bx[..., 3:6] * 2
        # re-encoded angle, as class2angle_batch does (:771-773): centre of the bin + residual, wrapped to (-pi, pi]
        raw_angles = cls.double() * per + res
        raw_angles = torch.where(raw_angles > math.pi, raw_angles - two_pi, raw_angles)
        # axis-aligned extent of the heading-rotated box (:722-741)
        c, s = torch.cos(-bx[..., 6]), torch.sin(-bx[..., 6])
        l, w, h = bx[..., 3], bx[..., 4], bx[..., 5]
        ex = (c * l).abs() + (s * w).abs()
        ey = (s * l).abs() + (c * w).abs()
        centers = bx[..., 0:3]
        dmin, dmax = dims[:, :3].double().unsqueeze(1), dims[:, 3:].double().unsqueeze(1)
        span = dmax - dmin
        out = {
            "gt_box_present": mask.float(),
            "gt_box_centers": centers.float(),
            "gt_box_centers_normalized": (((centers - dmin) / span) * mask.unsqueeze(-1)).float(),
            "gt_box_sizes": raw_sizes.float(),
            "gt_box_sizes_normalized": (raw_sizes / span).float(),
            "gt_box_angles": (raw_angles * mask).float(),
            "gt_angle_class_label": cls * present.long(),
            "gt_angle_residual_label": (res * mask).float(),
            "gt_box_extent": torch.stack((2 * ex, 2 * ey, 2 * h), -1).float(),
            "gt_box_sem_cls_label": bx[..., 7].long() * present.long(),
            "point_cloud_dims_min": dims[:, :3].contiguous(),
            "point_cloud_dims_max": dims[:, 3:].contiguous(),
        }
        if dataset_config is not None:
            out["gt_box_corners"] = dataset_config.box_parametrization_to_corners(
                centers.float(), raw_sizes.float(), raw_angles.float()) * mask.float().view(b, g, 1, 1)
        return out
