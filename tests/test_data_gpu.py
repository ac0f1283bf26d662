"""Device data layer (coda_neurips2023_b200/datasets/device_pipeline.py, include/coda_data.h) against its CPU
restatement oracle/data_ref.py (itself pinned to the reference's RandomCuboid in tests/test_data_cpu.py): integer
results -- chosen crop, kept boxes, sampled rows -- are exact, coordinates bit-exact (explicit float32 rounding order)."""
import numpy as np
import pytest
import torch

import data_ref
from coda_neurips2023_b200 import synthetic
from coda_neurips2023_b200.datasets import DeviceSceneAugmentor, draw_augmentation
from coda_neurips2023_b200.datasets.device_pipeline import rotz

pytestmark = pytest.mark.gpu


def _raw(batch, nmax, seed, stride=3):
    rng = np.random.default_rng(seed)
    npts = rng.integers(int(0.6 * nmax), nmax + 1, size=batch)
    pts = np.zeros((batch, nmax, stride), np.float32)
    for b in range(batch):
        pts[b, : npts[b], :3] = synthetic.point_clouds(1, int(npts[b]), seed=seed + b)[0]
        if stride > 3:
            pts[b, : npts[b], 3:] = rng.random((int(npts[b]), stride - 3))
    gmax = 12
    nbox = rng.integers(0, gmax + 1, size=batch)
    nbox[0] = 0                                                  # a scene without ground truth (common in SUN RGB-D)
    boxes = np.zeros((batch, gmax, 8), np.float32)
    for b in range(batch):
        g = int(nbox[b])
        boxes[b, :g, 0:3] = rng.uniform(synthetic.ROOM_MIN + 0.5, synthetic.ROOM_MAX - 0.5, size=(g, 3))
        boxes[b, :g, 3:6] = rng.uniform(0.2, 1.0, size=(g, 3))
        boxes[b, :g, 6] = rng.uniform(-3, 3, size=g)
        boxes[b, :g, 7] = rng.integers(0, 10, size=g)
    return pts, npts.astype(np.int32), boxes, nbox.astype(np.int32)


@pytest.mark.parametrize("batch,nmax,nsample,min_points,stride,seed",
                         [(4, 9000, 4000, 2500, 3, 0), (3, 50000, 20000, 30000, 6, 1), (2, 3000, 4000, 500, 3, 2),
                          (2, 6000, 3000, 5990, 3, 3)])
def test_points_pipeline_equals_cpu_restatement(built_lib, batch, nmax, nsample, min_points, stride, seed):
    pts, npts, boxes, nbox = _raw(batch, nmax, seed, stride)
    params = draw_augmentation(np.random.default_rng(seed + 50), batch, min_crop=0.5, max_crop=1.0)
    aug = DeviceSceneAugmentor(num_points=nsample, random_cuboid_min_points=min_points, aspect=0.8)
    got = aug.points(torch.from_numpy(pts).cuda(), torch.from_numpy(npts).cuda(), torch.from_numpy(boxes).cuda(),
                     torch.from_numpy(nbox).cuda(), params)
    rot = rotz(np.asarray(params["rot_angle"], np.float64)).astype(np.float32)
    nchosen = 0
    for b in range(batch):
        n, g = int(npts[b]), int(nbox[b])
        p = data_ref.scene_transform(pts[b, :n], params["flip"][b], rot[b], params["scale"][b])
        bx = got["boxes"][b, :g].cpu().numpy()                  # the few box rows: checked against float64 below
        chosen, crop, keep = data_ref.random_cuboid(p, bx, params["crop_range"][b], params["center_u"][b], min_points, 0.8)
        assert int(got["chosen"][b]) == chosen
        assert np.array_equal(got["box_keep"][b, :g].cpu().numpy(), keep)
        nchosen += chosen >= 0
        out, choice, count, dims = data_ref.sample_points(p, crop, int(params["seed"][b]), nsample)
        assert int(got["count"][b]) == count
        assert np.array_equal(got["choice"][b].cpu().numpy().astype(np.int64), choice)
        assert np.array_equal(got["point_clouds"][b].cpu().numpy(), out)          # bit-exact coordinates
        assert np.array_equal(got["dims"][b].cpu().numpy(), dims)
        # boxes vs the reference formulas in float64 (datasets/...:663-703)
        ref = boxes[b, :g].astype(np.float64).copy()
        if params["flip"][b] < 0:
            ref[:, 0] *= -1
            ref[:, 6] = np.pi - ref[:, 6]
        ref[:, 0:3] = np.dot(ref[:, 0:3], rot[b].astype(np.float64).T)
        ref[:, 6] -= params["rot_angle"][b]
        ref[:, 0:6] *= params["scale"][b]
        assert np.abs(bx - ref).max() < 1e-5 if g else True
    if min_points >= 5990:
        assert nchosen == 0            # no crop can keep that many points: every scene falls back to "no crop"
    elif nsample < nmax // 2:
        assert nchosen > 0


def test_image_augment_equals_cpu_restatement_and_reference_formula(built_lib):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(3, 121, 97, 3), dtype=np.uint8)
    params = draw_augmentation(rng, 3)
    params["image_flip"] = np.array([1, 0, 1], np.uint8)
    got = DeviceSceneAugmentor().images(torch.from_numpy(img).cuda(), params).cpu().numpy()
    for b in range(3):
        exp = data_ref.image_augment(img[b], bool(params["image_flip"][b]), params["image_gain"][b], params["image_shift"][b],
                                     int(params["image_seed"][b]))
        assert np.array_equal(got[b], exp)
        # the reference's float64 arithmetic without the jitter: within the jitter's 0.025 * 255 + 1 levels
        ref = (img[b][:, ::-1] if params["image_flip"][b] else img[b]) / 255.0
        ref = np.clip(ref * params["image_gain"][b].astype(np.float64) + params["image_shift"][b], 0, 1) * 255.0
        assert np.abs(got[b].astype(np.float64) - ref).max() <= 0.025 * 255 + 1.0


def test_labels_follow_the_dataset_encoding(built_lib):
    """angle class / residual re-encoding, sizes, normalised centres of datasets/...:707-795 from augmented boxes"""
    pts, npts, boxes, nbox = _raw(3, 5000, 4)
    params = draw_augmentation(np.random.default_rng(1), 3, min_crop=0.5)
    aug = DeviceSceneAugmentor(num_points=2000, random_cuboid_min_points=800, aspect=0.8)
    got = aug.points(torch.from_numpy(pts).cuda(), torch.from_numpy(npts).cuda(), torch.from_numpy(boxes).cuda(),
                     torch.from_numpy(nbox).cuda(), params)
    lab = aug.labels(got["boxes"], got["box_keep"], got["dims"], None)
    bx, keep = got["boxes"].cpu().numpy().astype(np.float64), got["box_keep"].cpu().numpy()
    for b in range(3):
        kept = bx[b][keep[b]]
        n = len(kept)
        assert int(lab["gt_box_present"][b].sum()) == n
        per = 2 * np.pi / 12
        for i in range(n):
            ang = kept[i, 6] % (2 * np.pi)
            shifted = (ang + per / 2) % (2 * np.pi)
            cid = int(shifted / per)
            assert int(lab["gt_angle_class_label"][b, i]) == cid
            assert abs(float(lab["gt_angle_residual_label"][b, i]) - (shifted - (cid * per + per / 2))) < 1e-6
            assert np.allclose(lab["gt_box_sizes"][b, i].cpu().numpy(), kept[i, 3:6] * 2, atol=1e-6)
            assert np.allclose(lab["gt_box_centers"][b, i].cpu().numpy(), kept[i, 0:3], atol=1e-6)
        assert float(lab["gt_box_sizes"][b, n:].abs().sum()) == 0.0
