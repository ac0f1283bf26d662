"""Generates tests/golden/eval_ap.npz from the REFERENCE's own evaluation code (utils/ap_calculator.py,
utils/eval_det.py, utils/nms.py, utils/box_util.py imported from /root/reference through _reference_harness) run on
CPU: synthetic predictions around synthetic ground truth -> parse_predictions (empty-box removal by the scipy hull
test, 3-D NMS, thresholds) -> APCalculator.step / compute_metrics.  Inputs AND outputs are stored, so the parity test
(tests/test_eval_gpu.py) needs nothing but the file.

    python tests/golden/make_eval_golden.py
"""
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(HERE))

import _reference_harness as H  # noqa: E402
from coda_neurips2023_b200 import synthetic  # noqa: E402

CONFIGS = {
    # the defaults the reference evaluates with (get_ap_config_dict): same-class 3-D NMS, per-class proposals
    "default": dict(),
    # class-agnostic 3-D NMS, one detection per box scored by its objectness, higher confidence threshold
    "agnostic": dict(cls_nms=False, per_class_proposal=False, conf_thresh=0.3),
    # bird's-eye-view NMS, class confidence as the score, no empty-box removal
    "bev": dict(use_3d_nms=False, per_class_proposal=False, use_cls_confidence_only=True, remove_empty_box=False),
}


def make_inputs(seed=0, batch=4, k=96, ncls=10, npoints=3000):
    box_util = H.load("utils.box_util")
    rng = np.random.default_rng(seed)
    d = synthetic.make_batch(batch, npoints, seed=seed, ncls_seen=ncls)
    present = d["gt_box_present"] > 0
    gt_cls = d["gt_box_seen_sem_cls_label"].astype(np.int64)
    centers = np.zeros((batch, k, 3), np.float32)
    sizes = np.zeros((batch, k, 3), np.float32)
    angles = np.zeros((batch, k), np.float32)
    probs = rng.random((batch, k, ncls)).astype(np.float32)
    for b in range(batch):
        ng = int(present[b].sum())
        for j in range(k):
            kind = rng.random()
            if kind < 0.55 and ng:          # around a ground-truth box
                g = int(rng.integers(0, ng))
                centers[b, j] = d["gt_box_centers"][b, g] + rng.normal(0, 0.12, 3)
                sizes[b, j] = d["gt_box_sizes"][b, g] * rng.uniform(0.8, 1.25, 3)
                angles[b, j] = d["gt_box_angles"][b, g] + rng.normal(0, 0.15)
                probs[b, j, gt_cls[b, g]] += rng.uniform(0.5, 3.0)
            elif kind < 0.8:                # anywhere in the room
                centers[b, j] = rng.uniform(synthetic.ROOM_MIN + 0.3, synthetic.ROOM_MAX - 0.3)
                sizes[b, j] = rng.uniform(0.2, 1.8, 3)
                angles[b, j] = rng.uniform(-np.pi, np.pi)
            elif kind < 0.9:                # too small to hold five points
                centers[b, j] = rng.uniform(synthetic.ROOM_MIN + 0.3, synthetic.ROOM_MAX - 0.3)
                sizes[b, j] = rng.uniform(0.01, 0.04, 3)
            # else: an all-zero box
        # near-duplicates of earlier boxes (NMS food), nudged so that no two boxes are identical
        for j in range(k - 8, k):
            src = int(rng.integers(0, k - 8))
            centers[b, j] = centers[b, src] + rng.normal(0, 0.01, 3)
            sizes[b, j], angles[b, j] = sizes[b, src], angles[b, src]
            probs[b, j] = probs[b, src] * rng.uniform(0.9, 1.1, ncls)
    probs = probs / probs.sum(-1, keepdims=True)
    obj = rng.random((batch, k)).astype(np.float32)
    cam = box_util.flip_axis_to_camera_np(centers)
    corners = box_util.get_3d_box_batch_np(sizes, angles, cam).astype(np.float32)
    corners[np.all(sizes == 0, axis=-1)] = 0
    return dict(point_clouds=d["point_clouds"], gt_box_corners=d["gt_box_corners"].astype(np.float32),
                gt_box_sem_cls_label=gt_cls, gt_box_present=d["gt_box_present"], box_corners=corners,
                sem_cls_prob=probs.astype(np.float32), objectness_prob=obj)


def main():
    apm = H.load("utils.ap_calculator")
    inp = make_inputs()
    t = {k: torch.from_numpy(v) for k, v in inp.items()}
    blob = {f"in.{k}": v for k, v in inp.items()}
    cfg_ds = SimpleNamespace(num_semcls=inp["sem_cls_prob"].shape[-1])
    for name, over in CONFIGS.items():
        cfg = apm.get_ap_config_dict(dataset_config=cfg_ds, **over)
        lists = apm.parse_predictions(t["box_corners"], t["sem_cls_prob"], t["objectness_prob"], t["point_clouds"], cfg)
        # which boxes became detections (any class): recover the box index from its corners
        b, k = inp["objectness_prob"].shape
        det = np.zeros((b, k), np.uint8)
        for i, cur in enumerate(lists):
            for cls_id, box, score in cur:
                j = np.where(np.all(np.all(inp["box_corners"][i] == box, axis=-1), axis=-1))[0]
                if len(j) > 1:      # the all-zero boxes share their corners: tell them apart by the score
                    if cfg["per_class_proposal"]:
                        want = inp["sem_cls_prob"][i, j, cls_id] * inp["objectness_prob"][i, j]
                    elif cfg["use_cls_confidence_only"]:
                        want = inp["sem_cls_prob"][i, j, cls_id]
                    else:
                        want = inp["objectness_prob"][i, j]
                    j = j[np.isclose(want, score, rtol=1e-6, atol=0)]
                assert len(j) == 1, "boxes must be distinguishable"
                det[i, j[0]] = 1
        blob[f"{name}.det_mask"] = det
        blob[f"{name}.ndet"] = np.array([len(cur) for cur in lists], np.int64)
        calc = apm.APCalculator(cfg_ds, ap_iou_thresh=[0.25, 0.5], class2type_map=None, exact_eval=True,
                                args=SimpleNamespace(dataset_name="sunrgbd"), ap_config_dict=cfg)
        # two steps of two scenes each: accumulation across steps is part of the contract
        for lo in (0, 2):
            calc.step_meter({"outputs": {"box_corners": t["box_corners"][lo:lo + 2],
                                         "sem_cls_prob": t["sem_cls_prob"][lo:lo + 2],
                                         "objectness_prob": t["objectness_prob"][lo:lo + 2]}},
                            {"point_clouds": t["point_clouds"][lo:lo + 2], "gt_box_corners": t["gt_box_corners"][lo:lo + 2],
                             "gt_box_sem_cls_label": t["gt_box_sem_cls_label"][lo:lo + 2],
                             "gt_box_present": t["gt_box_present"][lo:lo + 2]})
        if name == "default":
            # the pairwise IoUs the evaluation is built on (utils/box_util.py:156-183), for every detection x present
            # ground-truth box: pins the IoU kernel on its own
            box_util = H.load("utils.box_util")
            g = inp["gt_box_corners"].shape[1]
            iou = np.full((b, k, g), -1.0, np.float32)
            for i in range(b):
                for j in range(k):
                    if det[i, j]:
                        for q in range(g):
                            if inp["gt_box_present"][i, q] > 0:
                                iou[i, j, q] = box_util.box3d_iou(inp["box_corners"][i, j], inp["gt_box_corners"][i, q])[0]
            blob["default.iou"] = iou
        ret = calc.compute_metrics()
        for thr, rd in ret.items():
            keys = list(rd.keys())
            blob[f"{name}.{thr}.keys"] = np.array(keys)
            blob[f"{name}.{thr}.values"] = np.array([float(rd[kk]) for kk in keys], np.float64)
        print(name, {thr: (round(float(rd["mAP"]), 4), round(float(rd["AR"]), 4)) for thr, rd in ret.items()},
              "detections", int(det.sum()))
    np.savez_compressed(HERE / "eval_ap.npz", **blob)
    print("wrote", HERE / "eval_ap.npz")


if __name__ == "__main__":
    main()
