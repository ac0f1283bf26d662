"""CPU side of the evaluation path: the AP integration rule (utils/eval_det.py:23-55) and the golden fixture that the
GPU parity test (tests/test_eval_gpu.py) compares the device kernels with."""
from pathlib import Path

import numpy as np

from coda_neurips2023_b200.utils import ap_calculator as apc

GOLDEN = Path(__file__).resolve().parent / "golden" / "eval_ap.npz"


def test_voc_ap_area_under_the_precision_envelope():
    rec = np.array([0.25, 0.25, 0.5, 0.75, 0.75, 1.0])
    prec = np.array([1.0, 0.5, 2 / 3, 0.75, 0.6, 4 / 6])
    # envelope: 1.0 up to 0.25, 0.75 up to 0.75, 2/3 up to 1.0
    assert abs(apc.voc_ap(rec, prec) - (0.25 * 1.0 + 0.5 * 0.75 + 0.25 * (4 / 6))) < 1e-12
    assert apc.voc_ap(np.zeros(3), np.zeros(3)) == 0.0
    # the 11-point variant (:30-38)
    assert abs(apc.voc_ap(rec, prec, use_07_metric=True) - (3 * 1.0 + 5 * 0.75 + 3 * (4 / 6)) / 11) < 1e-12


def test_eval_golden_fixture_is_complete():
    g = np.load(GOLDEN)
    for name in ("default", "agnostic", "bev"):
        assert g[f"{name}.det_mask"].shape == g["in.objectness_prob"].shape
        for thr in ("0.25", "0.5"):
            keys = list(g[f"{name}.{thr}.keys"])
            assert "mAP" in keys and "AR" in keys and "Prec" in keys and len(keys) == len(g[f"{name}.{thr}.values"])
    cfg = apc.get_ap_config_dict()
    assert cfg["use_3d_nms"] and cfg["cls_nms"] and cfg["per_class_proposal"] and cfg["conf_thresh"] == 0.05
