"""Pins the numpy restatement of the crop / pad / antialiased-bicubic resize (oracle/ref_crop.py, the arithmetic
csrc/image_kernels.cu implements) to the reference's own op sequence for one box (models/model_3detr.py:1034-1088:
crop, 255-filled square canvas, torchvision Resize(BICUBIC) on the uint8 tensor) run on the CPU."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
import ref_crop  # noqa: E402

pytest.importorskip("torchvision")


def _image(seed, h, w):
    rng = np.random.default_rng(seed)
    base = torch.from_numpy(rng.integers(0, 256, size=(h // 8, w // 8, 3), dtype=np.uint8)).permute(2, 0, 1)[None].float()
    img = torch.nn.functional.interpolate(base, size=(h, w), mode="bilinear").round().clamp(0, 255)
    return img.to(torch.uint8)[0].permute(1, 2, 0).contiguous()


@pytest.mark.parametrize("res", [64, 36])
def test_numpy_crop_restatement_equals_the_torchvision_sequence(res):
    img = _image(res, 160, 212)
    boxes = [[10, 20, 200, 150],      # down-sampling, wide
             [0, 0, 212, 160],        # the whole image
             [50, 60, 80, 70],        # up-sampling
             [5, 5, 10, 150],         # a sliver: almost all white canvas
             [100, 3, 101, 4]]        # one pixel
    for box in boxes:
        got = ref_crop.crop_resize_uint8(img.numpy(), box, res)
        exp = ref_crop.torchvision_sequence(img, box, res).permute(1, 2, 0).numpy()
        diff = np.abs(got.astype(np.int32) - exp.astype(np.int32))
        assert diff.max() <= 1, (box, int(diff.max()))
        assert (diff > 0).mean() < 1e-3, (box, float((diff > 0).mean()))
