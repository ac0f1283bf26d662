"""CPU (no GPU) plumbing parity: OUR model + criterion Python code, with the CUDA ops
replaced by test-only CPU stand-ins (tests/cpu_shims.py), against golden outputs of
the REFERENCE model + criterion run on CPU (tests/golden/make_model_golden.py).
This pins everything that is not a kernel: module wiring, parameter names, box
decoding, projection to the image, crop selection, matching cost, every active
loss and its gradient.  The kernels themselves are tested on the GPU."""
import numpy as np
import pytest
import torch

import cpu_shims
import model_parity_common as mpc
import ref_crop


def _giou_cpu(c1, c2, nums_k2, rotated, rot_k2_limit=None):
    # numpy/torch restatement of the kernel's formulas is exercised on the GPU; on CPU use a
    # simple exact implementation for the axis-aligned part and shapely-free polygon clipping
    from giou_ref import giou3d_ref

    rot = bool(rotated.item()) if isinstance(rotated, torch.Tensor) else bool(rotated)
    return giou3d_ref(c1, c2, nums_k2, rot, rot_k2_limit)


def _crop_cpu(images, scene, boxes, valid, res, dtype=torch.float32, mean=None, std=None):
    from coda_neurips2023_b200.ops import CLIP_MEAN, CLIP_STD

    out = torch.zeros((boxes.shape[0], 3, res, res), dtype=torch.float32)
    m = torch.tensor(CLIP_MEAN).view(3, 1, 1)
    s = torch.tensor(CLIP_STD).view(3, 1, 1)
    for i in range(boxes.shape[0]):
        if not bool(valid[i]):
            continue
        u8 = ref_crop.crop_resize_uint8(images[int(scene[i])].numpy(), [int(v) for v in boxes[i]], res)
        out[i] = (torch.from_numpy(u8).permute(2, 0, 1).float() / 255.0 - m) / s
    return out


@pytest.mark.parametrize("name", list(mpc.CASES))
def test_model_and_criterion_match_reference_on_cpu(name):
    torch.manual_seed(0)
    with cpu_shims.installed(giou_fn=_giou_cpu, crop_fn=_crop_cpu):
        model, out, loss, loss_dict, golden = mpc.run(name, "cpu")
        errs = mpc.compare(model, out, loss, loss_dict, golden, rtol=2e-4, atol=1e-5)
    worst = max(errs, key=errs.get)
    print(f"{name}: worst {worst} = {errs[worst]:.2e}")
