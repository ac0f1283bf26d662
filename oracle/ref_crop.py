"""numpy restatement of the crop/pad/resize kernel's arithmetic (CHECKER for
csrc/image_kernels.cu; mirrors ATen's upsample_bicubic2d_aa weight rule) and the
reference op sequence built from torchvision (models/model_3detr.py:1034-1088)."""
import numpy as np
import torch


def _cubic(x):
    a = np.float32(-0.5)
    x = np.abs(x).astype(np.float32)
    return np.where(x < 1, ((a + 2) * x - (a + 3)) * x * x + 1,
                    np.where(x < 2, (((x - 5) * x + 8) * x - 4) * a, 0)).astype(np.float32)


def _weights(scale, out_size, in_size):
    support = np.float32(2 * scale if scale >= 1 else 2)
    inv = np.float32(1 / scale if scale >= 1 else 1)
    out = []
    for i in range(out_size):
        center = np.float32(scale) * np.float32(i + 0.5)
        lo = max(int(center - support + np.float32(0.5)), 0)
        n = min(int(center + support + np.float32(0.5)), in_size) - lo
        w = _cubic((np.arange(n, dtype=np.float32) + np.float32(lo) - center + np.float32(0.5)) * inv)
        tot = w.sum(dtype=np.float32)
        if tot != 0:
            w = w / tot
        out.append((lo, w.astype(np.float32)))
    return out


def crop_resize_uint8(img_hwc: np.ndarray, box, res: int) -> np.ndarray:
    """uint8 (res, res, 3) -> what the kernel computes before normalisation."""
    xmin, ymin, xmax, ymax = box
    crop = img_hwc[ymin:ymax, xmin:xmax].astype(np.float32)
    hc, wc = crop.shape[:2]
    e = max(hc, wc)
    canvas = np.full((e, e, 3), 255, np.float32)
    yb, xb = (e - hc) // 2, (e - wc) // 2
    canvas[yb:yb + hc, xb:xb + wc] = crop
    scale = np.float32(e) / np.float32(res)
    wts = _weights(scale, res, e)
    tmp = np.zeros((e, res, 3), np.float32)
    for ox, (lo, w) in enumerate(wts):
        tmp[:, ox] = (canvas[:, lo:lo + len(w)] * w[None, :, None]).sum(1)
    out = np.zeros((res, res, 3), np.float32)
    for oy, (lo, w) in enumerate(wts):
        out[oy] = (tmp[lo:lo + len(w)] * w[:, None, None]).sum(0)
    return np.rint(np.clip(out, 0, 255)).astype(np.uint8)


def torchvision_sequence(img_hwc: torch.Tensor, box, res: int) -> torch.Tensor:
    """The reference's op sequence for one box: crop, 255-canvas, Resize(BICUBIC) -> (3, res, res) uint8."""
    from torchvision.transforms import InterpolationMode, Resize

    xmin, ymin, xmax, ymax = box
    img_crop = img_hwc[ymin:ymax, xmin:xmax]
    w, h = ymax - ymin, xmax - xmin
    e = max(w, h)
    bg = torch.ones(e, e, 3, dtype=torch.uint8, device=img_hwc.device) * 255
    yb, xb = (e - w) // 2, (e - h) // 2
    bg[yb:yb + w, xb:xb + h, :] = img_crop
    return Resize(res, interpolation=InterpolationMode.BICUBIC)(bg.permute(2, 0, 1))
