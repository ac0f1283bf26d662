"""CPU stand-ins for the CUDA ops: the CPU restatement of the reference's training step.

TEST INFRASTRUCTURE ONLY (same rule as oracle/pointnet2_oracle.c): imported by
tests/ (plumbing parity against the reference goldens in the GPU-less container),
by __graft_entry__.smoke() and by bench.py's `cpu_baseline` / `--impl reference`
legs, never by the package.  With these installed, the model / criterion Python
code runs the reference's own CPU arithmetic: torch CPU ops for everything the
reference reaches through PyTorch, the C oracle for the pointnet2 ops (which the
reference only has on CUDA), scipy for the assignment (criterion.py:73).
The product package has no such path: its ops raise on CPU tensors."""
from __future__ import annotations

import contextlib

import numpy as np
import torch

import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import oracle_pointnet2 as orc  # noqa: E402


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _fourier(xyz, gauss_B, d_out, input_range=None):
    x = xyz.clone()
    if input_range is not None:
        x = (x - input_range[0][:, None, :]) / (input_range[1][:, None, :] - input_range[0][:, None, :])
    x = x * (2 * np.pi)
    proj = torch.mm(x.reshape(-1, 3), gauss_B[:, :d_out]).view(x.shape[0], x.shape[1], d_out)
    return torch.cat([proj.sin(), proj.cos()], dim=2).permute(0, 2, 1)


def _hungarian(cost, nactual):
    from scipy.optimize import linear_sum_assignment

    b, nprop, _ = cost.shape
    inds = torch.zeros((b, nprop), dtype=torch.int64)
    mask = torch.zeros((b, nprop), dtype=torch.float32)
    c = cost.detach().cpu().numpy()
    for i in range(b):
        n = int(nactual[i])
        if n > 0:
            r, col = linear_sum_assignment(c[i, :, :n])
            inds[i, r] = _t(col).long()
            mask[i, r] = 1
    return inds, mask


def _giou_cpu(c1, c2, nums_k2, rotated, rot_k2_limit=None):
    from giou_ref import giou3d_ref

    rot = bool(rotated.item()) if isinstance(rotated, torch.Tensor) else bool(rotated)
    return giou3d_ref(c1, c2, nums_k2, rot, rot_k2_limit)


def _crop_cpu(images, scene, boxes, valid, res, dtype=torch.float32, mean=None, std=None):
    """The reference's own per-box op sequence (models/model_3detr.py:1034-1088) with torchvision on CPU."""
    import ref_crop
    from coda_neurips2023_b200.ops import CLIP_MEAN, CLIP_STD

    out = torch.zeros((boxes.shape[0], 3, res, res), dtype=torch.float32)
    m = torch.tensor(CLIP_MEAN).view(3, 1, 1)
    s = torch.tensor(CLIP_STD).view(3, 1, 1)
    for i in range(boxes.shape[0]):
        if bool(valid[i]):
            u8 = ref_crop.torchvision_sequence(images[int(scene[i])], [int(v) for v in boxes[i]], res)
            out[i] = (u8.float() / 255.0 - m) / s
    return out


def _project_corners_to_image(corners_xyz, inputs):
    """Undo the point-cloud augmentation, project the 8 corners into the image in fp64 and undo the image-side
    crop / flip (reference models/model_3detr.py:912-968 and datasets/sunrgbd_utils.py:611-635), as the same chain
    of fp64 tensor ops the reference runs.  (B, Q, 8, 3) -> uv (B, Q, 8, 2) f64, depth (B, Q, 8)."""
    flip = inputs["flip_array"].unsqueeze(-1)                   # (B, 1, 1)
    pts = corners_xyz.to(torch.double) * inputs["scale_array"].unsqueeze(1).to(torch.double)
    pts = torch.matmul(pts, inputs["rot_array"].unsqueeze(1).to(torch.double))
    if "zx_flip_array" in inputs:
        pts = torch.cat((pts[..., :1], pts[..., 1:2] * inputs["zx_flip_array"].view(-1, 1, 1, 1), pts[..., 2:]), -1)
    pts = torch.cat((pts[..., :1] * flip.to(torch.double).view(-1, 1, 1, 1), pts[..., 1:]), dim=-1)
    K = inputs["K"].unsqueeze(1).to(torch.double)
    Rtilt = inputs["Rtilt"].unsqueeze(1).to(torch.double)
    pc2 = torch.matmul(Rtilt.transpose(2, 3), pts.transpose(2, 3)).transpose(2, 3)
    pc2 = torch.stack((pc2[..., 0], -pc2[..., 2], pc2[..., 1]), dim=-1)  # depth -> camera axes
    uv = torch.matmul(pc2, K.transpose(2, 3))
    depth = uv[..., 2]
    u = uv[..., 0] / (depth + 1e-32)
    v = uv[..., 1] / (depth + 1e-32)
    wmax = (inputs["ori_width"].to(torch.double) - 1).view(-1, 1, 1)
    hmax = (inputs["ori_height"].to(torch.double) - 1).view(-1, 1, 1)
    zero = torch.zeros((), dtype=torch.double, device=u.device)
    u = torch.minimum(torch.maximum(u, zero), wmax) + inputs["y_offset"].to(torch.double).view(-1, 1, 1)
    v = torch.minimum(torch.maximum(v, zero), hmax) + inputs["x_offset"].to(torch.double).view(-1, 1, 1)
    img_flip = inputs["image_flip_array"].to(torch.double).view(-1, 1, 1)
    flip_len = inputs["flip_length"].to(torch.double).view(-1, 1, 1)
    u = u * img_flip + (1 - img_flip) * (flip_len - 1 - u)
    return torch.stack((u, v), dim=-1), depth


def _boxes_in_image(corners_xyz, size_unnorm, inputs):
    uv, depth = _project_corners_to_image(corners_xyz, inputs)
    xmin = uv[..., 0].amin(-1).to(torch.int32)
    ymin = uv[..., 1].amin(-1).to(torch.int32)
    xmax = uv[..., 0].amax(-1).to(torch.int32)
    ymax = uv[..., 1].amax(-1).to(torch.int32)
    valid = ((xmax - xmin) > 0) & ((ymax - ymin) > 0) & (depth.amin(-1) >= 0) & ~(size_unnorm.amax(-1) < 1e-16)
    return torch.stack((xmin, ymin, xmax, ymax), dim=-1), valid


@contextlib.contextmanager
def installed(giou_fn=_giou_cpu, crop_fn=_crop_cpu):
    """Patches coda_neurips2023_b200.{ops, pointnet2._ext} with CPU math."""
    from coda_neurips2023_b200 import attention_sm100, ops
    from coda_neurips2023_b200.pointnet2 import _ext

    saved = {}

    def patch(mod, name, fn):
        saved[(mod, name)] = getattr(mod, name)
        setattr(mod, name, fn)

    patch(ops, "layer_norm", lambda x, w, b, eps=1e-5: torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, eps))
    patch(ops, "softmax_rows", lambda x, log=False: (torch.log_softmax if log else torch.softmax)(x, dim=-1))
    patch(ops, "fourier_pos_embed", _fourier)
    patch(ops, "hungarian", _hungarian)
    patch(ops, "boxes_in_image", _boxes_in_image)
    import discovery_ref

    patch(ops, "novel_candidates", discovery_ref.novel_candidates_ref)

    def _linear(x, weight, bias=None, relu=False, nsplit=None):
        y = torch.nn.functional.linear(x, weight.reshape(weight.shape[0], -1), bias)
        return torch.relu(y) if relu else y

    patch(ops, "linear", _linear)
    F = torch.nn.functional
    patch(ops, "dropout_add", lambda x, resid, p, training: resid + F.dropout(x, p, training))
    patch(ops, "dropout", lambda x, p, training: F.dropout(x, p, training))

    def _bn_act_rows(h, bn, relu, drop_p, training):
        out = bn(h)                         # nn.BatchNorm1d on (rows, C): the reference's own module
        out = torch.relu(out) if relu else out
        return F.dropout(out, drop_p, training)

    patch(ops, "bn_act_rows", _bn_act_rows)
    # attention masks travel as (dense boolean (B, Lq, Lk), same) on the CPU path: the reference's own dense form
    def _attention(q, k, v, nhead, dropout_p=0.0, training=False, causal=False, mask=None):
        return attention_sm100._math(q, k, v, nhead, dropout_p, training, causal,
                                     attn_mask=None if mask is None else mask[0])

    def _attention_fused(a, b, layout, nhead, dropout_p=0.0, training=False, mask=None):
        e = a.shape[-1] // (3 if layout == "qkv" else 2)
        q, k = a[..., :e], a[..., e: 2 * e]
        v = a[..., 2 * e:] if layout == "qkv" else b
        return _attention(q, k, v, nhead, dropout_p, training, False, mask)

    def _mask_bits(mask, batch):
        m = mask.to(torch.bool)
        m = m.unsqueeze(0) if m.dim() == 2 else m
        m = m.expand(batch, -1, -1)
        return m, m

    def _radius_mask(xyz, radius):       # reference models/transformer.py:155-162
        m = torch.cdist(xyz, xyz, p=2) >= radius
        return m, m

    patch(ops, "attention", _attention)
    patch(ops, "attention_fused", _attention_fused)
    patch(ops, "attention_mask_bits", _mask_bits)
    patch(ops, "radius_mask_bits", _radius_mask)
    if giou_fn is not None:
        patch(ops, "giou3d", giou_fn)
    if crop_fn is not None:
        patch(ops, "crop_resize_normalize", crop_fn)
    patch(_ext, "furthest_point_sampling", lambda p, n: _t(orc.furthest_point_sampling(p.detach().numpy(), int(n))))
    patch(_ext, "gather_points", lambda p, i: _t(orc.gather_points(p.detach().numpy(), i.numpy())))
    patch(_ext, "gather_points_grad", lambda g, i, n: _t(orc.gather_points_grad(g.numpy(), i.numpy(), int(n))))
    patch(_ext, "ball_query", lambda nx, x, r, ns: _t(orc.ball_query(nx.numpy(), x.numpy(), float(r), int(ns))))
    patch(_ext, "group_points", lambda p, i: _t(orc.group_points(p.detach().numpy(), i.numpy())))
    patch(_ext, "group_points_grad", lambda g, i, n: _t(orc.group_points_grad(g.numpy(), i.numpy(), int(n))))

    def qg(xyz, new_xyz, radius, nsample, normalize):
        idx, g = orc.query_and_group_xyz(xyz.numpy(), new_xyz.numpy(), float(radius), int(nsample), normalize)
        # the CPU reference divides (torch CPU true division); keep the oracle's GPU form here
        return _t(idx), _t(g)

    patch(_ext, "query_and_group_xyz", qg)
    try:
        yield
    finally:
        for (mod, name), fn in saved.items():
            setattr(mod, name, fn)
