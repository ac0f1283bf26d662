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
    patch(ops, "attention", lambda q, k, v, nhead, dropout_p=0.0, training=False, causal=False:
          attention_sm100._math(q, k, v, nhead, dropout_p, training, causal))
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
