"""Autograd surface of the set-abstraction ops.

Same public names and call signatures as
third_party_pointnet2/pointnet2/pointnet2_utils.py (`furthest_point_sample`,
`gather_operation`, `three_nn`, `three_interpolate`, `grouping_operation`,
`ball_query`, `QueryAndGroup`, `GroupAll`, `RandomDropout`), implemented on the
C-ABI kernels in ``_ext``.  Index-producing ops are non-differentiable, as in the
reference (pointnet2_utils.py:69, :280).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext
from . import pytorch_utils as pt_utils  # noqa: F401  (re-exported name of the reference module)


class RandomDropout(nn.Module):
    """Feature dropout with a random rate in [0, p) (reference pointnet2_utils.py:37-45)."""

    def __init__(self, p: float = 0.5, inplace: bool = False):
        super().__init__()
        self.p, self.inplace = p, inplace

    def forward(self, X):
        theta = float(torch.empty(1).uniform_(0, self.p)[0])
        return nn.functional.dropout2d(X, theta, self.training, self.inplace) * (1.0 - theta)


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        """xyz (B, N, 3) -> int32 (B, npoint) indices of the furthest-point set."""
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B, C, N), idx (B, npoint) int32 -> (B, C, npoint)."""
        ctx.n = features.size(2)
        ctx.save_for_backward(idx)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.gather_points_grad(grad_out.contiguous(), idx, ctx.n), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor):
        """unknown (B, n, 3), known (B, m, 3) -> (dist (B, n, 3) L2, idx (B, n, 3) int32)."""
        dist2, idx = _ext.three_nn(unknown, known)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        """features (B, c, m), idx/weight (B, n, 3) -> (B, c, n)."""
        ctx.m = features.size(2)
        ctx.save_for_backward(idx, weight)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B, C, N), idx (B, npoint, nsample) int32 -> (B, C, npoint, nsample)."""
        ctx.n = features.size(2)
        ctx.save_for_backward(idx)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.group_points_grad(grad_out.contiguous(), idx, ctx.n), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        """xyz (B, N, 3), new_xyz (B, npoint, 3) -> int32 (B, npoint, nsample)."""
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)  # note: new_xyz first, as in the reference
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class _QueryAndGroupXYZ(Function):
    """Fused ball_query + grouping of xyz + centre subtraction (+ 1/radius)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, radius, nsample, normalize_xyz):
        idx, grouped = _ext.query_and_group_xyz(xyz, new_xyz, radius, nsample, normalize_xyz)
        ctx.mark_non_differentiable(idx)
        ctx.save_for_backward(idx)
        ctx.n = xyz.size(1)
        ctx.scale = (1.0 / radius) if normalize_xyz else 1.0
        return idx, grouped

    @staticmethod
    def backward(ctx, _gidx, grad_grouped):
        (idx,) = ctx.saved_tensors
        grad_xyz = grad_new = None
        g = grad_grouped if ctx.scale == 1.0 else grad_grouped * ctx.scale
        if ctx.needs_input_grad[0]:
            grad_xyz = _ext.group_points_grad(g.contiguous(), idx, ctx.n).transpose(1, 2)
        if ctx.needs_input_grad[1]:
            grad_new = -g.sum(dim=3).transpose(1, 2)
        return grad_xyz, grad_new, None, None, None


class QueryAndGroup(nn.Module):
    """Ball query of `radius` around each centre followed by grouping.

    Same constructor / forward contract as reference pointnet2_utils.py:291-373.
    When only coordinates are grouped and gradients w.r.t. them are not needed
    through a differentiable path other than the fused one, the fused kernel is
    used; it is bit-identical to the unfused op sequence.
    """

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if self.ret_unique_cnt:
            assert self.sample_uniformly

    def _resample_uniformly(self, idx):
        # reference pointnet2_utils.py:333-342: refill each ball with random repeats of its unique members
        unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
        for b in range(idx.shape[0]):
            for r in range(idx.shape[1]):
                uniq = torch.unique(idx[b, r, :])
                nu = uniq.shape[0]
                unique_cnt[b, r] = nu
                pick = torch.randint(0, nu, (self.nsample - nu,), dtype=torch.long)
                idx[b, r, :] = torch.cat((uniq, uniq[pick]))
        return unique_cnt

    def forward(self, xyz, new_xyz, features=None):
        """xyz (B, N, 3), new_xyz (B, npoint, 3), features (B, C, N) or None
        -> new_features (B, 3 + C, npoint, nsample) [, grouped_xyz][, unique_cnt]"""
        unique_cnt = None
        if self.sample_uniformly:
            idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
            unique_cnt = self._resample_uniformly(idx)
            xyz_trans = xyz.transpose(1, 2).contiguous()
            grouped_xyz = grouping_operation(xyz_trans, idx)
            grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
            if self.normalize_xyz:
                grouped_xyz /= self.radius
        else:
            idx, grouped_xyz = _QueryAndGroupXYZ.apply(
                xyz.contiguous(), new_xyz.contiguous(), float(self.radius), int(self.nsample),
                bool(self.normalize_xyz))

        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz

        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    """Groups every point into a single ball (reference pointnet2_utils.py:376-422)."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            new_features = grouped_xyz
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features
