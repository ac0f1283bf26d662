"""Drop-in for the reference's pybind module ``pointnet2._ext``.

Same nine functions, names, argument order and tensor contracts as
third_party_pointnet2/pointnet2/_ext_src/src/bindings.cpp:9-22 (host functions in
sampling.cpp / ball_query.cpp / group_points.cpp / interpolate.cpp): inputs must
be contiguous fp32 / int32 CUDA tensors (the reference's CHECK_CONTIGUOUS /
CHECK_IS_FLOAT / CHECK_IS_INT, utils.h:8-28, which raise ``RuntimeError``),
outputs are freshly allocated on the input's device, kernels are enqueued on the
current stream without synchronising.  A CPU tensor raises "CPU not supported"
exactly like the reference (e.g. sampling.cpp:36).

This file is the whole "binding": shape checks + allocation + one C-ABI call
into libcoda_b200.so (include/coda_pointnet2.h).
"""
from __future__ import annotations

import ctypes

import torch

from .._lib import check, lib, ptr, stream_of

_c_int = ctypes.c_int
_c_float = ctypes.c_float


def _req(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)


def _check_float(t: torch.Tensor, name: str) -> None:
    _req(t.is_contiguous(), f"{name} must be a contiguous tensor")
    _req(t.dtype == torch.float32, f"{name} must be a float tensor")


def _check_int(t: torch.Tensor, name: str) -> None:
    _req(t.is_contiguous(), f"{name} must be a contiguous tensor")
    _req(t.dtype == torch.int32, f"{name} must be an int tensor")


def _check_cuda(t: torch.Tensor, name: str) -> None:
    _req(t.is_cuda, "CPU not supported" if name is None else f"{name} must be a CUDA tensor")


def furthest_point_sampling(points: torch.Tensor, nsamples: int) -> torch.Tensor:
    """sampling.cpp:67-88.  points (B, N, 3) -> int32 (B, nsamples)."""
    _check_float(points, "points")
    _req(points.is_cuda, "CPU not supported")
    _req(points.dim() == 3 and points.size(2) == 3, "points must be (B, N, 3)")
    b, n = points.size(0), points.size(1)
    nsamples = int(nsamples)
    out = torch.empty((b, max(nsamples, 0)), dtype=torch.int32, device=points.device)
    if b == 0 or nsamples <= 0:
        return out
    with torch.cuda.device(points.device):
        st = lib().coda_furthest_point_sampling(
            _c_int(b), _c_int(n), _c_int(nsamples), ptr(points), ptr(out), stream_of(points))
    check(st, "furthest_point_sampling")
    return out


def gather_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """sampling.cpp:17-40.  (B, C, N), (B, M) -> (B, C, M)."""
    _check_float(points, "points")
    _check_int(idx, "idx")
    _req(points.is_cuda, "CPU not supported")
    _check_cuda(idx, "idx")
    b, c, n = points.shape
    m = idx.size(1)
    out = torch.empty((b, c, m), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        st = lib().coda_gather_points(_c_int(b), _c_int(c), _c_int(n), _c_int(m), ptr(points),
                                      ptr(idx), ptr(out), stream_of(points))
    check(st, "gather_points")
    return out


def gather_points_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    """sampling.cpp:42-66.  (B, C, M), (B, M), n -> (B, C, n)."""
    _check_float(grad_out, "grad_out")
    _check_int(idx, "idx")
    _req(grad_out.is_cuda, "CPU not supported")
    _check_cuda(idx, "idx")
    b, c, m = grad_out.shape
    out = torch.zeros((b, c, int(n)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        st = lib().coda_gather_points_grad(_c_int(b), _c_int(c), _c_int(int(n)), _c_int(m),
                                           ptr(grad_out), ptr(idx), ptr(out), stream_of(grad_out))
    check(st, "gather_points_grad")
    return out


def ball_query(new_xyz: torch.Tensor, xyz: torch.Tensor, radius: float, nsample: int) -> torch.Tensor:
    """ball_query.cpp:11-35 -- note the argument order: new_xyz FIRST."""
    _check_float(new_xyz, "new_xyz")
    _check_float(xyz, "xyz")
    _req(new_xyz.is_cuda, "CPU not supported")
    _check_cuda(xyz, "xyz")
    b, m = new_xyz.size(0), new_xyz.size(1)
    n = xyz.size(1)
    nsample = int(nsample)
    idx = torch.empty((b, m, nsample), dtype=torch.int32, device=new_xyz.device)
    with torch.cuda.device(new_xyz.device):
        st = lib().coda_ball_query(_c_int(b), _c_int(n), _c_int(m), _c_float(radius), _c_int(nsample),
                                   ptr(new_xyz), ptr(xyz), ptr(idx), stream_of(new_xyz))
    check(st, "ball_query")
    return idx


def group_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """group_points.cpp:15-37.  (B, C, N), (B, M, S) -> (B, C, M, S)."""
    _check_float(points, "points")
    _check_int(idx, "idx")
    _req(points.is_cuda, "CPU not supported")
    _check_cuda(idx, "idx")
    b, c, n = points.shape
    m, s = idx.size(1), idx.size(2)
    out = torch.empty((b, c, m, s), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        st = lib().coda_group_points(_c_int(b), _c_int(c), _c_int(n), _c_int(m), _c_int(s),
                                     ptr(points), ptr(idx), ptr(out), stream_of(points))
    check(st, "group_points")
    return out


def group_points_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    """group_points.cpp:39-63.  (B, C, M, S), (B, M, S), n -> (B, C, n)."""
    _check_float(grad_out, "grad_out")
    _check_int(idx, "idx")
    _req(grad_out.is_cuda, "CPU not supported")
    _check_cuda(idx, "idx")
    b, c = grad_out.size(0), grad_out.size(1)
    m, s = idx.size(1), idx.size(2)
    out = torch.zeros((b, c, int(n)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        st = lib().coda_group_points_grad(_c_int(b), _c_int(c), _c_int(int(n)), _c_int(m), _c_int(s),
                                          ptr(grad_out), ptr(idx), ptr(out), stream_of(grad_out))
    check(st, "group_points_grad")
    return out


def three_nn(unknowns: torch.Tensor, knows: torch.Tensor):
    """interpolate.cpp:14-44.  (B, n, 3), (B, m, 3) -> [dist2 (B, n, 3), idx (B, n, 3)]."""
    _check_float(unknowns, "unknowns")
    _check_float(knows, "knows")
    _req(unknowns.is_cuda, "CPU not supported")
    _check_cuda(knows, "knows")
    b, n = unknowns.size(0), unknowns.size(1)
    m = knows.size(1)
    idx = torch.empty((b, n, 3), dtype=torch.int32, device=unknowns.device)
    dist2 = torch.empty((b, n, 3), dtype=torch.float32, device=unknowns.device)
    with torch.cuda.device(unknowns.device):
        st = lib().coda_three_nn(_c_int(b), _c_int(n), _c_int(m), ptr(unknowns), ptr(knows),
                                 ptr(dist2), ptr(idx), stream_of(unknowns))
    check(st, "three_nn")
    return [dist2, idx]


def three_interpolate(points: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """interpolate.cpp:46-73.  (B, C, m), (B, n, 3), (B, n, 3) -> (B, C, n)."""
    _check_float(points, "points")
    _check_int(idx, "idx")
    _check_float(weight, "weight")
    _req(points.is_cuda, "CPU not supported")
    _check_cuda(idx, "idx")
    _check_cuda(weight, "weight")
    b, c, m = points.shape
    n = idx.size(1)
    out = torch.empty((b, c, n), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        st = lib().coda_three_interpolate(_c_int(b), _c_int(c), _c_int(m), _c_int(n), ptr(points),
                                          ptr(idx), ptr(weight), ptr(out), stream_of(points))
    check(st, "three_interpolate")
    return out


def three_interpolate_grad(grad_out: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor, m: int) -> torch.Tensor:
    """interpolate.cpp:75-101.  (B, C, n), (B, n, 3), (B, n, 3), m -> (B, C, m)."""
    _check_float(grad_out, "grad_out")
    _check_int(idx, "idx")
    _check_float(weight, "weight")
    _req(grad_out.is_cuda, "CPU not supported")
    _check_cuda(idx, "idx")
    _check_cuda(weight, "weight")
    b, c, n = grad_out.shape
    out = torch.zeros((b, c, int(m)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        st = lib().coda_three_interpolate_grad(_c_int(b), _c_int(c), _c_int(n), _c_int(int(m)),
                                               ptr(grad_out), ptr(idx), ptr(weight), ptr(out),
                                               stream_of(grad_out))
    check(st, "three_interpolate_grad")
    return out


# --- not in the reference module: fused QueryAndGroup for the xyz-only SA layer ---
def query_and_group_xyz(xyz: torch.Tensor, new_xyz: torch.Tensor, radius: float, nsample: int,
                        normalize_xyz: bool):
    """ball_query + group(xyz^T) + "-= new_xyz" (+ "/= radius") in one kernel.

    Returns (idx int32 (B, M, S), grouped_xyz fp32 (B, 3, M, S)); bit-identical to
    the op sequence of pointnet2_utils.py:331-349 run with torch CUDA ops.
    """
    _check_float(xyz, "xyz")
    _check_float(new_xyz, "new_xyz")
    _req(xyz.is_cuda, "CPU not supported")
    _check_cuda(new_xyz, "new_xyz")
    b, n = xyz.size(0), xyz.size(1)
    m = new_xyz.size(1)
    nsample = int(nsample)
    idx = torch.empty((b, m, nsample), dtype=torch.int32, device=xyz.device)
    grouped = torch.empty((b, 3, m, nsample), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        st = lib().coda_query_and_group_xyz(_c_int(b), _c_int(n), _c_int(m), _c_float(radius),
                                            _c_int(nsample), _c_int(1 if normalize_xyz else 0),
                                            ptr(xyz), ptr(new_xyz), ptr(idx), ptr(grouped),
                                            stream_of(xyz))
    check(st, "query_and_group_xyz")
    return idx, grouped
