"""numpy front-end of oracle/pointnet2_oracle.c (CHECKER ONLY -- never imported by the package)."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = ORACLE_DIR / "liboracle_pointnet2.so"
        src = ORACLE_DIR / "pointnet2_oracle.c"
        if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
            subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
        _LIB = ctypes.CDLL(str(so))
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def opt_n_threads(n: int) -> int:
    return _lib().oracle_opt_n_threads(ctypes.c_int(n))


def furthest_point_sampling(xyz: np.ndarray, m: int) -> np.ndarray:
    xyz, px = _f(xyz)
    b, n, _ = xyz.shape
    out = np.zeros((b, max(m, 0)), dtype=np.int32)
    _lib().oracle_furthest_point_sampling(b, n, m, px, out.ctypes.data_as(ctypes.c_void_p))
    return out


def gather_points(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.zeros((b, c, m), dtype=np.float32)
    _lib().oracle_gather_points(b, c, n, m, pp, pi, out.ctypes.data_as(ctypes.c_void_p))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    b, c, m = grad_out.shape
    out = np.zeros((b, c, n), dtype=np.float32)
    _lib().oracle_gather_points_grad(b, c, n, m, pg, pi, out.ctypes.data_as(ctypes.c_void_p))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, pn = _f(new_xyz)
    xyz, px = _f(xyz)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = np.zeros((b, m, nsample), dtype=np.int32)
    _lib().oracle_ball_query(b, n, m, ctypes.c_float(radius), nsample, pn, px,
                             out.ctypes.data_as(ctypes.c_void_p))
    return out


def group_points(points, idx):
    points, pp = _f(points)
    idx, pi = _i(idx)
    b, c, n = points.shape
    _, m, s = idx.shape
    out = np.zeros((b, c, m, s), dtype=np.float32)
    _lib().oracle_group_points(b, c, n, m, s, pp, pi, out.ctypes.data_as(ctypes.c_void_p))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    b, c, m, s = grad_out.shape
    out = np.zeros((b, c, n), dtype=np.float32)
    _lib().oracle_group_points_grad(b, c, n, m, s, pg, pi, out.ctypes.data_as(ctypes.c_void_p))
    return out


def three_nn(unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d = np.zeros((b, n, 3), dtype=np.float32)
    i = np.zeros((b, n, 3), dtype=np.int32)
    _lib().oracle_three_nn(b, n, m, pu, pk, d.ctypes.data_as(ctypes.c_void_p),
                           i.ctypes.data_as(ctypes.c_void_p))
    return d, i


def three_interpolate(points, idx, weight):
    points, pp = _f(points)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.zeros((b, c, n), dtype=np.float32)
    _lib().oracle_three_interpolate(b, c, m, n, pp, pi, pw, out.ctypes.data_as(ctypes.c_void_p))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    b, c, n = grad_out.shape
    out = np.zeros((b, c, m), dtype=np.float32)
    _lib().oracle_three_interpolate_grad(b, c, n, m, pg, pi, pw, out.ctypes.data_as(ctypes.c_void_p))
    return out


def query_and_group_xyz(xyz, new_xyz, radius, nsample, normalize):
    xyz, px = _f(xyz)
    new_xyz, pn = _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((b, m, nsample), dtype=np.int32)
    g = np.zeros((b, 3, m, nsample), dtype=np.float32)
    _lib().oracle_query_and_group_xyz(b, n, m, ctypes.c_float(radius), nsample, int(bool(normalize)), px, pn,
                                      idx.ctypes.data_as(ctypes.c_void_p), g.ctypes.data_as(ctypes.c_void_p))
    return idx, g
