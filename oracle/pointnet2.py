"""TEST INFRASTRUCTURE ONLY -- numpy front-end over the C oracle (liboracle.so).

Mirrors the reference's Python operators (modules/pointnet2/pointnet2/pointnet2_utils.py)
including the pre-fills the reference's Python side performs before calling the
kernels (temp = 1e10, :26; ball-query idx zero-filled, :218; grads zero-filled,
:67,146,190) and the sqrt after three_nn (:98).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)
_c = ctypes.c_int


def build(force=False):
    """Compile liboracle.so with the recipe in oracle/Makefile."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pointnet2_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_opt_n_threads.argtypes = [_c]
        L.orc_opt_n_threads.restype = _c
        L.orc_num_threads.restype = _c
        L.orc_set_num_threads.argtypes = [_c]
        L.orc_set_num_threads.restype = None
        for name, args in {
            "orc_furthest_point_sampling": [_c, _c, _c, _f, _f, _i],
            "orc_furthest_point_sampling_fast": [_c, _c, _c, _f, _f, _i],
            "orc_gather_points": [_c, _c, _c, _c, _f, _i, _f],
            "orc_gather_points_grad": [_c, _c, _c, _c, _f, _i, _f],
            "orc_ball_query": [_c, _c, _c, ctypes.c_float, _c, _f, _f, _i],
            "orc_group_points": [_c, _c, _c, _c, _c, _f, _i, _f],
            "orc_group_points_grad": [_c, _c, _c, _c, _c, _f, _i, _f],
            "orc_three_nn": [_c, _c, _c, _f, _f, _f, _i],
            "orc_three_interpolate": [_c, _c, _c, _c, _f, _i, _f, _f],
            "orc_three_interpolate_grad": [_c, _c, _c, _c, _f, _i, _f, _f],
        }.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = None
        _LIB = L
    return _LIB


def _fp(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_f)


def _ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def opt_n_threads(n):
    return int(lib().orc_opt_n_threads(int(n)))


def furthest_point_sample(xyz, npoint, fast=False):
    """pointnet2_utils.py:10-36.  xyz (B,N,3) f32 -> idx (B,npoint) i32."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    idx = np.zeros((B, npoint), np.int32)
    temp = np.full((B, N), 1e10, np.float32)
    fn = lib().orc_furthest_point_sampling_fast if fast else lib().orc_furthest_point_sampling
    fn(B, N, npoint, _fp(xyz), _fp(temp), _ip(idx))
    return idx


def gather_operation(features, idx):
    """pointnet2_utils.py:39-73.  (B,C,N),(B,m) -> (B,C,m)."""
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    m = idx.shape[1]
    out = np.empty((B, C, m), np.float32)
    lib().orc_gather_points(B, C, N, m, _fp(features), _ip(idx), _fp(out))
    return out


def gather_operation_grad(grad_out, idx, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, m = grad_out.shape
    g = np.zeros((B, C, N), np.float32)
    lib().orc_gather_points_grad(B, C, N, m, _fp(grad_out), _ip(idx), _fp(g))
    return g


def ball_query(radius, nsample, xyz, new_xyz):
    """pointnet2_utils.py:200-229.  -> idx (B,npoint,nsample) i32."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = np.zeros((B, m, nsample), np.int32)
    lib().orc_ball_query(B, N, m, float(np.float32(radius)), nsample, _fp(new_xyz), _fp(xyz), _ip(idx))
    return idx


def grouping_operation(features, idx):
    """pointnet2_utils.py:156-197.  (B,C,N),(B,P,S) -> (B,C,P,S)."""
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    _, P, S = idx.shape
    out = np.empty((B, C, P, S), np.float32)
    lib().orc_group_points(B, C, N, P, S, _fp(features), _ip(idx), _fp(out))
    return out


def grouping_operation_grad(grad_out, idx, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, P, S = grad_out.shape
    g = np.zeros((B, C, N), np.float32)
    lib().orc_group_points_grad(B, C, N, P, S, _fp(grad_out), _ip(idx), _fp(g))
    return g


def three_nn(unknown, known):
    """pointnet2_utils.py:76-105.  Returns (sqrt(dist2), idx), both (B,n,3)."""
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = np.empty((B, n, 3), np.float32)
    idx = np.empty((B, n, 3), np.int32)
    lib().orc_three_nn(B, n, m, _fp(unknown), _fp(known), _fp(dist2), _ip(idx))
    return np.sqrt(dist2), idx


def three_interpolate(features, idx, weight):
    """pointnet2_utils.py:108-153.  (B,C,m),(B,n,3),(B,n,3) -> (B,C,n)."""
    features, idx, weight = _f32(features), _i32(idx), _f32(weight)
    B, C, m = features.shape
    n = idx.shape[1]
    out = np.empty((B, C, n), np.float32)
    lib().orc_three_interpolate(B, C, m, n, _fp(features), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, n = grad_out.shape
    g = np.zeros((B, C, m), np.float32)
    lib().orc_three_interpolate_grad(B, C, n, m, _fp(grad_out), _ip(idx), _fp(weight), _fp(g))
    return g


def query_and_group(radius, nsample, xyz, new_xyz, features=None, use_xyz=True):
    """QueryAndGroup.forward, pointnet2_utils.py:243-265 -> (B, 3+C, P, S)."""
    idx = ball_query(radius, nsample, xyz, new_xyz)
    xyz_t = np.ascontiguousarray(np.transpose(_f32(xyz), (0, 2, 1)))
    grouped_xyz = grouping_operation(xyz_t, idx)
    grouped_xyz -= np.transpose(_f32(new_xyz), (0, 2, 1))[..., None]
    if features is not None:
        gf = grouping_operation(features, idx)
        return np.concatenate([grouped_xyz, gf], axis=1) if use_xyz else gf
    assert use_xyz
    return grouped_xyz


# ---- independent pure-numpy statements (small cases only) used to cross-check the C code ----

def fps_numpy_keyed(xyz, npoint):
    """FPS via the closed-form tie-break key (bitrev(k mod bs), k div bs); float32 fma emulated in float64.

    d = fma(dz,dz,fma(dx,dx,dy*dy)) in float32: each fma is computed exactly in
    float64 (products of two float32 are exact in float64; the sum of an exact
    product and a float32 rounds once to float64, then once to float32 --
    double rounding is harmless here only with overwhelming probability, so
    this is a cross-check for small random inputs, not the oracle).
    """
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    bs = opt_n_threads(N)
    lg = bs.bit_length() - 1
    k = np.arange(N)
    t = k & (bs - 1)
    rev = np.zeros(N, np.int64)
    for bit in range(lg):
        rev |= ((t >> bit) & 1) << (lg - 1 - bit)
    key = rev * (1 << 32) + (k >> lg)
    out = np.zeros((B, npoint), np.int32)
    for b in range(B):
        temp = np.full(N, 1e10, np.float32)
        old = 0
        for j in range(1, npoint):
            d = xyz[b] - xyz[b, old]
            dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
            yy = (dy * dy).astype(np.float32)
            a = (dx.astype(np.float64) * dx.astype(np.float64) + yy.astype(np.float64)).astype(np.float32)
            dd = (dz.astype(np.float64) * dz.astype(np.float64) + a.astype(np.float64)).astype(np.float32)
            temp = np.minimum(dd, temp)
            cand = np.flatnonzero(temp == temp.max())
            old = int(cand[np.argmin(key[cand])])
            out[b, j] = old
    return out
