"""Drop-in for the reference's compiled extension module ``pointnet2_cuda``.

The reference's Python layer does ``import pointnet2_cuda as pointnet2``
(modules/pointnet2/pointnet2/pointnet2_utils.py:7) and calls nine functions, all positionally, with
caller-allocated CUDA tensors (modules/pointnet2/pointnet2/src/pointnet2_api.cpp:10-23).  This module
exposes the same nine names with the same argument order and meaning on top of the C ABI in
include/garment4d_b200.h.  Put the repository root on ``sys.path`` (it holds a top-level
``pointnet2_cuda.py`` that re-exports this module) and the reference's own ``pointnet2_utils.py`` runs
unmodified on the B200 kernels.

Differences from the reference, all deliberate:
  * a failed launch raises ``G4DError`` instead of printing to stderr and calling ``exit(-1)``
    (e.g. sampling_gpu.cu:39-43);
  * tensors are checked (CUDA, contiguous, fp32/int32): the reference only asserts contiguity in
    Python and checks nothing in C++ except in ball_query (ball_query.cpp:10-17).
"""
import torch

from . import _lib

__all__ = [
    "ball_query_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "gather_points_wrapper",
    "gather_points_grad_wrapper", "furthest_point_sampling_wrapper", "three_nn_wrapper",
    "three_interpolate_wrapper", "three_interpolate_grad_wrapper",
]


def _chk(t, dtype, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")       # CHECK_CUDA, ball_query.cpp:10
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")          # CHECK_CONTIGUOUS, ball_query.cpp:11
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    return _lib.ptr(t)


_F, _I = torch.float32, torch.int32


def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    """sampling.cpp:36-46"""
    rc = _lib.lib().g4d_furthest_point_sampling(b, n, m, _chk(points_tensor, _F, "points"), _chk(temp_tensor, _F, "temp"),
                                                _chk(idx_tensor, _I, "idx"), _lib.stream_ptr())
    _lib.check(rc, "furthest_point_sampling_wrapper")
    return 1


def gather_points_wrapper(b, c, n, npoints, points_tensor, idx_tensor, out_tensor):
    """sampling.cpp:11-21"""
    rc = _lib.lib().g4d_gather_points(b, c, n, npoints, _chk(points_tensor, _F, "points"), _chk(idx_tensor, _I, "idx"),
                                      _chk(out_tensor, _F, "out"), _lib.stream_ptr())
    _lib.check(rc, "gather_points_wrapper")
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out_tensor, idx_tensor, grad_points_tensor):
    """sampling.cpp:24-34"""
    rc = _lib.lib().g4d_gather_points_grad(b, c, n, npoints, _chk(grad_out_tensor, _F, "grad_out"), _chk(idx_tensor, _I, "idx"),
                                           _chk(grad_points_tensor, _F, "grad_points"), _lib.stream_ptr())
    _lib.check(rc, "gather_points_grad_wrapper")
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz_tensor, xyz_tensor, idx_tensor):
    """ball_query.cpp:14-25 -- note: new_xyz before xyz."""
    rc = _lib.lib().g4d_ball_query(b, n, m, float(radius), nsample, _chk(new_xyz_tensor, _F, "new_xyz"),
                                   _chk(xyz_tensor, _F, "xyz"), _chk(idx_tensor, _I, "idx"), _lib.stream_ptr())
    _lib.check(rc, "ball_query_wrapper")
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points_tensor, idx_tensor, out_tensor):
    """group_points.cpp:25-36"""
    rc = _lib.lib().g4d_group_points(b, c, n, npoints, nsample, _chk(points_tensor, _F, "points"), _chk(idx_tensor, _I, "idx"),
                                     _chk(out_tensor, _F, "out"), _lib.stream_ptr())
    _lib.check(rc, "group_points_wrapper")
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out_tensor, idx_tensor, grad_points_tensor):
    """group_points.cpp:11-22"""
    rc = _lib.lib().g4d_group_points_grad(b, c, n, npoints, nsample, _chk(grad_out_tensor, _F, "grad_out"),
                                          _chk(idx_tensor, _I, "idx"), _chk(grad_points_tensor, _F, "grad_points"),
                                          _lib.stream_ptr())
    _lib.check(rc, "group_points_grad_wrapper")
    return 1


def three_nn_wrapper(b, n, m, unknown_tensor, known_tensor, dist2_tensor, idx_tensor):
    """interpolate.cpp:14-23"""
    rc = _lib.lib().g4d_three_nn(b, n, m, _chk(unknown_tensor, _F, "unknown"), _chk(known_tensor, _F, "known"),
                                 _chk(dist2_tensor, _F, "dist2"), _chk(idx_tensor, _I, "idx"), _lib.stream_ptr())
    _lib.check(rc, "three_nn_wrapper")


def three_interpolate_wrapper(b, c, m, n, points_tensor, idx_tensor, weight_tensor, out_tensor):
    """interpolate.cpp:26-39"""
    rc = _lib.lib().g4d_three_interpolate(b, c, m, n, _chk(points_tensor, _F, "points"), _chk(idx_tensor, _I, "idx"),
                                          _chk(weight_tensor, _F, "weight"), _chk(out_tensor, _F, "out"), _lib.stream_ptr())
    _lib.check(rc, "three_interpolate_wrapper")


def three_interpolate_grad_wrapper(b, c, n, m, grad_out_tensor, idx_tensor, weight_tensor, grad_points_tensor):
    """interpolate.cpp:42-54"""
    rc = _lib.lib().g4d_three_interpolate_grad(b, c, n, m, _chk(grad_out_tensor, _F, "grad_out"), _chk(idx_tensor, _I, "idx"),
                                               _chk(weight_tensor, _F, "weight"), _chk(grad_points_tensor, _F, "grad_points"),
                                               _lib.stream_ptr())
    _lib.check(rc, "three_interpolate_grad_wrapper")
