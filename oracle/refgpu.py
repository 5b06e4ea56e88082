"""TEST INFRASTRUCTURE ONLY -- the reference's own pointnet2 CUDA kernels on the GPU.

Loads oracle/_ref/libpointnet2_ref.so: the four reference .cu files compiled
UNMODIFIED for sm_100a by oracle/build_ref.sh, behind oracle/ref_shim.cu.  The
operators below repeat what the reference's Python layer does around each
kernel call (pointnet2_utils.py) -- allocation, pre-fills, sqrt -- so that the
outputs are "the reference pointnet2 CUDA ops" the north star asks parity with.
Needs a GPU; used by the -m gpu parity tests and by tests/golden/make_golden.py.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libpointnet2_ref.so")
_LIB = None


def available():
    # G4D_NO_REFGPU=1: leave the reference kernels out (under compute-sanitizer's instrumentation the reference FPS kernel, which
    # relies on implicit warp-synchronous execution in its last reduction steps, returns wrong indices)
    return os.path.exists(SO) and os.environ.get("G4D_NO_REFGPU", "0") != "1"


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(SO)
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(rc):
    if rc != 0:
        raise RuntimeError(f"reference kernel launch failed: cudaError {rc}")


def furthest_point_sample(xyz, npoint):
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    _chk(lib().ref_furthest_point_sampling(B, N, npoint, _p(xyz), _p(temp), _p(idx), _s()))
    return idx


def gather_operation(features, idx):
    B, C, N = features.shape
    m = idx.shape[1]
    out = torch.empty(B, C, m, dtype=torch.float32, device=features.device)
    _chk(lib().ref_gather_points(B, C, N, m, _p(features), _p(idx), _p(out), _s()))
    return out


def gather_operation_grad(grad_out, idx, N):
    B, C, m = grad_out.shape
    g = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
    _chk(lib().ref_gather_points_grad(B, C, N, m, _p(grad_out), _p(idx), _p(g), _s()))
    return g


def ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(B, m, nsample, dtype=torch.int32, device=xyz.device)
    _chk(lib().ref_ball_query(B, N, m, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s()))
    return idx


def grouping_operation(features, idx):
    B, C, N = features.shape
    _, P, S = idx.shape
    out = torch.empty(B, C, P, S, dtype=torch.float32, device=features.device)
    _chk(lib().ref_group_points(B, C, N, P, S, _p(features), _p(idx), _p(out), _s()))
    return out


def grouping_operation_grad(grad_out, idx, N):
    B, C, P, S = grad_out.shape
    g = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
    _chk(lib().ref_group_points_grad(B, C, N, P, S, _p(grad_out), _p(idx), _p(g), _s()))
    return g


def three_nn_raw(unknown, known):
    """Returns (dist2, idx) exactly as the kernel writes them (squared distances)."""
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
    _chk(lib().ref_three_nn(B, n, m, _p(unknown), _p(known), _p(d2), _p(idx), _s()))
    return d2, idx


def three_nn(unknown, known):
    d2, idx = three_nn_raw(unknown, known)
    return torch.sqrt(d2), idx


def three_interpolate(features, idx, weight):
    B, C, m = features.shape
    n = idx.shape[1]
    out = torch.empty(B, C, n, dtype=torch.float32, device=features.device)
    _chk(lib().ref_three_interpolate(B, C, m, n, _p(features), _p(idx), _p(weight), _p(out), _s()))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    B, C, n = grad_out.shape
    g = torch.zeros(B, C, m, dtype=torch.float32, device=grad_out.device)
    _chk(lib().ref_three_interpolate_grad(B, C, n, m, _p(grad_out), _p(idx), _p(weight), _p(g), _s()))
    return g
