"""The reference's point-cloud operators (modules/pointnet2/pointnet2/pointnet2_utils.py) on the B200 kernels.

Same public names, argument order, shapes, dtypes and autograd contract as the reference:

    furthest_point_sample(xyz, npoint) -> idx (B,npoint) int32            non-differentiable
    gather_operation(features (B,C,N), idx (B,m)) -> (B,C,m)              grad w.r.t. features
    three_nn(unknown (B,n,3), known (B,m,3)) -> (dist (B,n,3), idx)       non-differentiable
    three_interpolate(features (B,C,m), idx, weight) -> (B,C,n)           grad w.r.t. features
    grouping_operation(features (B,C,N), idx (B,P,S)) -> (B,C,P,S)        grad w.r.t. features
    ball_query(radius, nsample, xyz, new_xyz) -> idx (B,P,nsample) int32  non-differentiable
    QueryAndGroup(radius, nsample, use_xyz)(xyz, new_xyz, features)       -> (B, 3+C, P, nsample)
    GroupAll(use_xyz)(xyz, new_xyz, features)                             -> (B, 3+C, 1, N)

The nine low-level calls go through ``garment4d_b200.pointnet2_cuda`` (the mirror of the reference's compiled
module).  Two fused entry points are added behind the same operators:
``furthest_point_sample_and_gather`` (FPS + centroid gather, one launch) and the single-kernel
``QueryAndGroup.forward``.
"""
import os
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib
from .. import pointnet2_cuda as pointnet2


def _i32(*shape, device):
    return torch.empty(*shape, dtype=torch.int32, device=device)


def _f32(*shape, device):
    return torch.empty(*shape, dtype=torch.float32, device=device)


def attach_point_major(features: torch.Tensor, pm: torch.Tensor) -> None:
    """Remembers the fp16 point-major copy ``pm (B,N,C)`` of ``features (B,C,N)`` on the tensor, together with the tensor's
    version counter: the copy is only trusted while `features` has not been modified in place (see point_major_of)."""
    try:
        features._g4d_pm = (pm, features._version)
    except Exception:
        pass


def point_major_of(features: torch.Tensor):
    """The fp16 point-major copy attached by attach_point_major, or None when there is none or it is stale: `features`
    was written in place since (version counter), or shape / dtype / device no longer match."""
    hit = getattr(features, "_g4d_pm", None)
    if not isinstance(hit, tuple) or len(hit) != 2:
        return None
    pm, ver = hit
    B, C, N = features.shape
    if (ver != features._version or pm.dtype != torch.float16 or pm.device != features.device or tuple(pm.shape) != (B, N, C)
            or not pm.is_contiguous()):
        try:
            del features._g4d_pm
        except Exception:
            pass
        return None
    return pm


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    """The checks the compiled-module mirror applies (garment4d_b200/pointnet2_cuda.py) for tensors whose pointers go
    straight to the C ABI on the grid / fused routes: CUDA, exact dtype, contiguous.  The reference raises on a dtype
    mismatch too (``tensor.data<float>()``)."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


# Backward kernels: "det" (default) = deterministic segmented reduction (scatter_det.cu); "atomic" = the reference's atomicAdd
# scatters (1:1 entry points g4d_*_grad), whose result depends on the order the atomics land in.
BACKWARD = os.environ.get("G4D_BACKWARD", "det")


def scatter_add_deterministic(dst: torch.Tensor, n_dst: int, grad: torch.Tensor, weight: torch.Tensor = None, grad_div: int = 1) -> torch.Tensor:
    """out[b, c, dst[b, e]] += weight[b, e] * grad[b, c, e // grad_div], summed in ascending e: the three backward scatters of the
    reference (group / gather / three_interpolate) as one deterministic segmented reduction.  dst (B, n_src) int32 (any trailing
    shape, flattened), grad (B, C, n_src // grad_div) fp32 -> (B, C, n_dst).  The index structure is cached on `dst`."""
    B = dst.shape[0]
    dflat = dst.reshape(B, -1)
    n_src = dflat.shape[1]
    C = grad.shape[1]
    grad = grad.reshape(B, C, -1)
    _chk(dflat, torch.int32, "dst"); _chk(grad, torch.float32, "grad")
    if weight is not None:
        weight = _chk(weight.reshape(B, -1), torch.float32, "weight")
    L = _lib.lib()
    hit = getattr(dst, "_g4d_csr", None)
    if hit is None or hit[1] != dst._version or hit[2] != n_dst:
        ws = torch.empty(L.g4d_scatter_det_workspace_bytes(B, n_src, n_dst) // 4, dtype=torch.int32, device=dst.device)
        _lib.check(L.g4d_scatter_det_build(B, n_src, n_dst, _lib.ptr(dflat), _lib.ptr(ws), _lib.stream_ptr()), "g4d_scatter_det_build")
        hit = (ws, dst._version, n_dst)
        try:
            dst._g4d_csr = hit
        except Exception:
            pass
    out = torch.empty(B, C, n_dst, dtype=torch.float32, device=grad.device)
    rc = L.g4d_scatter_det_apply(B, C, n_src, n_dst, grad_div, _lib.ptr(hit[0]), _lib.ptr(weight), _lib.ptr(grad), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "g4d_scatter_det_apply")
    return out


GRID_MIN_POINTS = 2048     # below this the brute-force scans are already cheap (measured at 1024 points, 256 queries, r = 0.1 / 0.2 on body clouds: brute force 0.145 ms, grid route with its two builds 0.194 ms)
FPS_PRUNE_MIN_POINTS = 2048
# "rows": Morton-ordered warp-row pruned kernel (fps_rows.cu, default); "pruned": the cell-sorted thread-per-clump kernel of
# round 1 (fps_pruned.cu, needs a grid, n <= 8192); "plain": no pruning (fps.cu)
FPS_KERNEL = os.environ.get("G4D_FPS", "rows")


QUERY_ORDER = os.environ.get("G4D_BQ_QUERY_ORDER", "1") != "0"
# three_nn: cells along the longest axis of the grid over the m known points = NN_CELLS_FACTOR * sqrt(m)
NN_CELLS_FACTOR = float(os.environ.get("G4D_NN_CELLS", "1.2"))    # swept on B200 (tools/nn_sweep.py): 0.7 -> 1.2 is 15 % faster on body scans, even on cube clouds


def build_grid(xyz: torch.Tensor, min_cell: float) -> torch.Tensor:
    """Cell-sorted copy of every cloud (g4d_grid_build).  min_cell > 0: cell edge >= min_cell; min_cell <= -1: that many
    cells along the longest axis.  The grid is remembered on the tensor (``xyz._g4d_grid``) for later searches."""
    assert xyz.is_contiguous() and xyz.is_cuda and xyz.dtype == torch.float32
    B, N, _ = xyz.shape
    L = _lib.lib()
    grid = torch.empty(L.g4d_grid_bytes(B, N) // 4, dtype=torch.float32, device=xyz.device)
    _lib.check(L.g4d_grid_build(B, N, _lib.ptr(xyz), float(min_cell), _lib.ptr(grid), _lib.stream_ptr()), "g4d_grid_build")
    try:
        xyz._g4d_grid = (grid, float(min_cell), xyz._version)
    except Exception:
        pass
    return grid


def _cached_grid(xyz, need_cell=None):
    """A grid previously built over this very tensor (same version counter); need_cell: minimum cell edge required."""
    hit = getattr(xyz, "_g4d_grid", None)
    if hit is None or hit[2] != xyz._version:
        return None
    if need_cell is not None and not (hit[1] >= need_cell):
        return None
    return hit[0]


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        """xyz (B,N,3) -> (B,npoint) int32 indices; idx[:,0] == 0 (pointnet2_utils.py:10-29)."""
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        output = _i32(B, npoint, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        pointnet2.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, output)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


def furthest_point_sample_and_gather(xyz: torch.Tensor, npoint: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused FPS + gather: (idx (B,npoint) int32, new_xyz (B,npoint,3)).  Same indices as furthest_point_sample;
    new_xyz == gather_operation(xyz^T, idx)^T bit for bit (a copy).  No gradient flows to xyz (as in the reference,
    where new_xyz comes out of gather_operation applied to a non-differentiable index)."""
    assert xyz.is_contiguous() and xyz.dtype == torch.float32 and xyz.is_cuda
    B, N, _ = xyz.size()
    idx = _i32(B, npoint, device=xyz.device)
    new_xyz = _f32(B, npoint, 3, device=xyz.device)
    if FPS_KERNEL == "rows" and FPS_PRUNE_MIN_POINTS <= N <= 16384:
        L = _lib.lib()
        ws = torch.empty(L.g4d_fps_workspace_bytes(B, N) // 4, dtype=torch.float32, device=xyz.device)
        rc = L.g4d_fps_gather_ws(B, N, npoint, _lib.ptr(xyz), _lib.ptr(idx), _lib.ptr(new_xyz), _lib.ptr(ws), _lib.stream_ptr())
        _lib.check(rc, "g4d_fps_gather_ws")
        return idx, new_xyz
    if FPS_KERNEL == "pruned" and FPS_PRUNE_MIN_POINTS <= N <= 8192:
        # exact spatial pruning over the cell-sorted order (any grid over xyz will do; an SA module builds it with its
        # largest ball-query radius first, so the ball query that follows re-uses it)
        grid = _cached_grid(xyz)
        if grid is None:
            grid = build_grid(xyz, -24.0)
        rc = _lib.lib().g4d_fps_gather_grid(B, N, npoint, _lib.ptr(grid), _lib.ptr(idx), _lib.ptr(new_xyz), _lib.stream_ptr())
        _lib.check(rc, "g4d_fps_gather_grid")
        return idx, new_xyz
    scratch = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device) if N > 16384 else None
    rc = _lib.lib().g4d_fps_gather(B, N, npoint, _lib.ptr(xyz), _lib.ptr(idx), _lib.ptr(new_xyz), _lib.ptr(scratch),
                                   _lib.stream_ptr())
    _lib.check(rc, "g4d_fps_gather")
    return idx, new_xyz


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """(B,C,N), (B,npoint) -> (B,C,npoint)  (pointnet2_utils.py:42-60)"""
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = _f32(B, C, npoint, device=features.device)
        pointnet2.gather_points_wrapper(B, C, N, npoint, features, idx, output)
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        if BACKWARD == "det":
            return scatter_add_deterministic(idx, N, grad_out.contiguous()), None
        grad_features = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        pointnet2.gather_points_grad_wrapper(B, C, N, npoint, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (dist (B,n,3) = sqrt of squared distances, idx (B,n,3) int32)  (pointnet2_utils.py:78-98)"""
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = _f32(B, N, 3, device=unknown.device)
        idx = _i32(B, N, 3, device=unknown.device)
        three_nn_raw(unknown, known, dist2, idx)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


def three_nn_raw(unknown, known, dist2, idx):
    """Fills dist2 (SQUARED distances) and idx like the reference kernel; large problems go through the uniform grid
    (identical results), small ones through the 1:1 replacement of three_nn_kernel_fast."""
    B, N, _ = unknown.size()
    m = known.size(1)
    if N >= GRID_MIN_POINTS and 512 <= m <= (1 << 20):
        _chk(unknown, torch.float32, "unknown"); _chk(known, torch.float32, "known")
        _chk(dist2, torch.float32, "dist2"); _chk(idx, torch.int32, "idx")
        kgrid = build_grid(known, -max(4.0, round(NN_CELLS_FACTOR * float(m) ** 0.5)))
        ugrid = _cached_grid(unknown)      # only a processing order: any grid over `unknown` will do
        rc = _lib.lib().g4d_three_nn_grid(B, N, m, _lib.ptr(unknown), _lib.ptr(kgrid), _lib.ptr(ugrid), _lib.ptr(dist2),
                                          _lib.ptr(idx), _lib.stream_ptr())
        _lib.check(rc, "g4d_three_nn_grid")
    else:
        pointnet2.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        """features (B,C,m), idx/weight (B,n,3) -> (B,C,n)  (pointnet2_utils.py:110-131)"""
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = _f32(B, c, n, device=features.device)
        pointnet2.three_interpolate_wrapper(B, c, m, n, features, idx, weight, output)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        if BACKWARD == "det":
            return scatter_add_deterministic(idx, m, grad_out.contiguous(), weight=weight.contiguous(), grad_div=3), None, None
        grad_features = torch.zeros(B, c, m, dtype=torch.float32, device=grad_out.device)
        pointnet2.three_interpolate_grad_wrapper(B, c, n, m, grad_out.contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample)  (pointnet2_utils.py:158-176)"""
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = _f32(B, C, nfeatures, nsample, device=features.device)
        pointnet2.group_points_wrapper(B, C, N, nfeatures, nsample, features, idx, output)
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        if BACKWARD == "det":
            return scatter_add_deterministic(idx, N, grad_out.contiguous()), None
        grad_features = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        pointnet2.group_points_grad_wrapper(B, C, N, npoint, nsample, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        """-> idx (B,npoint,nsample) int32; all-zero row when no point is in range  (pointnet2_utils.py:202-222)"""
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.zeros(B, npoint, nsample, dtype=torch.int32, device=xyz.device)
        if GRID_MIN_POINTS <= N <= 65536:
            _chk(xyz, torch.float32, "xyz"); _chk(new_xyz, torch.float32, "new_xyz")
            grid = _cached_grid(xyz, radius)
            if grid is None:
                grid = build_grid(xyz, radius)
            qgrid = None                                     # processing order of the queries (see ball_query_pair)
            if QUERY_ORDER and npoint >= 256:
                qgrid = _cached_grid(new_xyz)
                if qgrid is None:
                    qgrid = build_grid(new_xyz, radius)
            rc = _lib.lib().g4d_ball_query2_grid_ordered(B, N, npoint, float(radius), nsample, _lib.ptr(idx), 0.0, 0, None,
                                                         _lib.ptr(new_xyz), _lib.ptr(grid),
                                                         _lib.ptr(qgrid) if qgrid is not None else None, _lib.stream_ptr())
            _lib.check(rc, "g4d_ball_query2_grid_ordered")
        else:
            pointnet2.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


def ball_query_pair(radius0, nsample0, radius1, nsample1, xyz, new_xyz):
    """Both scales of an MSG module from one scan of the cloud (g4d_ball_query2); each result equals ball_query's."""
    _chk(xyz, torch.float32, "xyz"); _chk(new_xyz, torch.float32, "new_xyz")
    B, N, _ = xyz.size()
    P = new_xyz.size(1)
    idx0 = torch.zeros(B, P, nsample0, dtype=torch.int32, device=xyz.device)
    idx1 = torch.zeros(B, P, nsample1, dtype=torch.int32, device=xyz.device)
    if GRID_MIN_POINTS <= N <= 65536:
        rmax = max(float(radius0), float(radius1))
        grid = _cached_grid(xyz, rmax)
        if grid is None:
            grid = build_grid(xyz, rmax)
        # FPS hands the centroids over far apart from one another: walk them in the cell order of a small grid of their own
        qgrid = None
        if QUERY_ORDER and P >= 256:
            qgrid = _cached_grid(new_xyz) if _cached_grid(new_xyz) is not None else build_grid(new_xyz, rmax)
        rc = _lib.lib().g4d_ball_query2_grid_ordered(B, N, P, float(radius0), nsample0, _lib.ptr(idx0), float(radius1), nsample1,
                                                     _lib.ptr(idx1), _lib.ptr(new_xyz), _lib.ptr(grid),
                                                     _lib.ptr(qgrid) if qgrid is not None else None, _lib.stream_ptr())
        _lib.check(rc, "g4d_ball_query2_grid_ordered")
    else:
        rc = _lib.lib().g4d_ball_query2(B, N, P, float(radius0), nsample0, _lib.ptr(idx0), float(radius1), nsample1,
                                        _lib.ptr(idx1), _lib.ptr(new_xyz), _lib.ptr(xyz), _lib.stream_ptr())
        _lib.check(rc, "g4d_ball_query2")
    return idx0, idx1


_FUSED_NSAMPLE = (4, 8, 16, 32, 64, 128)


class _QueryAndGroupFused(Function):
    """QueryAndGroup.forward as the ball query plus ONE grouping pass (nsample % 4 == 0).  Backward reproduces the reference graph: features and xyz receive the
    scatter-add of the grouped gradient (GroupingOperation.backward), new_xyz receives minus the sum over samples
    (the in-place ``grouped_xyz -= new_xyz`` of pointnet2_utils.py:253)."""

    @staticmethod
    def forward(ctx, radius, nsample, use_xyz, xyz, new_xyz, features):
        _chk(xyz, torch.float32, "xyz"); _chk(new_xyz, torch.float32, "new_xyz")
        if features is not None:
            _chk(features, torch.float32, "features")
        B, N, _ = xyz.size()
        P = new_xyz.size(1)
        C = 0 if features is None else features.size(1)
        cout = (3 if use_xyz else 0) + C
        out = _f32(B, cout, P, nsample, device=xyz.device)
        if C >= 16 and B <= 65535:
            # ball query, then the grouping pass from a POINT-major fp32 copy of the features (made once per feature tensor and
            # remembered on it): 256-byte row gathers, shared-memory transpose, 256-byte runs per output channel
            idx = BallQuery.apply(radius, nsample, xyz, new_xyz)
            hit = getattr(features, "_g4d_pm32", None)
            if hit is None or hit[1] != features._version:
                hit = (features.detach().transpose(1, 2).contiguous(), features._version)
                try:
                    features._g4d_pm32 = hit
                except Exception:
                    pass
            rc = _lib.lib().g4d_group_fused_pm(B, N, P, C, nsample, int(use_xyz), _lib.ptr(xyz), _lib.ptr(new_xyz),
                                               _lib.ptr(hit[0]), _lib.ptr(idx), _lib.ptr(out), _lib.stream_ptr())
            _lib.check(rc, "g4d_group_fused_pm")
        else:
            # ball query (uniform grid for large clouds), then ONE grouping pass (16-byte stores) for xyz, features and the cat
            idx = BallQuery.apply(radius, nsample, xyz, new_xyz)
            rc = _lib.lib().g4d_group_fused(B, N, P, C, nsample, int(use_xyz), _lib.ptr(xyz), _lib.ptr(new_xyz),
                                            _lib.ptr(features), _lib.ptr(idx), _lib.ptr(out), _lib.stream_ptr())
            _lib.check(rc, "g4d_group_fused")
        ctx.meta = (idx, N, C, use_xyz)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, N, C, use_xyz = ctx.meta
        B, _, P, S = grad_out.size()
        grad_out = grad_out.contiguous()
        g_xyz = g_new = g_feat = None
        off = 0
        if use_xyz:
            gx = grad_out[:, :3].contiguous()
            if ctx.needs_input_grad[3]:
                if BACKWARD == "det":
                    g = scatter_add_deterministic(idx, N, gx)
                else:
                    g = torch.zeros(B, 3, N, dtype=torch.float32, device=grad_out.device)
                    pointnet2.group_points_grad_wrapper(B, 3, N, P, S, gx, idx, g)
                g_xyz = g.transpose(1, 2).contiguous()
            if ctx.needs_input_grad[4]:
                g_new = -gx.sum(dim=3).transpose(1, 2).contiguous()
            off = 3
        if C > 0 and ctx.needs_input_grad[5]:
            gf = grad_out[:, off:].contiguous()
            if BACKWARD == "det":
                g_feat = scatter_add_deterministic(idx, N, gf)
            else:
                g_feat = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
                pointnet2.group_points_grad_wrapper(B, C, N, P, S, gf, idx, g_feat)
        return None, None, None, g_xyz, g_new, g_feat


class QueryAndGroup(nn.Module):
    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        """radius of the ball, maximum number of neighbours, whether to prepend the relative xyz (pointnet2_utils.py:233-240)"""
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None) -> torch.Tensor:
        """xyz (B,N,3), new_xyz (B,npoint,3), features (B,C,N) or None -> (B, 3+C, npoint, nsample)"""
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
        f32cuda = lambda t: t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        if self.nsample in _FUSED_NSAMPLE and f32cuda(xyz) and f32cuda(new_xyz) and (features is None or f32cuda(features)):
            return _QueryAndGroupFused.apply(self.radius, self.nsample, self.use_xyz, xyz, new_xyz, features)
        # generic composition, operator by operator (pointnet2_utils.py:250-265)
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping_operation(features.contiguous(), idx)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        return grouped_xyz


class GroupAll(nn.Module):
    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        """xyz (B,N,3), features (B,C,N) -> (B, C+3, 1, N); new_xyz ignored (pointnet2_utils.py:273-291)"""
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        return grouped_xyz
