"""Set-abstraction / feature-propagation modules with the reference's interface
(modules/pointnet2/pointnet2/pointnet2_modules.py): same constructor keywords, same sub-module and parameter
names (``groupers``, ``mlps.{i}.layer{j}.conv.weight`` ...), same ``forward`` signatures and results.

``_PointnetSAModuleBase.forward(xyz, features=None, new_xyz=None) -> (new_xyz, new_features)`` has two routes:

* fused (eval mode, no autograd graph needed through the MLP, max-pool, 3-layer conv+BN+ReLU stacks):
      g4d_fps_gather  ->  g4d_ball_query2 (both MSG scales from one scan)  ->  per scale g4d_sa_mlp_max
  (tcgen05 grouped-MLP + max, writes its channel window of the concatenated output directly).  This is the
  route the reference's posed-garment stage and all inference take: the encoder runs under ``torch.no_grad()``
  with BatchNorm in eval mode (modules/mesh_encoder.py:416-417, train_temporal.py:227-233).
* general (training-mode BN / gradients / avg-pool / other MLP depths): the reference's operator sequence,
  each operator on the B200 kernels (fused QueryAndGroup, then torch conv/bn/relu and pooling).
"""
import ctypes
import os
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from . import pointnet2_utils
from . import pytorch_utils as pt_utils


# Feature-propagation MLPs in eval mode: "tc" (default) = our two-layer tcgen05 kernel with streamed weights (g4d_mlp2_rows) on the
# point-major rows our prologue kernel writes -- no library GEMM; shapes it does not take fall through to "half" = one fp16
# library GEMM per layer over the whole batch with our bias+ReLU epilogues; "conv" = per-cloud fp32/TF32 library convolution.
_FP_GEMM = os.environ.get("G4D_FP_GEMM", "tc")


class _Mlp2Params:
    """Device-resident packed parameters of a 2-layer folded MLP for g4d_mlp2_rows (None-able: see mlp2_supported)."""

    def __init__(self, layers, device):
        (w1, b1), (w2, b2) = [(w.detach().float().cpu().contiguous(), b.detach().float().cpu().contiguous()) for w, b in layers]
        L = _lib.lib()
        self.desc = _lib.Mlp2Desc(w1.shape[1], w1.shape[0], w2.shape[0])
        nbytes = L.g4d_mlp2_param_bytes(ctypes.byref(self.desc))
        if nbytes == 0:
            raise _lib.G4DError("g4d_mlp2_param_bytes: " + L.g4d_last_error().decode())
        blob = torch.empty(nbytes, dtype=torch.uint8)
        rc = L.g4d_mlp2_pack_params(ctypes.byref(self.desc), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), blob.data_ptr())
        _lib.check(rc, "g4d_mlp2_pack_params")
        self.params = blob.to(device)
        torch.cuda.current_stream(device).synchronize()      # other streams may run this module next: the blob must have landed
        self.c_out = w2.shape[0]


def mlp2_supported(layers):
    if layers is None or len(layers) != 2:
        return False
    (w1, _), (w2, _) = layers
    c1, c_in = w1.shape
    c2 = w2.shape[0]
    return c_in % 32 == 0 and c_in >= 32 and c1 % 64 == 0 and 64 <= c1 <= 512 and c2 % 64 == 0 and 64 <= c2 <= 256


def _as_point_major_half(features: torch.Tensor) -> torch.Tensor:
    """(B,C,N) fp32 channel-major -> (B,N,C) fp16 point-major, the gather layout of the fused kernel.
    A fused SA module attaches this copy to its output as ``_g4d_pm`` so the next level does not recompute it."""
    pm = pointnet2_utils.point_major_of(features)     # trusted only while `features` is unmodified (version counter)
    if pm is not None:
        return pm
    return features.detach().transpose(1, 2).to(torch.float16).contiguous()


def _FP_GEMM_NOW():
    return _FP_GEMM          # module attribute: the tests switch routes by assigning it


class _FusedBranch:
    """Device-resident packed parameters of one (grouper, SharedMLP) scale for g4d_sa_mlp_max."""

    def __init__(self, mlp: pt_utils.SharedMLP, nsample: int, c_in: int, device):
        folded = pt_utils.fold_shared_mlp(mlp)
        assert folded is not None and len(folded) in (2, 3)
        folded = [(w.cpu().contiguous(), b.cpu().contiguous()) for w, b in folded]
        if len(folded) == 2:
            # two-layer MLP (the garment encoder's branches, mesh_encoder.py:58-70) on the three-layer kernel: an identity middle
            # layer is exact -- its input is post-ReLU fp16, the products 1*h accumulate exactly in fp32, ReLU and the rounding
            # back to fp16 change nothing
            c1 = folded[0][0].shape[0]
            folded = [folded[0], (torch.eye(c1, dtype=torch.float32), torch.zeros(c1, dtype=torch.float32)), folded[1]]
        (w1, b1), (w2, b2), (w3, b3) = folded
        assert w1.shape[1] == c_in + 3
        L = _lib.lib()
        self.desc = _lib.SaMlpDesc(c_in, w1.shape[0], w2.shape[0], w3.shape[0], nsample, L.g4d_sa_mlp_k0(c_in))
        nbytes = L.g4d_sa_mlp_param_bytes(ctypes.byref(self.desc))
        if nbytes == 0:
            raise _lib.G4DError("g4d_sa_mlp_param_bytes: " + L.g4d_last_error().decode())
        blob = torch.empty(nbytes, dtype=torch.uint8)
        rc = L.g4d_sa_mlp_pack_params(ctypes.byref(self.desc), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                      w3.data_ptr(), b3.data_ptr(), blob.data_ptr())
        _lib.check(rc, "g4d_sa_mlp_pack_params")        # raises when a folded weight / bias does not fit fp16
        self.params = blob.to(device)
        self.c_out = w3.shape[0]


def fused_branch_supported(mlp, grouper, c_in):
    if not isinstance(grouper, pointnet2_utils.QueryAndGroup) or not grouper.use_xyz:
        return False
    if grouper.nsample not in (8, 16, 32, 64, 128) or c_in % 8 != 0 or len(mlp) not in (2, 3):
        return False
    folded = pt_utils.fold_shared_mlp(mlp)
    if folded is None:
        return False
    widths = [w.shape[0] for w, _ in folded]
    c1, c2, c3 = widths if len(widths) == 3 else (widths[0], widths[0], widths[1])
    return c1 % 16 == 0 and c2 % 16 == 0 and 16 <= c1 <= 256 and 16 <= c2 <= 256 and c3 <= 256


class _PointnetSAModuleBase(nn.Module):

    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None
        self.pool_method = 'max_pool'
        self.fused = True          # set False to force the operator-by-operator route
        self._fused_cache = {}

    # ---- fused route ------------------------------------------------------------------------------------
    def _can_fuse(self, xyz, features, new_xyz):
        if not self.fused or self.training or self.pool_method != 'max_pool' or self.npoint is None:
            return False
        if not xyz.is_cuda or xyz.dtype != torch.float32:
            return False
        if torch.is_grad_enabled() and (xyz.requires_grad or (features is not None and features.requires_grad)
                                        or (new_xyz is not None and new_xyz.requires_grad)
                                        or any(p.requires_grad for p in self.mlps.parameters())):
            return False
        c_in = 0 if features is None else features.shape[1]
        return all(self._branch(i, c_in, xyz.device) is not None for i in range(len(self.groupers)))

    def _branch(self, i, c_in, device):
        """Packed parameters of scale i (None if that scale cannot take the fused route); rebuilt when any
        weight / BN statistic changes (load_state_dict, optimizer step, .to())."""
        key = (i, c_in, str(device))
        ver = pt_utils.shared_mlp_version(self.mlps[i])
        hit = self._fused_cache.get(key)
        if hit is None or hit[0] != ver:
            br = None
            if fused_branch_supported(self.mlps[i], self.groupers[i], c_in):
                try:
                    br = _FusedBranch(self.mlps[i], self.groupers[i].nsample, c_in, device)
                except _lib.G4DError as e:
                    if "fp16 range" not in str(e) and "shared memory footprint" not in str(e):
                        raise
                    br = None       # fp16 operands cannot hold these folded weights, or the resident weights do not fit shared
                                    # memory (e.g. 256-wide two-layer branches): operator route (fp32 / TF32 layers)
            hit = (ver, br)
            self._fused_cache[key] = hit
        return hit[1]

    def _forward_fused(self, xyz, features, new_xyz):
        B, N, _ = xyz.shape
        xyz = xyz.contiguous()
        if N >= pointnet2_utils.GRID_MIN_POINTS and pointnet2_utils._cached_grid(xyz, max(g.radius for g in self.groupers)) is None:
            pointnet2_utils.build_grid(xyz, max(g.radius for g in self.groupers))      # shared by FPS pruning and the ball query
        if new_xyz is None:
            _, new_xyz = pointnet2_utils.furthest_point_sample_and_gather(xyz, self.npoint)
        else:
            new_xyz = new_xyz.contiguous()
        P = new_xyz.shape[1]
        c_in = 0 if features is None else features.shape[1]
        feat_pm = None if features is None else _as_point_major_half(features)
        ng = len(self.groupers)
        if ng == 2:
            g0, g1 = self.groupers
            idxs = pointnet2_utils.ball_query_pair(g0.radius, g0.nsample, g1.radius, g1.nsample, xyz, new_xyz)
        else:
            idxs = [pointnet2_utils.ball_query(g.radius, g.nsample, xyz, new_xyz) for g in self.groupers]
        branches = [self._branch(i, c_in, xyz.device) for i in range(ng)]
        ctot = sum(br.c_out for br in branches)
        out_cm = torch.empty(B, ctot, P, dtype=torch.float32, device=xyz.device)
        out_pm = torch.empty(B, P, ctot, dtype=torch.float16, device=xyz.device)
        L = _lib.lib()
        off = 0
        for br, idx in zip(branches, idxs):
            rc = L.g4d_sa_mlp_max(ctypes.byref(br.desc), _lib.ptr(br.params), B, N, P, _lib.ptr(xyz), _lib.ptr(new_xyz),
                                  _lib.ptr(idx), _lib.ptr(feat_pm), _lib.ptr(out_cm), _lib.ptr(out_pm), ctot, off,
                                  _lib.stream_ptr())
            _lib.check(rc, "g4d_sa_mlp_max")
            off += br.c_out
        pointnet2_utils.attach_point_major(out_cm, out_pm)
        return new_xyz, out_cm

    # ---- public forward ----------------------------------------------------------------------------------
    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None, new_xyz=None) -> (torch.Tensor, torch.Tensor):
        """xyz (B,N,3), features (B,C,N) or None -> new_xyz (B,npoint,3), new_features (B, sum_k mlps[k][-1], npoint)"""
        if self._can_fuse(xyz, features, new_xyz):
            return self._forward_fused(xyz, features, new_xyz)

        new_features_list = []
        if new_xyz is None and self.npoint is not None:
            if xyz.requires_grad and torch.is_grad_enabled():
                # keep the reference graph: gather_operation is differentiable w.r.t. xyz (pointnet2_modules.py:30-35)
                xyz_flipped = xyz.transpose(1, 2).contiguous()
                new_xyz = pointnet2_utils.gather_operation(
                    xyz_flipped, pointnet2_utils.furthest_point_sample(xyz.contiguous(), self.npoint)
                ).transpose(1, 2).contiguous()
            else:
                _, new_xyz = pointnet2_utils.furthest_point_sample_and_gather(xyz.contiguous(), self.npoint)
        for i in range(len(self.groupers)):
            new_features = self.groupers[i](xyz, new_xyz, features)       # (B, C, npoint, nsample)
            new_features = self.mlps[i](new_features)                      # (B, mlp[-1], npoint, nsample)
            if self.pool_method == 'max_pool':
                new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            elif self.pool_method == 'avg_pool':
                new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
            else:
                raise NotImplementedError
            new_features_list.append(new_features.squeeze(-1))             # (B, mlp[-1], npoint)
        return new_xyz, torch.cat(new_features_list, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Set abstraction with multi-scale grouping (pointnet2_modules.py:58-91)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]], bn: bool = True,
                 use_xyz: bool = True, pool_method='max_pool', instance_norm=False):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        for radius, nsample, mlp_spec in zip(radii, nsamples, mlps):
            self.groupers.append(pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz)
                                 if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            if use_xyz:
                mlp_spec[0] += 3           # in place, like the reference (callers observe the mutated list)
            self.mlps.append(pt_utils.SharedMLP(mlp_spec, bn=bn, instance_norm=instance_norm))
        self.pool_method = pool_method


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (pointnet2_modules.py:94-113)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pool_method='max_pool', instance_norm=False):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz,
                         pool_method=pool_method, instance_norm=instance_norm)


class PointnetFPModule(nn.Module):
    """Feature propagation: 3-NN inverse-distance interpolation + skip concat + SharedMLP (pointnet2_modules.py:116-156)."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)

    def forward(self, unknown: torch.Tensor, known: torch.Tensor, unknow_feats: torch.Tensor,
                known_feats: torch.Tensor) -> torch.Tensor:
        """unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m) -> (B, mlp[-1], n)"""
        folded = self._folded_mlp(known_feats, unknow_feats)
        L = _lib.lib() if folded is not None else None
        if (folded is not None and known is not None and known_feats.dtype == torch.float32
                and (unknow_feats is None or unknow_feats.dtype == torch.float32)):
            # eval route: three_nn -> ONE kernel for weights + three_interpolate + cat
            unknown, known, known_feats = unknown.contiguous(), known.contiguous(), known_feats.contiguous()
            B, n, _ = unknown.shape
            m, c2 = known.shape[1], known_feats.shape[1]
            c1 = 0 if unknow_feats is None else unknow_feats.shape[1]
            skip = None if unknow_feats is None else unknow_feats.contiguous()
            dev = unknown.device
            dist2 = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
            idx = torch.empty(B, n, 3, dtype=torch.int32, device=dev)
            pointnet2_utils.three_nn_raw(unknown, known, dist2, idx)
            if _FP_GEMM_NOW() in ("half", "tc") and folded["half"] is not None and (B * n) % 8 == 0 and B <= 65535:
                # The module's 1x1 convolutions as ONE library GEMM per layer over the whole batch: the concatenated input is
                # written fp16 in (C, B*n) layout, hidden layers stay fp16 (fp32 accumulation; same 11-bit operand precision
                # as the TF32 convolutions torch runs by default), bias+ReLU passes are ours, the last layer's GEMM returns
                # fp32 and its epilogue writes the reference layout (B, C, n) (+ the fp16 point-major copy for the next level).
                kpm = pointnet2_utils.point_major_of(known_feats)      # fp16 point-major copy emitted by the producing level
                spm = None if unknow_feats is None else pointnet2_utils.point_major_of(unknow_feats)
                pm_ok = lambda t, shape: (t is not None and t.dtype == torch.float16 and tuple(t.shape) == shape and t.is_contiguous())
                layers = folded["half"]
                if (pm_ok(kpm, (B, m, c2)) and (c1 == 0 or pm_ok(spm, (B, n, c1))) and c2 % 8 == 0 and c1 % 8 == 0
                        and all(w.shape[0] % 8 == 0 for w, _ in layers[:-1])):
                    # point-major all the way: x (B*n, C) rows -> x @ W^T per layer (the library's linear-layer form)
                    x = torch.empty(B * n, c2 + c1, dtype=torch.float16, device=dev)
                    rc = L.g4d_fp_interp_concat_rows_h(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx), _lib.ptr(kpm), _lib.ptr(spm),
                                                       _lib.ptr(x), _lib.stream_ptr())
                    _lib.check(rc, "g4d_fp_interp_concat_rows_h")
                    tc = folded.get("tc") if _FP_GEMM_NOW() == "tc" else None
                    if tc is not None:
                        # both 1x1 convolutions + bias + ReLU in ONE tcgen05 kernel, weights streamed through shared memory
                        out = torch.empty(B, tc.c_out, n, dtype=torch.float32, device=dev)
                        pm = torch.empty(B, n, tc.c_out, dtype=torch.float16, device=dev) if self.emit_point_major else None
                        rc = L.g4d_mlp2_rows(ctypes.byref(tc.desc), _lib.ptr(tc.params), B, n, _lib.ptr(x), _lib.ptr(out), _lib.ptr(pm),
                                             _lib.stream_ptr())
                        _lib.check(rc, "g4d_mlp2_rows")
                        if pm is not None:
                            pointnet2_utils.attach_point_major(out, pm)
                        return out
                    for li, (w16, b) in enumerate(layers):
                        if li < len(layers) - 1:
                            x = F.linear(x, w16)
                            rc = L.g4d_bias_relu_rows_h(B * n, w16.shape[0], _lib.ptr(x), _lib.ptr(b), 1, _lib.stream_ptr())
                            _lib.check(rc, "g4d_bias_relu_rows_h")
                        else:
                            y = torch.mm(x, w16.t(), out_dtype=torch.float32)
                            out = torch.empty(B, w16.shape[0], n, dtype=torch.float32, device=dev)
                            pm = torch.empty(B, n, w16.shape[0], dtype=torch.float16, device=dev) if self.emit_point_major else None
                            rc = L.g4d_bias_relu_rows_unpack(B, w16.shape[0], n, _lib.ptr(y), _lib.ptr(b), 1, _lib.ptr(out), _lib.ptr(pm),
                                                             _lib.stream_ptr())
                            _lib.check(rc, "g4d_bias_relu_rows_unpack")
                            if pm is not None:
                                pointnet2_utils.attach_point_major(out, pm)
                    return out
                x = torch.empty(c2 + c1, B * n, dtype=torch.float16, device=dev)
                if (kpm is not None and c2 % 8 == 0 and kpm.dtype == torch.float16 and tuple(kpm.shape) == (B, m, c2)
                        and kpm.is_contiguous()):
                    rc = L.g4d_fp_interp_concat_pm_cbn_h(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx), _lib.ptr(kpm), _lib.ptr(skip),
                                                         _lib.ptr(x), _lib.stream_ptr())
                else:
                    rc = L.g4d_fp_interp_concat_cbn_h(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx), _lib.ptr(known_feats), _lib.ptr(skip),
                                                      _lib.ptr(x), _lib.stream_ptr())
                _lib.check(rc, "g4d_fp_interp_concat_cbn_h")
                for li, (w16, b) in enumerate(layers):
                    if li < len(layers) - 1:
                        x = torch.mm(w16, x)
                        rc = L.g4d_bias_relu_h(w16.shape[0], B * n, _lib.ptr(x), _lib.ptr(b), 1, _lib.stream_ptr())
                        _lib.check(rc, "g4d_bias_relu_h")
                    else:
                        y = torch.mm(w16, x, out_dtype=torch.float32)
                        out = torch.empty(B, w16.shape[0], n, dtype=torch.float32, device=dev)
                        pm = torch.empty(B, n, w16.shape[0], dtype=torch.float16, device=dev) if self.emit_point_major else None
                        rc = L.g4d_bias_relu_unpack(B, w16.shape[0], n, _lib.ptr(y), 0, _lib.ptr(b), 1, _lib.ptr(out), _lib.ptr(pm),
                                                    _lib.stream_ptr())
                        _lib.check(rc, "g4d_bias_relu_unpack")
                        if pm is not None:
                            pointnet2_utils.attach_point_major(out, pm)
                return out
            new_features = torch.empty(B, c2 + c1, n, dtype=torch.float32, device=dev)
            rc = L.g4d_fp_interp_concat(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx), _lib.ptr(known_feats), _lib.ptr(skip),
                                        _lib.ptr(new_features), _lib.stream_ptr())
            _lib.check(rc, "g4d_fp_interp_concat")
        else:
            if known is not None:
                dist, idx = pointnet2_utils.three_nn(unknown.contiguous(), known.contiguous())
                dist_recip = 1.0 / (dist + 1e-8)
                norm = torch.sum(dist_recip, dim=2, keepdim=True)
                weight = dist_recip / norm
                interpolated_feats = pointnet2_utils.three_interpolate(known_feats.contiguous(), idx, weight)
            else:
                interpolated_feats = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
            if unknow_feats is not None:
                new_features = torch.cat([interpolated_feats, unknow_feats], dim=1)
            else:
                new_features = interpolated_feats
        if folded is not None:
            # eval mode, no autograd, per-cloud route (G4D_FP_GEMM=conv): BatchNorm folded into the 1x1 convolutions -> library
            # convolution (no bias) + ONE in-place bias+ReLU pass of ours per layer; the last layer's pass also emits the fp16
            # point-major copy that the next (finer) level's fused kernel gathers from (attached as ``_g4d_pm``)
            y = new_features
            layers = folded["conv"]
            for li, (w, b) in enumerate(layers):
                last = li == len(layers) - 1
                y = F.conv2d(y.unsqueeze(-1), w).squeeze(-1)
                if last and self.emit_point_major and y.is_contiguous() and y.shape[0] <= 65535:
                    pm = torch.empty(y.shape[0], y.shape[2], y.shape[1], dtype=torch.float16, device=y.device)
                    rc = L.g4d_bias_relu_pm(y.shape[0], y.shape[1], y.shape[2], _lib.ptr(y), _lib.ptr(b), 1, _lib.ptr(pm), _lib.stream_ptr())
                    _lib.check(rc, "g4d_bias_relu_pm")
                    pointnet2_utils.attach_point_major(y, pm)
                elif y.shape[0] * y.shape[1] <= 65535 and y.is_contiguous():
                    rc = L.g4d_bias_relu_inplace(y.shape[0], y.shape[1], y.shape[2], _lib.ptr(y), _lib.ptr(b), 1, _lib.stream_ptr())
                    _lib.check(rc, "g4d_bias_relu_inplace")
                else:
                    y = F.relu_(y + b[None, :, None])
            return y
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)

    emit_point_major = False      # set by the encoder on the level that feeds the fused finest-level kernel
    fused = True                  # set False to force the reference's operator sequence (three_nn, three_interpolate, cat, SharedMLP)

    def _folded_mlp(self, x, skip=None):
        if not self.fused or self.training or not x.is_cuda or (torch.is_grad_enabled() and (x.requires_grad or (skip is not None and skip.requires_grad)
                                                                           or any(p.requires_grad for p in self.mlp.parameters()))):
            return None
        ver = pt_utils.shared_mlp_version(self.mlp)
        hit = getattr(self, "_fold_cache", None)
        if hit is None or hit[0] != ver:
            f = pt_utils.fold_shared_mlp(self.mlp)
            hit = (ver, None)
            if f is not None:
                half = [(w.to(torch.float16).contiguous(), b) for w, b in f]
                if not all(bool(torch.isfinite(w).all()) for w, _ in half):     # folded weight outside the fp16 range: keep fp32
                    half = None
                tc = None
                if half is not None and x.is_cuda and mlp2_supported(f):
                    tc = _Mlp2Params(f, x.device)           # (packed on the host, synchronised: safe to use from any stream)
                hit = (ver, {"conv": [(w[:, :, None, None].contiguous(), b) for w, b in f], "half": half, "tc": tc})
            self._fold_cache = hit
            if x.is_cuda:
                # the folded tensors were produced by asynchronous ops on THIS stream; the runner calls the module from several
                # streams: wait until they exist before any other stream may pick them up from the cache
                torch.cuda.current_stream(x.device).synchronize()
        return hit[1]
