"""Callers of the hot path inside the garment model, on the B200 kernels, with the reference's
function signatures (modules/mesh_encoder.py).

  calc_segmentation_results   PCAGarmentEncoderSeg.calc_segmentation_results (mesh_encoder.py:109-125): garment point selection.
                              The reference loops over the B*T frames in Python with boolean-mask indexing (one device->host
                              sync per frame); here ONE kernel (g4d_select_points) does the arg-max over the class logits and an
                              order-preserving stream compaction for all frames.
  GarmentEncoderStack         the second set-abstraction stack on the selected garment points (mesh_encoder.py:54-78, 149-161):
                              two PointnetSAModuleMSG levels + the GroupAll summary module -- the fused tcgen05 route of
                              garment4d_b200.pointnet2 in eval mode.
"""
import torch
import torch.nn as nn

from . import _lib
from .pointnet2.pointnet2_modules import PointnetSAModule, PointnetSAModuleMSG

# utils/dataloader.py:24-33 (class_num, label_dict): the garment's class id is label_dict[name] - 1 (mesh_encoder.py:117)
CLASS_NUM = 7
LABEL_DICT = {"Skin": 1, "Top": 2, "Tshirt": 2, "Dress": 3, "Skirt": 4, "Trousers": 5, "Jumpsuit": 6}


def calc_segmentation_results(x, sem_logits, n, nbatch, T, feature, garment_label, return_counts=False):
    """x (.., N, 3) coordinates, sem_logits (.., N, classes), feature (nbatch*T, Cf, N) channel-major (l_features[0]);
    returns (garment_v (nbatch*T, n, 3), feat (nbatch*T, n, Cf)): per frame the points labelled `garment_label`
    (= label_dict[name] - 1 in the reference), first n in point order, zero-padded (mesh_encoder.py:109-125)."""
    C = nbatch * T
    x = x.reshape(C, -1, 3)
    N = x.shape[1]
    ncls = sem_logits.shape[-1]
    sem_logits = sem_logits.reshape(C, N, ncls)
    if not (x.is_cuda and sem_logits.is_cuda and feature.is_cuda):
        raise _lib.G4DError("calc_segmentation_results runs on CUDA tensors only; there is no CPU fallback")
    xs = x.detach().contiguous().float()
    lg = sem_logits.detach().contiguous().float()
    ft = feature.detach().contiguous().float()
    assert ft.shape[0] == C and ft.shape[2] == N
    Cf = ft.shape[1]
    garment_v = torch.empty(C, n, 3, dtype=torch.float32, device=x.device)
    feat = torch.empty(C, n, Cf, dtype=torch.float32, device=x.device)
    counts = torch.empty(C, dtype=torch.int32, device=x.device) if return_counts else None
    rc = _lib.lib().g4d_select_points(C, N, ncls, Cf, int(garment_label), n, _lib.ptr(lg), None, _lib.ptr(xs), _lib.ptr(ft),
                                      _lib.ptr(garment_v), _lib.ptr(feat), _lib.ptr(counts), _lib.stream_ptr())
    _lib.check(rc, "g4d_select_points")
    return (garment_v, feat, counts) if return_counts else (garment_v, feat)


class GarmentEncoderStack(nn.Module):
    """GarmentEncoder + GarmentSummarize of PCAGarmentEncoderSeg (mesh_encoder.py:54-78), same sub-module names and therefore
    the same state-dict keys; forward = mesh_encoder.py:149-161: (garment_v (C,n,3), garment_f (C,Cf,n)) ->
    (l_xyz [3], l_features [3], garment_summary (C, 512))."""

    def __init__(self, feat_channels=64):
        super().__init__()
        self.GarmentEncoder = nn.ModuleList()
        self.GarmentEncoder.append(PointnetSAModuleMSG(npoint=512, radii=[0.05, 0.1], nsamples=[16, 32],
                                                       mlps=[[feat_channels, 32, 32], [feat_channels, 64, 64]], use_xyz=True, bn=True))
        self.GarmentEncoder.append(PointnetSAModuleMSG(npoint=64, radii=[0.2, 0.4], nsamples=[32, 64],
                                                       mlps=[[32 + 64, 128, 128], [32 + 64, 256, 256]], use_xyz=True, bn=True))
        self.GarmentSummarize = PointnetSAModule(mlp=[128 + 256, 512, 512], use_xyz=True, bn=True)

    def forward(self, garment_v, garment_f):
        l_xyz, l_features = [garment_v], [garment_f]
        for sa in self.GarmentEncoder:
            li_xyz, li_features = sa(l_xyz[-1], l_features[-1])
            l_xyz.append(li_xyz)
            l_features.append(li_features)
        summary = self.GarmentSummarize(l_xyz[-1], l_features[-1])[1]
        return l_xyz, l_features, summary.reshape(summary.shape[0], -1)


# ------------------------------------------------------------------------------------------------------------------
# GCN-refinement positional encodings (PCALBSGarmentUseSegEncoderSeg, mesh_encoder.py:196-258, 450-466)

from .pointnet2 import pointnet2_utils as _pu   # noqa: E402


def _pe_reference_composition(group, mlp, xyz, new_xyz, features):
    """The reference's operator sequence for one unit (mesh_encoder.py:453-455 / 460-462), each operator on the B200 kernels:
    QueryAndGroup -> permute -> Linear -> ReLU -> Linear -> max over the samples."""
    B, P = new_xyz.shape[:2]
    qg = group(xyz=xyz, new_xyz=new_xyz, features=features)                 # (B, 3+C, P, K)
    qg = qg.reshape(B, qg.shape[1], P, group.nsample).permute(0, 2, 3, 1)
    return mlp(qg).max(-2)[0].reshape(B, P, -1)


class _PositionalEncodingFn(torch.autograd.Function):
    """Forward: ONE fused kernel (g4d_pe_mlp_max) after the ball query -- the (B, 3+C, P, K) grouped tensor and the two
    (B, P, K, 32) activations of the reference never exist.  Backward: the unit is RECOMPUTED through the reference's operator
    sequence with autograd on (activation checkpointing): gradients w.r.t. new_xyz (the garment vertices being refined),
    the two Linear layers and, when they require it, xyz / features are exactly those of the reference graph."""

    @staticmethod
    def forward(ctx, group, mlp, xyz, new_xyz, features, w1, b1, w2, b2):
        B, N, _ = xyz.shape
        P = new_xyz.shape[1]
        C = 0 if features is None else features.shape[1]
        xs, ns = xyz.detach().contiguous().float(), new_xyz.detach().contiguous().float()
        idx = _pu.ball_query(group.radius, group.nsample, xs, ns)
        feat_pm = None if features is None else features.detach().transpose(1, 2).contiguous().float()
        out = torch.empty(B, P, w2.shape[0], dtype=torch.float32, device=xyz.device)
        w1t, w2t = w1.detach().t().contiguous().float(), w2.detach().t().contiguous().float()
        rc = _lib.lib().g4d_pe_mlp_max(B, N, P, C, group.nsample, _lib.ptr(xs), _lib.ptr(ns), _lib.ptr(feat_pm), _lib.ptr(idx),
                                       _lib.ptr(w1t), _lib.ptr(b1.detach().contiguous().float()), _lib.ptr(w2t),
                                       _lib.ptr(b2.detach().contiguous().float()), _lib.ptr(out), None, _lib.stream_ptr())
        _lib.check(rc, "g4d_pe_mlp_max")
        ctx.group, ctx.mlp = group, mlp
        ctx.save_for_backward(xyz, new_xyz, features if features is not None else torch.empty(0))
        ctx.has_feat = features is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xyz, new_xyz, features = ctx.saved_tensors
        features = features if ctx.has_feat else None
        with torch.enable_grad():
            xs = xyz.detach().requires_grad_(ctx.needs_input_grad[2])
            ns = new_xyz.detach().requires_grad_(ctx.needs_input_grad[3])
            fs = None if features is None else features.detach().requires_grad_(ctx.needs_input_grad[4])
            out = _pe_reference_composition(ctx.group, ctx.mlp, xs, ns, fs)
            params = list(ctx.mlp.parameters())               # Linear1.weight, Linear1.bias, Linear2.weight, Linear2.bias
            wanted = [t for t, need in ((xs, ctx.needs_input_grad[2]), (ns, ctx.needs_input_grad[3]), (fs, ctx.needs_input_grad[4])) if need and t is not None]
            wanted_p = [p for p, need in zip(params, ctx.needs_input_grad[5:9]) if need]
            grads = torch.autograd.grad(out, wanted + wanted_p, grad_out, allow_unused=True)
        it = iter(grads)
        g_xyz = next(it) if ctx.needs_input_grad[2] else None
        g_new = next(it) if ctx.needs_input_grad[3] else None
        g_feat = next(it) if (ctx.needs_input_grad[4] and features is not None) else None
        g_params = [next(it) if need else None for need in ctx.needs_input_grad[5:9]]
        return (None, None, g_xyz, g_new, g_feat, *g_params)


class PositionalEncoding(nn.Module):
    """One (QueryAndGroup, Linear -> ReLU -> Linear, max over samples) unit of the GCN refinement; sub-module names `group` and
    `mlp` (an nn.Sequential like the reference's body_positional_encoding{i} / garment_positional_encoding{i}, so its state dict
    loads into `mlp`).  forward(xyz (B,N,3), new_xyz (B,P,3), features (B,C,N) or None) -> (B, P, feat_out)."""

    def __init__(self, radius, nsample, in_dim, feat_num=32, feat_out=32):
        super().__init__()
        self.group = _pu.QueryAndGroup(radius=radius, nsample=nsample, use_xyz=True)
        self.mlp = nn.Sequential(nn.Linear(in_dim, feat_num), nn.ReLU(), nn.Linear(feat_num, feat_out))
        self.fused = True

    def forward(self, xyz, new_xyz, features=None):
        l1, l2 = self.mlp[0], self.mlp[2]
        ok = (self.fused and xyz.is_cuda and self.group.nsample in (4, 8, 16, 32) and l1.out_features == 32 and l2.in_features == 32
              and l2.out_features == 32 and xyz.dtype == torch.float32)
        if not ok:
            return _pe_reference_composition(self.group, self.mlp, xyz, new_xyz, features)
        return _PositionalEncodingFn.apply(self.group, self.mlp, xyz, new_xyz, features, l1.weight, l1.bias, l2.weight, l2.bias)
