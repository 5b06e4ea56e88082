"""Callers of the hot path inside the garment model, on the B200 kernels, with the reference's
function signatures (modules/mesh_encoder.py).

  calc_segmentation_results   PCAGarmentEncoderSeg.calc_segmentation_results (mesh_encoder.py:109-125): garment point selection.
                              The reference loops over the B*T frames in Python with boolean-mask indexing (one device->host
                              sync per frame); here ONE kernel (g4d_select_points) does the arg-max over the class logits and an
                              order-preserving stream compaction for all frames.
  GarmentEncoderStack         the second set-abstraction stack on the selected garment points (mesh_encoder.py:54-78, 149-161):
                              two PointnetSAModuleMSG levels + the GroupAll summary module -- the fused tcgen05 route of
                              garment4d_b200.pointnet2 in eval mode.
  knn_points                  chamferdist.knn_points as the model calls it (mesh_encoder.py:321-324, 541): g4d_knn_points.
  smoothing_operator          normalize(adj_old) - I of mesh_encoder.py:386 as a CSR triple on the device (built once per mesh).
  lbs_garment_interpolation   MeshEncoder.lbs_garment_interpolation (mesh_encoder.py:312-410): one K-NN search instead of three,
                              the weighted gather of the skinning weights without the (F, body_v, K, 24) intermediate, the 100
                              smoothing steps and both skinning passes (g4d_lbs_skin) -- no torch math on the data path.
"""
from collections import namedtuple
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


# ---- lbs_garment_interpolation (mesh_encoder.py:312-410) ------------------------------------------------------------------------
KNN = namedtuple("KNN", ["dists", "idx"])      # the fields of chamferdist's / pytorch3d's knn_points result the model reads


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


def knn_points(p1, p2, K=1):
    """chamferdist.knn_points(p1, p2, K) as called at mesh_encoder.py:321-324: p1 (B,N,3), p2 (B,P,3) ->
    KNN(dists (B,N,K) squared distances ascending, idx (B,N,K) int64).  Equal distances come in ascending index order.
    K <= min(256, P), P <= 8192."""
    if not (p1.is_cuda and p2.is_cuda):
        raise _lib.G4DError("knn_points: CUDA tensors required (there is no CPU path)")
    a, b = _f32c(p1), _f32c(p2)
    B, N, _ = a.shape
    P = b.shape[1]
    d = torch.empty(B, N, K, dtype=torch.float32, device=a.device)
    i = torch.empty(B, N, K, dtype=torch.int32, device=a.device)
    _lib.check(_lib.lib().g4d_knn_points(B, N, P, K, _lib.ptr(a), _lib.ptr(b), _lib.ptr(d), _lib.ptr(i), _lib.stream_ptr()), "g4d_knn_points")
    return KNN(d, i.long())


def smoothing_operator(adj, device):
    """normalize(adj_old) - I (mesh_encoder.py:386; pygcn/utils.py:56-63 row normalisation) as (rowptr, col, val) int32/int32/fp32
    device tensors.  adj: the symmetric 0/1 garment-mesh adjacency as a dense (G,G) tensor, a torch sparse tensor, or anything
    with .tocoo() (scipy)."""
    if hasattr(adj, "tocoo"):
        coo = adj.tocoo()
        row, col = torch.as_tensor(coo.row, dtype=torch.int64), torch.as_tensor(coo.col, dtype=torch.int64)
        val, G = torch.as_tensor(coo.data, dtype=torch.float32), int(coo.shape[0])
    elif adj.layout == torch.strided:
        A = adj.detach().cpu().to(torch.float32)
        row, col = torch.nonzero(A, as_tuple=True)
        val, G = A[row, col], int(A.shape[0])
    else:
        A = adj.detach().cpu().coalesce()
        (row, col), val, G = A.indices(), A.values().to(torch.float32), int(A.shape[0])
    rowsum = torch.zeros(G, dtype=torch.float32).index_add_(0, row, val)
    r_inv = torch.where(rowsum != 0, 1.0 / rowsum, torch.zeros_like(rowsum))      # r_inv[isinf] = 0
    diag = torch.arange(G, dtype=torch.int64)
    key = torch.cat([row, diag]) * G + torch.cat([col, diag])                      # entries of r_inv * A - I, duplicates summed
    v = torch.cat([val * r_inv[row], -torch.ones(G)])
    ukey, inv = torch.unique(key, sorted=True, return_inverse=True)
    uval = torch.zeros(ukey.numel(), dtype=torch.float32).index_add_(0, inv, v)
    urow, ucol = ukey // G, ukey % G
    rowptr = torch.zeros(G + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(urow, minlength=G), 0)
    return rowptr.to(torch.int32).to(device), ucol.to(torch.int32).to(device), uval.to(device)


def lbs_garment_interpolation(pred_template_garment_v, Tpose_vertices, Tpose_root_joints, zeropose_vertices, body_model, gt_pose,
                              T_J_regressor, T_lbs_weights, K=3, smooth=None, smooth_iters=100, coeff=0.1):
    """MeshEncoder.lbs_garment_interpolation (mesh_encoder.py:312-410), same arguments and results:
        pred_template_garment_v (B,G,3), Tpose_vertices (B,P,3) (or (B,1,P,3)), Tpose_root_joints (B,3), zeropose_vertices (B,T,P,3),
        body_model: anything with .parents, gt_pose (B,T,72), T_J_regressor (B,T,J,P), T_lbs_weights (B,T,P,J)
        -> lbs_pred_garment_v (B,T,G,3), nn (KNN with K = 1), stage-1 (inverse-posed) garment (B,T,G,3)
    smooth: the (rowptr, col, val) triple of smoothing_operator() -- the reference rebuilds normalize(self.adj_old) - I on the host
    on every call (:386-387); required when K > 1."""
    from . import lbs as L
    dev = pred_template_garment_v.device
    if dev.type != "cuda":
        raise _lib.G4DError("lbs_garment_interpolation: CUDA tensors required (there is no CPU path)")
    assert pred_template_garment_v.dim() == 3 and pred_template_garment_v.shape[2] == 3
    assert gt_pose.dim() == 3 and gt_pose.shape[2] == 72
    lib = _lib.lib()
    st = _lib.stream_ptr()
    B, G, _ = pred_template_garment_v.shape
    T = gt_pose.shape[1]
    J = T_J_regressor.shape[2]
    parents = body_model.parents if hasattr(body_model, "parents") else body_model
    gt_pose_mat = L.batch_rodrigues(_f32c(gt_pose).reshape(-1, 3)).reshape(B * T, 24, 3, 3)                  # :318
    q = (_f32c(pred_template_garment_v) + _f32c(Tpose_root_joints).reshape(B, 1, 3)).contiguous()            # :320
    body = _f32c(Tpose_vertices).reshape(B, -1, 3)
    P = body.shape[1]
    # :321-324 -- one search: the K = min(64, K) and K = 1 results are prefixes of the K one
    dK = torch.empty(B, G, K, dtype=torch.float32, device=dev)
    iK = torch.empty(B, G, K, dtype=torch.int32, device=dev)
    _lib.check(lib.g4d_knn_points(B, G, P, K, _lib.ptr(q), _lib.ptr(body), _lib.ptr(dK), _lib.ptr(iK), st), "g4d_knn_points")
    K64 = min(64, K)
    nn = KNN(dK[:, :, :1].contiguous(), iK[:, :, :1].long())
    # :326-335 -- the fixed inverse template pose
    inv_pose = torch.zeros(B, 24, 3, dtype=torch.float32, device=dev)
    inv_pose[:, 0, 0] = -3.141592653589793 / 2
    inv_pose[:, 1, 1] = 0.15
    inv_pose[:, 2, 1] = -0.15
    inv_mat = L.batch_rodrigues(inv_pose.reshape(-1, 3)).reshape(B, 24, 3, 3)
    Jreg = _f32c(T_J_regressor)
    inv_J = L.vertices2jointsB(Jreg[:, 0].contiguous(), body)
    _, inv_A = L.batch_rigid_transform(inv_mat, inv_J, parents)
    # :339-347 -- skinning weights of the garment in the T pose: inverse-distance blend of the K64 nearest body vertices
    Wall = _f32c(T_lbs_weights)                                                                               # (B,T,P,J)
    w64 = torch.empty(B, G, K64, dtype=torch.float32, device=dev)
    _lib.check(lib.g4d_knn_inverse_weights(B * G, K, K64, _lib.ptr(dK), _lib.ptr(w64), st), "g4d_knn_inverse_weights")
    W0 = Wall[:, 0].contiguous()
    inv_nn_W = torch.empty(B, G, J, dtype=torch.float32, device=dev)
    _lib.check(lib.g4d_knn_blend_weights(B, 1, G, P, J, K64, K, _lib.ptr(iK), _lib.ptr(w64), _lib.ptr(W0), _lib.ptr(inv_nn_W), st),
               "g4d_knn_blend_weights")
    stage1_b = L.skin(q, inv_A, inv_nn_W)                                                                     # :357-362  (B,G,3)
    stage1 = stage1_b.reshape(B, 1, G, 3).repeat(1, T, 1, 1).reshape(B * T, G, 3).contiguous()
    # :364-369 -- the posed skeleton of every frame
    Jz = L.vertices2jointsB(Jreg.reshape(B * T, J, -1), _f32c(zeropose_vertices).reshape(B * T, -1, 3))
    _, A = L.batch_rigid_transform(gt_pose_mat, Jz, parents)
    # :371-379 -- per-frame skinning weights of the garment
    wK = torch.empty(B, G, K, dtype=torch.float32, device=dev)
    _lib.check(lib.g4d_knn_inverse_weights(B * G, K, K, _lib.ptr(dK), _lib.ptr(wK), st), "g4d_knn_inverse_weights")
    nn_W = torch.empty(B * T, G, J, dtype=torch.float32, device=dev)
    Wf = Wall.reshape(B * T, P, J)
    _lib.check(lib.g4d_knn_blend_weights(B, T, G, P, J, K, K, _lib.ptr(iK), _lib.ptr(wK), _lib.ptr(Wf), _lib.ptr(nn_W), st),
               "g4d_knn_blend_weights")
    if K > 1:                                                                                                 # :382-389
        if smooth is None:
            raise _lib.G4DError("lbs_garment_interpolation: K > 1 needs smooth = smoothing_operator(adj_old, device)")
        rowptr, col, val = smooth
        if rowptr.numel() != G + 1:
            raise _lib.G4DError("lbs_garment_interpolation: the smoothing operator does not match the garment's vertex count")
        tmp = torch.empty_like(nn_W)
        _lib.check(lib.g4d_smooth_weights(B * T, G, J, int(smooth_iters), float(coeff), _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val),
                                          _lib.ptr(nn_W), _lib.ptr(tmp), st), "g4d_smooth_weights")
    out = L.skin(stage1, A, nn_W)                                                                             # :391-408
    return out.reshape(B, T, G, 3), nn, stage1.reshape(B, T, G, 3)
