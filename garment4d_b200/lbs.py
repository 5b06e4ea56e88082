"""SMPL linear-blend skinning on the B200 kernels, with the reference's function signatures
(smplx/smplx/lbs.py): ``lbs``, ``blend_shapes``, ``vertices2joints``, ``vertices2jointsB``, ``batch_rodrigues``,
``transform_mat``, ``batch_rigid_transform``.

The forward pass of every function runs in libgarment4d_b200.so (csrc/lbs.cu) in fp32; tensors must be CUDA
float32.  CPU tensors raise: there is no CPU path in this package (the reference runs lbs() on CPU inside its
DataLoader workers, utils/dataloader.py:187-212 -- move those tensors to the GPU and batch the frames instead).
Inputs that require grad are served by a differentiable composition of torch CUDA ops with the same math, because
the reference's lbs() is differentiable; the kernels are the forward/inference path.
"""
from typing import Tuple

import torch
import torch.nn.functional as F

from . import _lib

Tensor = torch.Tensor


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.G4DError("garment4d_b200.lbs runs on CUDA tensors only (got a CPU tensor); there is no CPU fallback")


def _wants_grad(*ts):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def _c(t):
    return t.detach().contiguous().float()


_PARENTS_CACHE = {}


def _parents_i32(parents: Tensor, device):
    key = (parents.data_ptr(), parents._version, str(device), parents.numel())
    p = _PARENTS_CACHE.get(key)
    if p is None:
        if len(_PARENTS_CACHE) > 64:
            _PARENTS_CACHE.clear()
        p = parents.detach().to(device=device, dtype=torch.int32).contiguous()
        _PARENTS_CACHE[key] = p
    return p


def batch_rodrigues(rot_vecs: Tensor, epsilon: float = 1e-8) -> Tensor:
    """(N,3) axis-angle -> (N,3,3) rotation matrices (lbs.py:312-346)."""
    _need_cuda(rot_vecs)
    if _wants_grad(rot_vecs):
        return _batch_rodrigues_torch(rot_vecs)
    v = _c(rot_vecs)
    n = v.shape[0]
    out = torch.empty(n, 3, 3, dtype=torch.float32, device=v.device)
    _lib.check(_lib.lib().g4d_batch_rodrigues(n, _lib.ptr(v), _lib.ptr(out), _lib.stream_ptr()), "g4d_batch_rodrigues")
    return out


def vertices2joints(J_regressor: Tensor, vertices: Tensor) -> Tensor:
    """J_regressor (J,V), vertices (B,V,3) -> (B,J,3) (lbs.py:251-268)."""
    _need_cuda(J_regressor, vertices)
    if _wants_grad(J_regressor, vertices):
        return torch.einsum("bik,ji->bjk", [vertices, J_regressor])
    Jr, v = _c(J_regressor), _c(vertices)
    B, V, _ = v.shape
    J = Jr.shape[0]
    out = torch.empty(B, J, 3, dtype=torch.float32, device=v.device)
    _lib.check(_lib.lib().g4d_vertices2joints(B, V, J, 0, _lib.ptr(Jr), _lib.ptr(v), _lib.ptr(out), _lib.stream_ptr()),
               "g4d_vertices2joints")
    return out


def vertices2jointsB(J_regressor_B: Tensor, vertices: Tensor) -> Tensor:
    """Per-frame regressor (B,J,V), vertices (B,V,3) -> (B,J,3) (lbs.py:270-286)."""
    _need_cuda(J_regressor_B, vertices)
    if _wants_grad(J_regressor_B, vertices):
        return torch.einsum("bik,bji->bjk", [vertices, J_regressor_B])
    Jr, v = _c(J_regressor_B), _c(vertices)
    B, V, _ = v.shape
    J = Jr.shape[1]
    out = torch.empty(B, J, 3, dtype=torch.float32, device=v.device)
    _lib.check(_lib.lib().g4d_vertices2joints(B, V, J, 1, _lib.ptr(Jr), _lib.ptr(v), _lib.ptr(out), _lib.stream_ptr()),
               "g4d_vertices2joints")
    return out


def blend_shapes(betas: Tensor, shape_disps: Tensor) -> Tensor:
    """betas (B,NB), shape_disps (V,3,NB) -> (B,V,3) displacement (lbs.py:288-309)."""
    _need_cuda(betas, shape_disps)
    if _wants_grad(betas, shape_disps):
        return torch.einsum("bl,mkl->bmk", [betas, shape_disps])
    bt, sd = _c(betas), _c(shape_disps)
    B, NB = bt.shape
    V = sd.shape[0]
    out = torch.empty(B, V, 3, dtype=torch.float32, device=bt.device)
    _lib.check(_lib.lib().g4d_blend_shapes(B, V, NB, _lib.ptr(bt), _lib.ptr(sd), _lib.ptr(out), _lib.stream_ptr()), "g4d_blend_shapes")
    return out


def transform_mat(R: Tensor, t: Tensor) -> Tensor:
    """R (B,3,3), t (B,3,1) -> (B,4,4) (lbs.py:349-359)."""
    return torch.cat([F.pad(R, [0, 0, 0, 1]), F.pad(t, [0, 0, 0, 1], value=1)], dim=2)


def batch_rigid_transform(rot_mats: Tensor, joints: Tensor, parents: Tensor, dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """rot_mats (B,J,3,3), joints (B,J,3), parents (J) -> posed_joints (B,J,3), rel_transforms (B,J,4,4) (lbs.py:362-419)."""
    _need_cuda(rot_mats, joints)
    if _wants_grad(rot_mats, joints):
        return _batch_rigid_transform_torch(rot_mats, joints, parents)
    R, Jt = _c(rot_mats), _c(joints)
    B, J = Jt.shape[:2]
    posed = torch.empty(B, J, 3, dtype=torch.float32, device=R.device)
    A = torch.empty(B, J, 4, 4, dtype=torch.float32, device=R.device)
    rc = _lib.lib().g4d_batch_rigid_transform(B, J, _lib.ptr(R), _lib.ptr(Jt), _lib.ptr(_parents_i32(parents, R.device)),
                                              _lib.ptr(posed), _lib.ptr(A), _lib.stream_ptr())
    _lib.check(rc, "g4d_batch_rigid_transform")
    return posed, A


def skin(v_posed: Tensor, A: Tensor, lbs_weights: Tensor) -> Tensor:
    """The skinning tail of lbs() (lbs.py:233-246; same math at modules/mesh_encoder.py:347,362,393,408):
    verts = (W.A)[v_posed;1].  v_posed (B,V,3), A (B,J,4,4), lbs_weights (V,J) or per-frame (B,V,J)."""
    _need_cuda(v_posed, A, lbs_weights)
    vp, Ac, W = _c(v_posed), _c(A), _c(lbs_weights)
    B, V, _ = vp.shape
    J = Ac.shape[1]
    out = torch.empty(B, V, 3, dtype=torch.float32, device=vp.device)
    rc = _lib.lib().g4d_lbs_skin(B, V, J, int(W.dim() == 3), _lib.ptr(vp), _lib.ptr(Ac), _lib.ptr(W), _lib.ptr(out),
                                 _lib.stream_ptr())
    _lib.check(rc, "g4d_lbs_skin")
    return out


def lbs(betas: Tensor, pose: Tensor, v_template: Tensor, shapedirs: Tensor, posedirs: Tensor, J_regressor: Tensor,
        parents: Tensor, lbs_weights: Tensor, pose2rot: bool = True) -> Tuple[Tensor, Tensor]:
    """Linear blend skinning (lbs.py:152-248).

    betas (B,NB); pose (B,(J)*3) axis-angle when pose2rot else rotation matrices (B,J,3,3) / (B,J*9);
    v_template (V,3); shapedirs (V,3,NB); posedirs ((J-1)*9, V*3); J_regressor (J,V); parents (J); lbs_weights (V,J).
    Returns verts (B,V,3), joints (B,J,3).
    """
    _need_cuda(betas, pose, v_template, shapedirs, posedirs, J_regressor, lbs_weights)
    if _wants_grad(betas, pose, v_template, shapedirs, posedirs, J_regressor, lbs_weights):
        return _lbs_torch(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot)
    if betas.dtype != torch.float32:
        raise _lib.G4DError("garment4d_b200.lbs computes in float32 (the reference's training dtype)")
    Fb = max(betas.shape[0], pose.shape[0])
    V, J = v_template.shape[-2], J_regressor.shape[0]
    if v_template.dim() == 3:
        if v_template.shape[0] != 1:
            raise _lib.G4DError("per-frame v_template is not supported by the fused lbs kernel")
        v_template = v_template[0]
    NB = betas.shape[1]
    b, p = _c(betas), _c(pose)
    if p.shape[0] != Fb:
        p = p.expand(Fb, *p.shape[1:]).contiguous()
    vt, sd, pd, Jr, W = _c(v_template), _c(shapedirs), _c(posedirs), _c(J_regressor), _c(lbs_weights)
    assert sd.shape == (V, 3, NB) and pd.shape == ((J - 1) * 9, V * 3) and W.shape == (V, J)
    L = _lib.lib()
    nbytes = L.g4d_lbs_workspace_bytes(Fb, V, J)
    ws = torch.empty(nbytes // 4, dtype=torch.float32, device=b.device)
    verts = torch.empty(Fb, V, 3, dtype=torch.float32, device=b.device)
    joints = torch.empty(Fb, J, 3, dtype=torch.float32, device=b.device)
    rc = L.g4d_lbs(Fb, V, J, NB, b.shape[0], int(bool(pose2rot)), _lib.ptr(b), _lib.ptr(p), _lib.ptr(vt), _lib.ptr(sd),
                   _lib.ptr(pd), _lib.ptr(Jr), _lib.ptr(_parents_i32(parents, b.device)), _lib.ptr(W), _lib.ptr(verts),
                   _lib.ptr(joints), _lib.ptr(ws), nbytes, _lib.stream_ptr())
    _lib.check(rc, "g4d_lbs")
    return verts, joints


# ---- differentiable compositions (autograd only; same math, torch CUDA ops) -----------------------------

def _batch_rodrigues_torch(rot_vecs):
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    d = rot_vecs / angle
    cos, sin = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    rx, ry, rz = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    z = torch.zeros_like(rx)
    K = torch.cat([z, -rz, ry, rz, z, -rx, -ry, rx, z], dim=1).view(n, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
    return eye + sin * K + (1 - cos) * torch.bmm(K, K)


def _batch_rigid_transform_torch(rot_mats, joints, parents):
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] = rel[:, 1:] - joints[:, parents[1:]]
    tm = transform_mat(rot_mats.reshape(-1, 3, 3), rel.reshape(-1, 3, 1)).reshape(-1, joints.shape[1], 4, 4)
    chain = [tm[:, 0]]
    for i in range(1, parents.shape[0]):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed = transforms[:, :, :3, 3]
    jh = F.pad(joints, [0, 0, 0, 1])
    rel_t = transforms - F.pad(torch.matmul(transforms, jh), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed, rel_t


def _lbs_torch(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot):
    B = max(betas.shape[0], pose.shape[0])
    v_shaped = v_template + torch.einsum("bl,mkl->bmk", [betas, shapedirs])
    J = torch.einsum("bik,ji->bjk", [v_shaped, J_regressor])
    eye = torch.eye(3, dtype=betas.dtype, device=betas.device)
    rot = _batch_rodrigues_torch(pose.view(-1, 3)).view(B, -1, 3, 3) if pose2rot else pose.view(B, -1, 3, 3)
    pf = (rot[:, 1:] - eye).view(B, -1)
    v_posed = torch.matmul(pf, posedirs).view(B, -1, 3) + v_shaped
    Jt, A = _batch_rigid_transform_torch(rot, J, parents)
    W = lbs_weights.unsqueeze(0).expand(B, -1, -1)
    T = torch.matmul(W, A.view(B, J_regressor.shape[0], 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=betas.dtype, device=betas.device)], dim=2)
    return torch.matmul(T, vh.unsqueeze(-1))[:, :, :3, 0], Jt
