"""TEST INFRASTRUCTURE ONLY -- CPU port of the reference's encoder forward
(modules/pointnet2encoder.py:112-145 driving pointnet2_modules.py:19-55,131-156), used as the checker for the
fused GPU route and as the timed ``cpu_baseline`` / ``--impl reference`` arm of bench.py.

The point operators are the C oracle (oracle/pointnet2_oracle.c, OpenMP over clouds); the 1x1-conv / BatchNorm /
ReLU stacks, pooling and concatenations run under torch on CPU in fp32, layer by layer, exactly in the order the
reference's modules apply them.  `model` is a CPU copy of a garment4d_b200.encoder.Pointnet2MSGSEG (only its
parameters and sub-module containers are used; none of its forward code runs here).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import pointnet2 as orc


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def sa_module_cpu(sa, xyz, features):
    """_PointnetSAModuleBase.forward (pointnet2_modules.py:19-55) -- xyz (B,N,3) np, features (B,C,N) np or None."""
    idx = orc.furthest_point_sample(xyz, sa.npoint, fast=True)
    new_xyz = np.ascontiguousarray(orc.gather_operation(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx).transpose(0, 2, 1))
    outs = []
    for grouper, mlp in zip(sa.groupers, sa.mlps):
        grouped = orc.query_and_group(grouper.radius, grouper.nsample, xyz, new_xyz, features, use_xyz=grouper.use_xyz)
        y = mlp(_t(grouped))
        y = F.max_pool2d(y, kernel_size=[1, y.size(3)]).squeeze(-1)
        outs.append(y)
    return new_xyz, torch.cat(outs, dim=1).numpy()


def fp_module_cpu(fp, unknown, known, unknow_feats, known_feats):
    """PointnetFPModule.forward (pointnet2_modules.py:131-156)."""
    dist, idx = orc.three_nn(unknown, known)
    dist_recip = (np.float32(1.0) / (dist + np.float32(1e-8))).astype(np.float32)
    weight = (dist_recip / dist_recip.sum(axis=2, keepdims=True)).astype(np.float32)
    interp = orc.three_interpolate(known_feats, idx, weight)
    new = interp if unknow_feats is None else np.concatenate([interp, unknow_feats], axis=1)
    return fp.mlp(_t(new).unsqueeze(-1)).squeeze(-1).numpy()


@torch.no_grad()
def encoder_forward_cpu(model, pointcloud, sa_only=False):
    """pointcloud (B,N,3) float32 numpy.  Returns (sem_logits (B,N,classes) or None, l_features, l_xyz)."""
    model.eval()
    xyz = np.ascontiguousarray(pointcloud[..., :3], dtype=np.float32)
    l_xyz, l_features = [xyz], [None]
    for sa in model.SA_modules:
        nx, nf = sa_module_cpu(sa, l_xyz[-1], l_features[-1])
        l_xyz.append(nx)
        l_features.append(nf)
    if sa_only:
        return None, l_features, l_xyz
    for i in range(-1, -(len(model.FP_modules) + 1), -1):
        l_features[i - 1] = fp_module_cpu(model.FP_modules[i], l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i])
    sem = model.FC_layer(_t(l_features[0])).transpose(1, 2).contiguous().numpy()
    return sem, l_features, l_xyz
