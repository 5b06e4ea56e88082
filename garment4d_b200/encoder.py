"""``Pointnet2MSGSEG`` -- the point-cloud encoder of Garment4D (modules/pointnet2encoder.py:18-145) on the B200
set-abstraction stack.  Same constructor, same sub-module names (``SA_modules``, ``FP_modules``, ``FC_layer``,
``Middle_modules``) and therefore the same state-dict keys, same ``forward`` return tuple
``(middle_features, sem_logits, l_features, l_xyz)``.

The reference file cannot be imported on its own (it pulls ``utils.config`` -- argparse at import time -- and
``utils.dataloader``), so the 7 segmentation classes (utils/dataloader.py:24) are a constructor default here.
Garment4D builds it as ``Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False)`` (modules/mesh_encoder.py:49).
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from .pointnet2 import pointnet2_cuda_bridge as _bridge
from .pointnet2 import pytorch_utils as pt_utils
from .pointnet2.pointnet2_modules import PointnetFPModule, PointnetSAModule, PointnetSAModuleMSG

CLASS_NUM = 7   # utils/dataloader.py:24


class Pointnet2MSGSEG(nn.Module):
    def __init__(self, input_channels=3, use_xyz=True, bn=True, global_feat=True, class_num=CLASS_NUM):
        super().__init__()
        self.global_feat = global_feat
        c0 = input_channels
        self.SA_modules = nn.ModuleList()
        # six branches, pointnet2encoder.py:41-76
        self.SA_modules.append(PointnetSAModuleMSG(npoint=1024, radii=[0.05, 0.1], nsamples=[16, 32],
                                                   mlps=[[c0, 16, 16, 32], [c0, 32, 32, 64]], use_xyz=use_xyz, bn=bn))
        c1 = 32 + 64
        self.SA_modules.append(PointnetSAModuleMSG(npoint=256, radii=[0.1, 0.2], nsamples=[16, 32],
                                                   mlps=[[c1, 32, 32, 64], [c1, 64, 64, 128]], use_xyz=use_xyz, bn=bn))
        c2 = 64 + 128
        self.SA_modules.append(PointnetSAModuleMSG(npoint=64, radii=[0.2, 0.4], nsamples=[32, 64],
                                                   mlps=[[c2, 64, 64, 128], [c2, 128, 128, 256]], use_xyz=use_xyz, bn=bn))
        c3 = 128 + 256
        if global_feat:
            self.Middle_modules = PointnetSAModule(mlp=[c3, 256, 512], use_xyz=use_xyz, bn=bn)   # pointnet2encoder.py:80-84
        self.num_feat = 512
        self.pointwise_num_feat = 64 + 128 + 256 + 128 + 256
        self.feat_channels_list = [64, 128, 256, 128 + 256]
        self.FP_modules = nn.ModuleList()                                                       # pointnet2encoder.py:91-96
        self.FP_modules.append(PointnetFPModule(mlp=[128 + c0, 128, 64], bn=bn))
        self.FP_modules.append(PointnetFPModule(mlp=[256 + c1, 256, 128], bn=bn))
        self.FP_modules.append(PointnetFPModule(mlp=[c3 + c2, 512, 256], bn=bn))
        self.FP_modules[1].emit_point_major = True     # its output feeds the fused finest-level kernel (fp16 point-major gather)
        self.FP_modules[2].emit_point_major = True     # ... and this one feeds FP_modules[1]'s interpolation the same way
        self.FC_layer = nn.Sequential(pt_utils.Conv1d(64, 32, bn=True), nn.Dropout(),
                                      pt_utils.Conv1d(32, class_num, activation=None))           # pointnet2encoder.py:98-101

    def _break_up_pc(self, pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def sa_stack(self, pointcloud: torch.Tensor):
        """Only the three set-abstraction levels (the FPS / ball-query / grouped-MLP hot path):
        returns (l_xyz, l_features) lists of length 4."""
        xyz, features = self._break_up_pc(pointcloud)
        l_xyz, l_features = [xyz], [features]
        for sa in self.SA_modules:
            li_xyz, li_features = sa(l_xyz[-1], l_features[-1])
            l_xyz.append(li_xyz)
            l_features.append(li_features)
        return l_xyz, l_features

    fused = True      # set False to force the module-by-module route for the finest FP level + head

    def _fused_fp0_head(self, l_xyz, l_features):
        """Finest feature-propagation level + FC head in one tcgen05 kernel (eval mode, no autograd, no skip features):
        three_nn -> g4d_fp_interp_mlp.  Returns (l_features[0], sem_logits) or None when the route does not apply."""
        fp = self.FP_modules[0]
        if (not self.fused or self.training or l_features[0] is not None or not l_xyz[0].is_cuda
                or (torch.is_grad_enabled() and (l_features[1].requires_grad or any(p.requires_grad for p in self.parameters())))):
            return None
        ver = pt_utils.shared_mlp_version(fp.mlp) + pt_utils.shared_mlp_version(self.FC_layer)
        key = str(l_xyz[0].device)
        hit = getattr(self, "_fp0_cache", {}).get(key)
        if hit is None or hit[0] != ver:
            packed = _bridge.pack_fp_head(fp.mlp, self.FC_layer, l_xyz[0].device)
            self._fp0_cache = {key: (ver, packed)}
            hit = self._fp0_cache[key]
        if hit[1] is None:
            return None
        return _bridge.fp_interp_mlp(hit[1], l_xyz[0], l_xyz[1], l_features[1])

    def forward(self, pointcloud: torch.Tensor):
        """pointcloud (B, N, 3 + input_channels) -> (middle_features | None, sem_logits (B,N,class_num),
        l_features [4], l_xyz [4])  (pointnet2encoder.py:112-145)"""
        l_xyz, l_features = self.sa_stack(pointcloud)
        middle_features = self.Middle_modules(l_xyz[-1], l_features[-1])[1] if self.global_feat else None
        nfp = len(self.FP_modules)
        for i in range(-1, -nfp, -1):
            l_features[i - 1] = self.FP_modules[i](l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i])
        fused = self._fused_fp0_head(l_xyz, l_features)
        if fused is not None:
            l_features[0], sem_logits = fused
        else:
            l_features[0] = self.FP_modules[0](l_xyz[0], l_xyz[1], l_features[0], l_features[1])
            sem_logits = self.FC_layer(l_features[0]).transpose(1, 2).contiguous()
        return middle_features, sem_logits, l_features, l_xyz
