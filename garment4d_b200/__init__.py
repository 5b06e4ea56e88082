"""garment4d_b200 -- B200 (sm_100a) implementation of Garment4D's data-parallel hot path:
the PointNet++ set-abstraction stack and SMPL linear-blend skinning, behind the reference's
own Python operator API.

    garment4d_b200.pointnet2_cuda            the 9 functions of the reference's compiled extension
    garment4d_b200.pointnet2.pointnet2_utils furthest_point_sample, gather_operation, ball_query, ...
    garment4d_b200.pointnet2.pointnet2_modules  PointnetSAModuleMSG / PointnetSAModule / PointnetFPModule
    garment4d_b200.pointnet2.pytorch_utils   SharedMLP, Conv1d, Conv2d, ...
    garment4d_b200.lbs                       lbs, batch_rodrigues, batch_rigid_transform, ...
    garment4d_b200.encoder                   Pointnet2MSGSEG (modules/pointnet2encoder.py)

All compute runs in libgarment4d_b200.so (hand-written CUDA, C ABI in include/garment4d_b200.h).
There is no CPU fallback; missing library = error.
"""
from ._lib import G4DError, SO_PATH, lib  # noqa: F401

__version__ = "0.1.0"
