"""Top-level drop-in for the reference's compiled module of the same name
(``import pointnet2_cuda as pointnet2``, modules/pointnet2/pointnet2/pointnet2_utils.py:7).
With the repository root on sys.path the reference's Python layer binds to the B200 kernels unchanged."""
from garment4d_b200.pointnet2_cuda import *  # noqa: F401,F403
from garment4d_b200.pointnet2_cuda import __all__  # noqa: F401
