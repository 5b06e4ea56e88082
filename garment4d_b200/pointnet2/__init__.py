"""Mirror of the reference package modules/pointnet2/pointnet2 (pointnet2_utils, pointnet2_modules, pytorch_utils)."""
from . import pointnet2_utils, pointnet2_modules, pytorch_utils  # noqa: F401
