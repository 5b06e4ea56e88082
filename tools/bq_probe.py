#!/usr/bin/env python
"""Level-0 two-radius grid ball query at 240 clouds (an ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from garment4d_b200.pointnet2 import pointnet2_utils as pu
dev = torch.device("cuda:0")
x = torch.from_numpy(bench.make_inputs("body", 77, 240, 8192)).to(dev)
_, nx = pu.furthest_point_sample_and_gather(x, 1024)
for _ in range(3):
    pu.ball_query_pair(0.05, 16, 0.1, 32, x, nx)
torch.cuda.synchronize()
