#!/usr/bin/env python
"""three_nn level 0 (8192 unknown, 1024 known) against the grid density factor.  python tools/nn_sweep.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from garment4d_b200.pointnet2 import pointnet2_utils as pu
dev = torch.device("cuda:0")
C = 240
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for kind in ("body", "cube"):
    x = torch.from_numpy(bench.make_inputs(kind, 77, C, 8192)).to(dev)
    _, known = pu.furthest_point_sample_and_gather(x, 1024)
    pu.build_grid(x, 0.2)
    d = torch.empty(C, 8192, 3, device=dev); i = torch.empty(C, 8192, 3, dtype=torch.int32, device=dev)
    for f in (0.7, 0.9, 1.0, 1.1, 1.15, 1.2, 1.25, 1.3, 1.4):
        pu.NN_CELLS_FACTOR = f
        pu.three_nn_raw(x, known, d, i); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(5):
            flush.fill_(0.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); pu.three_nn_raw(x, known, d, i); e.record(); torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        print(f"{kind} factor {f}: {tot / 5:.4f} ms (grid build included)")
