#!/usr/bin/env python
"""FPS level-0 timing (8192 -> 1024) for several cloud counts and kernel variants (env G4D_FPS / G4D_FPS_WS / G4D_FPS_WIDE)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from garment4d_b200.pointnet2 import pointnet2_utils as pu
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for kind in ("body", "cube"):
    for C in (30, 60, 120, 148, 240):
        x = torch.from_numpy(bench.make_inputs(kind, 77, C, 8192)).to(dev)
        pu.furthest_point_sample_and_gather(x, 1024); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(5):
            flush.fill_(0.0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); pu.furthest_point_sample_and_gather(x, 1024); e.record(); torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        print(f"{kind} C={C}: {tot / 5:.3f} ms")
