#!/usr/bin/env python
"""Where a serial FPS step spends its cycles: clock() phase sums per warp of cloud 0 of the pruned kernel (measurement build,
G4D_FPS_PROF=1: one CTA per SM).   G4D_FPS_PROF=1 python tools/fps_phases.py [clouds]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("G4D_FPS_PROF", "1")
import torch
import bench
from garment4d_b200 import _lib
from garment4d_b200.pointnet2 import pointnet2_utils as pu
C = int(sys.argv[1]) if len(sys.argv) > 1 else 120
dev = torch.device("cuda:0")
names = ["test", "update", "warp argmax", "slot write", "barrier", "block argmax", "latch+thr"]
for kind in ("body", "cube"):
    x = torch.from_numpy(bench.make_inputs(kind, 77, C, 8192)).to(dev)
    pu.furthest_point_sample_and_gather(x, 1024); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); pu.furthest_point_sample_and_gather(x, 1024); e.record(); torch.cuda.synchronize()
    buf = (ctypes.c_uint * 640)()
    _lib.lib().g4d_debug_fps_phases(ctypes.cast(buf, ctypes.c_void_p))
    a = np.array(buf, dtype=np.int64).reshape(32, 20)[:16]
    print(f"== {kind}, {C} clouds: {s.elapsed_time(e):.3f} ms; cycles per step, mean over the 16 warps of cloud 0")
    for f, lab in ((1, "steps where the warp updates"), (0, "steps where it does not")):
        n = a[:, 16 + f].astype(np.float64)
        print(f"  {lab}: {n.mean():.0f} of 1022 steps per warp")
        tot = 0.0
        for i, nm in enumerate(names):
            v = (a[:, f * 8 + i] / np.maximum(n, 1)).mean(); tot += v
            print(f"      {nm:14s} {v:7.0f}")
        print(f"      {'sum':14s} {tot:7.0f}")
