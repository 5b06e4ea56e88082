#!/usr/bin/env python
"""One pass of every hot-path kernel at BASELINE config c3 (240 clouds x 8192 points, one frame group, no graph) inside
a cudaProfilerStart/Stop range -- the target of the `ncu --set full --profile-from-start off` capture in
tools/profile_round.sh.  Also runs the stand-alone fused ball-query+group operator (QueryAndGroup) per branch, which the
fused SA route never launches."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from garment4d_b200 import _lib, synthetic
from garment4d_b200 import lbs as glbs
from garment4d_b200.encoder import Pointnet2MSGSEG
from garment4d_b200.pointnet2 import pointnet2_utils as pu

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
B, T, N = bench.CONFIGS[cfg]
C = B * T
dev = torch.device("cuda:0")
_lib.lib()
torch.manual_seed(1234)
model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).eval()
smpl_np = synthetic.synthetic_smpl(seed=1234)
smpl = [torch.from_numpy(np.ascontiguousarray(smpl_np[k])).to(dev) for k in
        ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights")]
base = synthetic.body_clouds(1234 + 3000, min(C, 16), N)
pc = torch.from_numpy(np.tile(base, ((C + 15) // 16, 1, 1))[:C].copy()).to(dev)
b, p = synthetic.synthetic_frames(C, seed=2)
betas, pose = torch.from_numpy(b).to(dev), torch.from_numpy(p).to(dev)


def one_pass(profile):
    with torch.no_grad():
        if profile:
            torch.cuda.profiler.start()
        model(pc)
        glbs.lbs(betas, pose, *smpl)
        torch.cuda.synchronize()
        if profile:
            torch.cuda.profiler.stop()
        l_xyz, l_feat = model.sa_stack(pc)          # SA-level features (the FP stack overwrites l_features[1..2])
        torch.cuda.synchronize()
        if profile:
            torch.cuda.profiler.start()
        for lvl, sa in enumerate(model.SA_modules):
            for g in sa.groupers:
                pu.QueryAndGroup(g.radius, g.nsample)(l_xyz[lvl], l_xyz[lvl + 1], l_feat[lvl])
        torch.cuda.synchronize()
        if profile:
            torch.cuda.profiler.stop()


one_pass(False)                 # parameter caches, lazy kernel attributes
one_pass(True)
print("ncu_once: done")
