#!/usr/bin/env python
"""Per-kernel timing table of the hot path (same measurement as bench.py's "kernels" list, without the rest)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from garment4d_b200 import _lib, synthetic
from garment4d_b200.encoder import Pointnet2MSGSEG

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
B, T, N = bench.CONFIGS[cfg]
C = B * T
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).eval()
smpl_np = synthetic.synthetic_smpl(seed=1234)
smpl = [torch.from_numpy(np.ascontiguousarray(smpl_np[k])).to(dev) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights")]
base = synthetic.body_clouds(1, min(C, 16), N)
pc = torch.from_numpy(np.tile(base, ((C + 15) // 16, 1, 1))[:C].copy()).to(dev)
b, p = synthetic.synthetic_frames(C, seed=2)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
ks = bench.kernel_breakdown(torch, _lib.lib(), model, pc, torch.from_numpy(b).to(dev), torch.from_numpy(p).to(dev), smpl, flush, bench.read_peaks(), C, N)
for k in ks:
    r = k.get("roofline")
    print(f"{k['ms']:8.3f} ms  {k['name']:60s} {('%.4f %s' % (r['frac'], r['bound'])) if r else ''}  {k.get('note','')}")
