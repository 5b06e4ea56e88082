"""Prints the per-batch clock64 timeline of CTA 0 of each sa_mlp_max branch (debug aid)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from garment4d_b200 import _lib, synthetic
from garment4d_b200.encoder import Pointnet2MSGSEG
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).eval()
C = 240
pc = torch.from_numpy(np.tile(synthetic.body_clouds(1, 16, 8192), (C // 16, 1, 1))).to(dev)
L = _lib.lib()
with torch.no_grad():
    model.sa_stack(pc)
    xyz, feats = pc.contiguous(), None
    for lvl, sa in enumerate(model.SA_modules):
        buf = torch.zeros(25 * 16, dtype=torch.int64, device=dev)
        # run the module once with the timeline on: both branches write into the same buffer -> run branch by branch
        from garment4d_b200.pointnet2 import pointnet2_utils as pu
        _, new_xyz = pu.furthest_point_sample_and_gather(xyz, sa.npoint)
        g0, g1 = sa.groupers
        idxs = pu.ball_query_pair(g0.radius, g0.nsample, g1.radius, g1.nsample, xyz, new_xyz)
        c_in = 0 if feats is None else feats.shape[1]
        feat_pm = None if feats is None else feats._g4d_pm
        nx, nf = sa(xyz, feats)
        ctot = nf.shape[1]
        out_cm = torch.empty_like(nf); out_pm = torch.empty(C, sa.npoint, ctot, dtype=torch.float16, device=dev)
        off = 0
        for i, idx in enumerate(idxs):
            br = sa._branch(i, c_in, dev)
            buf.zero_(); torch.cuda.synchronize()
            L.g4d_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
            rc = L.g4d_sa_mlp_max(ctypes.byref(br.desc), _lib.ptr(br.params), C, xyz.shape[1], sa.npoint, _lib.ptr(xyz), _lib.ptr(new_xyz),
                                  _lib.ptr(idx), _lib.ptr(feat_pm), _lib.ptr(out_cm), _lib.ptr(out_pm), ctot, off, _lib.stream_ptr())
            torch.cuda.synchronize()
            L.g4d_debug_timeline(None)
            t = buf.cpu().numpy().reshape(25, 16)
            d = br.desc
            print(f"--- L{lvl} branch {i}: c_in={d.c_in} mlp=({d.c1},{d.c2},{d.c3}) K={d.nsample}")
            for b in range(2, 8):
                m, e = t[b, :6], t[b, 8:14]
                if m[0] == 0: continue
                base = m[0]
                print("  batch", b, "MMA:", [int(x - base) for x in m], " EPI:", [int(x - base) for x in e], " next batch starts:", int(t[b + 1, 0] - base))
            off += br.c_out
        xyz, feats = nx, nf
