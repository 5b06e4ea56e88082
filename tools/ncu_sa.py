#!/usr/bin/env python
"""One launch of each grouped-MLP (g4d_sa_mlp_max) branch at c3 sizes inside a cudaProfilerStart/Stop range (ncu target)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from garment4d_b200 import _lib, synthetic
from garment4d_b200.encoder import Pointnet2MSGSEG
from garment4d_b200.pointnet2 import pointnet2_utils as pu

C, N = 240, 8192
dev = torch.device("cuda:0")
L = _lib.lib()
torch.manual_seed(1234)
model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(dev).eval()
pc = torch.from_numpy(bench.make_inputs("body", 4234, C, N)).to(dev)
calls = []
with torch.no_grad():
    xyz, feats = pc, None
    for lvl, sa in enumerate(model.SA_modules):
        P = sa.npoint
        _, new_xyz = pu.furthest_point_sample_and_gather(xyz, P)
        g0, g1 = sa.groupers
        idxs = pu.ball_query_pair(g0.radius, g0.nsample, g1.radius, g1.nsample, xyz, new_xyz)
        c_in = 0 if feats is None else feats.shape[1]
        nx, nf = sa(xyz, feats)
        fpm = None if feats is None else pu.point_major_of(feats)
        ctot = nf.shape[1]
        out_cm = torch.empty_like(nf); out_pm = torch.empty(C, P, ctot, dtype=torch.float16, device=dev)
        off = 0
        for i, idx in enumerate(idxs):
            br = sa._branch(i, c_in, dev)
            calls.append((br, xyz, new_xyz, idx, fpm, out_cm, out_pm, ctot, off, xyz.shape[1], P))
            off += br.c_out
        xyz, feats = nx, nf
    def run_all():
        for br, x, nx_, idx, fpm, ocm, opm, ctot, off, n_in, P in calls:
            rc = L.g4d_sa_mlp_max(ctypes.byref(br.desc), _lib.ptr(br.params), C, n_in, P, _lib.ptr(x), _lib.ptr(nx_), _lib.ptr(idx), _lib.ptr(fpm),
                                  _lib.ptr(ocm), _lib.ptr(opm), ctot, off, _lib.stream_ptr())
            _lib.check(rc, "sa_mlp")
    run_all(); torch.cuda.synchronize()
    # plain timing (not under the profiler when run without ncu)
    for br, *_ in calls:
        pass
    evs = []
    for c in calls:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        br, x, nx_, idx, fpm, ocm, opm, ctot, off, n_in, P = c
        s.record()
        for _ in range(5):
            L.g4d_sa_mlp_max(ctypes.byref(br.desc), _lib.ptr(br.params), C, n_in, P, _lib.ptr(x), _lib.ptr(nx_), _lib.ptr(idx), _lib.ptr(fpm),
                             _lib.ptr(ocm), _lib.ptr(opm), ctot, off, _lib.stream_ptr())
        e.record(); evs.append((br.desc, s, e))
    torch.cuda.synchronize()
    print("timing ms:", ["%d+3->%d,%d,%d K=%d: %.3f" % (d.c_in, d.c1, d.c2, d.c3, d.nsample, s.elapsed_time(e) / 5) for d, s, e in evs],
          "total %.3f" % sum(s.elapsed_time(e) / 5 for _, s, e in evs))
    torch.cuda.profiler.start()
    run_all(); torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ncu_sa: done")
