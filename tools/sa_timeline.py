#!/usr/bin/env python
"""clock64() timeline of slot 0 of CTA 0 of g4d_sa_mlp_max (g4d_debug_timeline) for one branch at c3 sizes: where a tile's
three hand-offs spend their cycles.   python tools/sa_timeline.py [branch 0..5]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from garment4d_b200 import _lib
from garment4d_b200.encoder import Pointnet2MSGSEG
from garment4d_b200.pointnet2 import pointnet2_utils as pu

which = [int(a) for a in sys.argv[1:]] or [1, 3, 5]
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
    for bi in which:
        br, x, nx_, idx, fpm, ocm, opm, ctot, off, n_in, P = calls[bi]
        buf = torch.zeros(16 * 16, dtype=torch.int64, device=dev)
        def run():
            rc = L.g4d_sa_mlp_max(ctypes.byref(br.desc), _lib.ptr(br.params), C, n_in, P, _lib.ptr(x), _lib.ptr(nx_), _lib.ptr(idx), _lib.ptr(fpm),
                                  _lib.ptr(ocm), _lib.ptr(opm), ctot, off, _lib.stream_ptr())
            _lib.check(rc, "sa_mlp")
        run(); torch.cuda.synchronize()
        L.g4d_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
        run(); torch.cuda.synchronize()
        L.g4d_debug_timeline(None)
        t = buf.cpu().numpy().reshape(16, 16)
        d = br.desc
        print(f"== branch {bi}: {d.c_in}+3->{d.c1},{d.c2},{d.c3} K={d.nsample}   (cycles relative to the tile's layer-1 issue)")
        print("   tile:  L1 wait>commit | e1 wake-commit  e1 work | L2 wake-arrive L2 issue | e2 wake-commit e2 work | L3 wake-arrive L3 issue | e3 wake-commit e3 work | next L1 - this L1")
        for r in range(2, 14):
            w1, c1, w2, c2, w3, c3 = t[r, 0:6]
            e1w, e1a, e2w, e2a, e3w, e3a = t[r, 8:14]
            nxt = t[r + 1, 0]
            ld, st, fe, e3l = t[r, 6], t[r, 7], t[r, 14], t[r, 15]
            print(f"   {r:4d}: {c1 - w1:6d} | {e1w - c1:6d} {e1a - e1w:6d} | {w2 - e1a:6d} {c2 - w2:6d} | {e2w - c2:6d} {e2a - e2w:6d} | {w3 - e2a:6d} {c3 - w3:6d} | "
                  f"{e3w - c3:6d} {e3a - e3w:6d} | {nxt - w1:6d}   e1: ld {ld - e1w} sts {st - ld} fence {fe - st} arrive {e1a - fe}  e3 loop {e3l - e3w}")
