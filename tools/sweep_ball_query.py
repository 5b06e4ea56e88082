#!/usr/bin/env python
"""BASELINE config c5 (N = 16384 points per cloud): roofline sweep of the fused ball-query + group operator over radius and
nsample (QueryAndGroup: ball query + one grouping pass that materialises the (B, 3+C, P, K) tensor), plus FPS at N = 16384.
Prints a markdown table: time, algorithmic bytes (12N + 12P + 4CN + 4PK + 4(C+3)PK per cloud), GB/s
and the fraction of the measured HBM peak.   python tools/sweep_ball_query.py [clouds] > profiles/r02_c5_sweep.md"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from garment4d_b200.pointnet2 import pointnet2_utils as pu

C = int(sys.argv[1]) if len(sys.argv) > 1 else 240
N, P = 16384, 1024
dev = torch.device("cuda:0")
peak = bench.read_peaks()["hbm_gbs"]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(0.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps


print(f"# c5 sweep: {C} clouds x {N} points, {P} centroids per cloud (B200, HBM peak {peak:.0f} GB/s measured)\n")
for kind in ("body", "cube"):
    xyz = torch.from_numpy(bench.make_inputs(kind, 5234, C, N)).to(dev)
    ms = t(lambda: pu.furthest_point_sample_and_gather(xyz, P))
    _, new_xyz = pu.furthest_point_sample_and_gather(xyz, P)
    print(f"## {kind} clouds\n\nFPS {N} -> {P}: {ms:.3f} ms ({ms * 1e6 / (P - 1):.0f} ns per serial step)\n")
    print("| radius | nsample | feature channels | ball query ms | query+group ms | algorithmic MB/cloud | GB/s | of HBM peak | mean hits (<= nsample) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for cf in (0, 96):
        feats = torch.randn(C, cf, N, device=dev) if cf else None
        for radius in (0.025, 0.05, 0.1, 0.2, 0.4):
            for K in (16, 32, 64):
                ms_bq = t(lambda: pu.ball_query(radius, K, xyz, new_xyz))
                ms_qg = t(lambda: pu.QueryAndGroup(radius, K)(xyz, new_xyz, feats))
                idx = pu.ball_query(radius, K, xyz, new_xyz)
                hits = float((idx != idx[:, :, :1]).sum(2).float().mean().item()) + 1.0
                by = 12 * N + 12 * P + 4 * cf * N + 4 * P * K + 4 * (cf + 3) * P * K
                gbs = by * C / (ms_qg * 1e-3) / 1e9
                print(f"| {radius} | {K} | {cf} | {ms_bq:.3f} | {ms_qg:.3f} | {by / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.3f} | {hits:.1f} |")
    print()
