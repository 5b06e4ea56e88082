#!/usr/bin/env python
"""Small instances of the kernels added or changed this round, for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py [which ...]"""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from garment4d_b200 import mesh_ops
from garment4d_b200.pointnet2 import pointnet2_modules as pm, pointnet2_utils as pu
from garment4d_b200.synthetic import synthetic_smpl
from tests.util import clouds

which = set(sys.argv[1:]) or {"knn", "garment", "fps", "bq", "group", "fp", "sa"}
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rs = np.random.RandomState(0)
if "knn" in which:
    for (N, P, K) in ((70, 700, 128), (33, 6890, 256), (5, 40, 3)):
        r = mesh_ops.knn_points(t(rs.randn(2, N, 3).astype(np.float32)), t(rs.randn(2, P, 3).astype(np.float32)), K=K)
        torch.cuda.synchronize(); print("knn", N, P, K, float(r.dists.sum()))
if "garment" in which:
    B, T, G, P, J, K = 1, 2, 64, 700, 24, 16
    smpl = synthetic_smpl(V=P, J=J, seed=3)
    body = t(smpl["v_template"][None])
    adj = torch.zeros(G, G); i = torch.arange(G - 1); adj[i, i + 1] = 1; adj[i + 1, i] = 1
    out, nn, s1 = mesh_ops.lbs_garment_interpolation(body[:, :G] + 0.01, body, torch.zeros(B, 3, device=dev), body[:, None].repeat(1, T, 1, 1),
                                                     types.SimpleNamespace(parents=t(smpl["parents"].astype(np.int64))), t((rs.randn(B, T, 72) * 0.2).astype(np.float32)),
                                                     t(smpl["J_regressor"])[None, None].repeat(B, T, 1, 1), t(smpl["lbs_weights"])[None, None].repeat(B, T, 1, 1),
                                                     K=K, smooth=mesh_ops.smoothing_operator(adj, dev), smooth_iters=5)
    torch.cuda.synchronize(); print("garment", float(out.abs().sum()))
x = t(clouds(3, 2, 8192, "body"))
if "fps" in which:
    idx, nx = pu.furthest_point_sample_and_gather(x, 64)
    torch.cuda.synchronize(); print("fps", int(idx.sum()))
if "bq" in which:
    _, nx = pu.furthest_point_sample_and_gather(x, 256)
    a, b = pu.ball_query_pair(0.1, 16, 0.2, 32, x, nx)
    torch.cuda.synchronize(); print("bq", int(a.sum()), int(b.sum()))
if "group" in which:
    f = torch.randn(2, 70, 8192, device=dev)
    _, nx = pu.furthest_point_sample_and_gather(x, 50)
    g = pu.QueryAndGroup(0.2, 16)(x, nx, f)
    torch.cuda.synchronize(); print("group", float(g.abs().sum()))
if "fp" in which:
    torch.manual_seed(0)
    for n, m, c2, c1, mlp in ((256, 64, 384, 192, [576, 512, 256]), (300, 70, 256, 96, [352, 256, 128])):
        mod = pm.PointnetFPModule(mlp=list(mlp), bn=True).to(dev).eval()
        u = t(clouds(12, 2, n, "body")); k = u[:, :m].contiguous()
        out = mod(u, k, torch.randn(2, c1, n, device=dev), torch.randn(2, c2, m, device=dev))
        torch.cuda.synchronize(); print("fp", n, float(out.abs().mean()))
if "sa" in which:
    torch.manual_seed(0)
    sa = pm.PointnetSAModuleMSG(npoint=128, radii=[0.1, 0.2], nsamples=[16, 32], mlps=[[0, 16, 16, 32], [0, 32, 32, 64]], use_xyz=True, bn=True).to(dev).eval()
    nx, nf = sa(x[:, :2048].contiguous(), None)
    torch.cuda.synchronize(); print("sa", float(nf.abs().mean()))
