#!/usr/bin/env python
"""Timing of lbs_garment_interpolation at a CLOTH3D-like size against the reference's torch composition
(mesh_encoder.py:339-389: the (F, body_v, K, 24) repeat + gather + sum and 100 torch.spmm steps) run on the same K-NN result
(chamferdist is not installed, so the K-NN itself is timed for ours only).   python tools/garment_lbs_bench.py [B T G K]"""
import os, sys, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from garment4d_b200 import mesh_ops
from garment4d_b200.synthetic import synthetic_smpl

B, T, G, K = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 8, 4900, 256)
P, J = 6890, 24
dev = torch.device("cuda:0")
smpl = synthetic_smpl(V=P, J=J, seed=3)
rs = np.random.RandomState(0)
nu = int(round(G ** 0.5)); G = nu * nu
idx = np.arange(G).reshape(nu, nu)
rows = np.concatenate([idx[:-1, :].ravel(), idx[:, :-1].ravel(), idx[:-1, :-1].ravel()])
cols = np.concatenate([idx[1:, :].ravel(), idx[:, 1:].ravel(), idx[1:, 1:].ravel()])
adj = torch.sparse_coo_tensor(torch.tensor(np.stack([np.concatenate([rows, cols]), np.concatenate([cols, rows])])), torch.ones(2 * rows.size), (G, G)).coalesce()
op = mesh_ops.smoothing_operator(adj, dev)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
body = t((smpl["v_template"][None] + rs.randn(B, P, 3) * 0.002).astype(np.float32))
garment = body[:, rs.choice(P, G)] + t((rs.randn(B, G, 3) * 0.01).astype(np.float32))
root = t((rs.randn(B, 3) * 0.05).astype(np.float32))
zero = body[:, None].repeat(1, T, 1, 1).contiguous()
pose = t((rs.randn(B, T, 72) * 0.3).astype(np.float32))
Jreg = t(smpl["J_regressor"])[None, None].repeat(B, T, 1, 1).contiguous()
Wb = t(smpl["lbs_weights"])[None, None].repeat(B, T, 1, 1).contiguous()
model = types.SimpleNamespace(parents=t(smpl["parents"].astype(np.int64)))


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


ours = timed(lambda: mesh_ops.lbs_garment_interpolation(garment - root[:, None], body, root, zero, model, pose, Jreg, Wb, K=K, smooth=op))
knn = timed(lambda: mesh_ops.knn_points(garment, body, K=K))
nnk = mesh_ops.knn_points(garment, body, K=K)
rowptr, col, val = op
crow = torch.repeat_interleave(torch.arange(G, device=dev), (rowptr[1:] - rowptr[:-1]).long())
adj_t = torch.sparse_coo_tensor(torch.stack([crow, col.long()]), val, (G, G)).coalesce()


def reference_weights():          # mesh_encoder.py:371-389 as written
    iw = 1 / nnk.dists.reshape(B, -1, K, 1)
    iw[torch.where(torch.isinf(iw))] = 0
    iw = iw / iw.sum(-2, keepdim=True)
    iw[torch.where(torch.isinf(iw))] = 0
    W = Wb.reshape(B * T, -1, 1, J).repeat(1, 1, K, 1)
    nn_W = torch.gather(W, 1, nnk.idx.reshape(B, 1, -1, K, 1).repeat(1, T, 1, 1, J).reshape(B * T, -1, K, J))
    nn_W = (nn_W * iw.reshape(B, 1, -1, K, 1).repeat(1, T, 1, 1, 1).reshape(B * T, -1, K, 1)).sum(-2)
    for _ in range(100):
        nn_W = nn_W + 0.1 * torch.spmm(adj_t, nn_W.transpose(0, 1).reshape(-1, B * T * J)).reshape(-1, B * T, J).transpose(0, 1)
    return nn_W


try:
    ref = timed(reference_weights, reps=1)
    ref_s = f"{ref:.2f} ms"
except torch.OutOfMemoryError:
    ref_s = "out of memory"
print(f"lbs_garment_interpolation B={B} T={T} G={G} P={P} K={K}: ours (whole function) {ours:.2f} ms, of which K-NN {knn:.2f} ms; "
      f"the reference's weight gather + 100 spmm steps alone (torch, same K-NN result): {ref_s}")
