#!/usr/bin/env python
"""Generates the committed golden vectors under tests/golden/.

Two sources, both the reference itself:

1. ``lbs_ref_*.npz`` -- produced HERE (build container, CPU) by importing the
   reference's own ``smplx/smplx/lbs.py`` from /root/reference and running it on
   seeded synthetic SMPL-shaped inputs (oracle.lbs.synthetic_smpl).
       python tests/golden/make_golden.py lbs

2. ``pointnet2_ref_*.npz`` -- produced ON THE GPU BOX by running the reference's
   own CUDA kernels (oracle/_ref/libpointnet2_ref.so = the reference .cu files
   compiled unmodified by oracle/build_ref.sh) on seeded inputs; written to
   gpurun_out/golden/ and copied from there into tests/golden/.
       gpurun -- python tests/golden/make_golden.py pointnet2

The generating inputs are either stored in the .npz (small cases) or are a
pure function of the seed recorded in it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def make_lbs():
    import torch
    sys.path.insert(0, "/root/reference/smplx")
    from smplx.lbs import (lbs, batch_rodrigues, batch_rigid_transform,  # noqa: E402  (reference code)
                           vertices2joints, vertices2jointsB, blend_shapes)
    from oracle import lbs as olbs

    torch.set_num_threads(1)
    t = torch.from_numpy

    # (a) small model, everything stored
    m = olbs.synthetic_smpl(V=160, seed=11, sparse_weights=True)
    betas, pose = olbs.synthetic_frames(5, seed=12)
    with torch.no_grad():
        rot = batch_rodrigues(t(pose).view(-1, 3)).view(5, 24, 3, 3)
        v1, j1 = lbs(t(betas), t(pose), t(m["v_template"]), t(m["shapedirs"]), t(m["posedirs"]),
                     t(m["J_regressor"]), t(m["parents"]), t(m["lbs_weights"]), pose2rot=True)
        v2, j2 = lbs(t(betas), rot, t(m["v_template"]), t(m["shapedirs"]), t(m["posedirs"]),
                     t(m["J_regressor"]), t(m["parents"]), t(m["lbs_weights"]), pose2rot=False)
        v_shaped = t(m["v_template"]) + blend_shapes(t(betas), t(m["shapedirs"]))
        J = vertices2joints(t(m["J_regressor"]), v_shaped)
        JB = vertices2jointsB(t(m["J_regressor"])[None].expand(5, -1, -1).contiguous(), v_shaped)
        pj, A = batch_rigid_transform(rot, J, t(m["parents"]))
    np.savez_compressed(os.path.join(HERE, "lbs_ref_small.npz"), betas=betas, pose=pose, **m,
                        rot_mats=rot.numpy(), verts_pose2rot=v1.numpy(), joints_pose2rot=j1.numpy(),
                        verts_rotmat=v2.numpy(), joints_rotmat=j2.numpy(), v_shaped=v_shaped.numpy(),
                        J=J.numpy(), JB=JB.numpy(), posed_joints=pj.numpy(), A=A.numpy())

    # (b) SMPL-sized model (V=6890): inputs are a function of the seeds; outputs stored strided
    out = {}
    for tag, sparse in (("sparse", True), ("dense", False)):
        seed_model, seed_frames, F, stride = 21, 22, 3, 13
        m = olbs.synthetic_smpl(V=6890, seed=seed_model, sparse_weights=sparse)
        betas, pose = olbs.synthetic_frames(F, seed=seed_frames)
        with torch.no_grad():
            v, j = lbs(t(betas), t(pose), t(m["v_template"]), t(m["shapedirs"]), t(m["posedirs"]),
                       t(m["J_regressor"]), t(m["parents"]), t(m["lbs_weights"]), pose2rot=True)
        out[f"{tag}_verts_strided"] = v.numpy()[:, ::stride].copy()
        out[f"{tag}_joints"] = j.numpy()
        out[f"{tag}_verts_abs_sum"] = np.abs(v.numpy().astype(np.float64)).sum(axis=(1, 2))
    np.savez_compressed(os.path.join(HERE, "lbs_ref_smpl.npz"), seed_model=21, seed_frames=22, F=3, stride=13, **out)
    print("wrote lbs_ref_small.npz, lbs_ref_smpl.npz")


def golden_clouds(seed, B, N, dup_frac=0.05):
    """Seeded clouds: uniform cube and a 'body' surface, with exact duplicate points (FPS tie-breaks)."""
    rs = np.random.RandomState(seed)
    cube = rs.rand(B, N, 3).astype(np.float32)
    u = rs.rand(B, N).astype(np.float32) * np.float32(2 * np.pi)
    h = rs.rand(B, N).astype(np.float32) * np.float32(1.7)
    r = np.float32(0.15) + np.float32(0.05) * np.sin(h * np.float32(7.0)).astype(np.float32)
    body = np.stack([r * np.cos(u), h, r * np.sin(u)], axis=-1).astype(np.float32)
    nd = int(N * dup_frac)
    for arr in (cube, body):
        for b in range(B):
            src = rs.randint(0, N, nd)
            dst = rs.randint(0, N, nd)
            arr[b, dst] = arr[b, src]
    return cube, body


def make_pointnet2():
    """Runs the REFERENCE CUDA kernels (oracle/_ref) -- needs a GPU."""
    import torch
    from oracle import refgpu
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    dev = "cuda:0"
    res = {}
    for tag, (seed, B, N, m, radius, K) in {
        "c1": (101, 2, 1024, 256, 0.2, 32),           # BASELINE config 1 geometry
        "n1000": (102, 2, 1000, 200, 0.15, 16),       # bs=512 < N, ragged tail
        "n8192": (103, 1, 8192, 1024, 0.1, 32),       # SA1b geometry
        "n37": (104, 3, 37, 9, 0.5, 8),               # tiny, bs=32
    }.items():
        for kind, xyz in zip(("cube", "body"), golden_clouds(seed, B, N)):
            x = torch.from_numpy(xyz).to(dev)
            idx = refgpu.furthest_point_sample(x, m)
            new_xyz = refgpu.gather_operation(x.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
            bq = refgpu.ball_query(radius, K, x, new_xyz)
            d2, nn_idx = refgpu.three_nn_raw(x, new_xyz)
            w = torch.rand(B, N, 3, generator=torch.Generator().manual_seed(seed)).to(dev)
            feats = torch.randn(B, 5, m, generator=torch.Generator().manual_seed(seed + 1)).to(dev)
            interp = refgpu.three_interpolate(feats, nn_idx, w)
            p = f"{tag}_{kind}_"
            res[p + "fps_idx"] = idx.cpu().numpy()
            res[p + "ball_idx"] = bq.cpu().numpy()
            res[p + "nn_dist2"] = d2.cpu().numpy()
            res[p + "nn_idx"] = nn_idx.cpu().numpy()
            res[p + "interp"] = interp.cpu().numpy()
            res[p + "meta"] = np.array([seed, B, N, m, K], np.int64)
            res[p + "radius"] = np.float32(radius)
    np.savez_compressed(os.path.join(outdir, "pointnet2_ref_kernels.npz"), **res)
    print("wrote", os.path.join(outdir, "pointnet2_ref_kernels.npz"))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "lbs"
    {"lbs": make_lbs, "pointnet2": make_pointnet2}[what]()
