"""CPU tests (-m "not gpu"): the oracle against the golden vectors produced by the reference itself, and against
independent statements of the same algorithms."""
import os

import numpy as np
import pytest

from oracle import lbs as olbs
from oracle import pointnet2 as orc
from tests.golden.make_golden import golden_clouds
from tests.util import clouds

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_opt_n_threads_matches_reference_rule():
    # cuda_utils.h:10-14; SURVEY probe: 6890/8192/16384 -> 1024, 1024 -> 1024, 256 -> 256
    for n, want in [(6890, 1024), (8192, 1024), (16384, 1024), (1024, 1024), (256, 256), (1000, 512), (37, 32), (1, 1), (3, 2)]:
        assert orc.opt_n_threads(n) == want


@pytest.mark.parametrize("tag", ["c1", "n1000", "n37", "n8192"])
@pytest.mark.parametrize("kind", ["cube", "body"])
def test_oracle_vs_reference_kernel_golden(tag, kind):
    """tests/golden/pointnet2_ref_kernels.npz = outputs of the reference's own CUDA kernels (oracle/_ref) on B200."""
    G = np.load(os.path.join(GOLD, "pointnet2_ref_kernels.npz"))
    p = f"{tag}_{kind}_"
    seed, B, N, m, K = (int(v) for v in G[p + "meta"])
    radius = float(G[p + "radius"])
    xyz = golden_clouds(seed, B, N)[0 if kind == "cube" else 1]
    if tag == "n8192" and kind == "body":
        xyz, B = xyz[:1], 1
    idx = orc.furthest_point_sample(xyz, m)
    assert np.array_equal(idx, G[p + "fps_idx"][:B]), "oracle FPS != reference kernel"
    assert np.array_equal(orc.furthest_point_sample(xyz, m, fast=True), idx), "vectorised FPS != line-by-line FPS"
    new_xyz = np.ascontiguousarray(orc.gather_operation(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx).transpose(0, 2, 1))
    assert np.array_equal(orc.ball_query(radius, K, xyz, new_xyz), G[p + "ball_idx"][:B])
    dist, nn_idx = orc.three_nn(xyz, new_xyz)
    assert np.array_equal(nn_idx, G[p + "nn_idx"][:B])
    assert np.array_equal(dist, np.sqrt(G[p + "nn_dist2"][:B]))
    import torch
    w = torch.rand(G[p + "meta"][1], N, 3, generator=torch.Generator().manual_seed(seed)).numpy()[:B]
    feats = torch.randn(int(G[p + "meta"][1]), 5, m, generator=torch.Generator().manual_seed(seed + 1)).numpy()[:B]
    assert np.array_equal(orc.three_interpolate(feats, nn_idx, w), G[p + "interp"][:B])


def test_fps_tie_break_is_bit_reversed_slot():
    """All points coincide except index 0: every distance ties, so the pick order is purely the tie-break.
    Tree semantics (sampling_gpu.cu:86-91,143-203): smallest bit-reversed (k mod bs), then smallest k div bs."""
    N = 64
    xyz = np.zeros((1, N, 3), np.float32)
    idx = orc.furthest_point_sample(xyz, 4)
    # all temps equal 0 after step 1 -> winner is key-min = point 0 every time
    assert idx.tolist() == [[0, 0, 0, 0]]
    xyz[0, 1] = xyz[0, 2] = 1.0          # slots 1 and 2 tie at distance 3: bitrev6(1)=32 > bitrev6(2)=16 -> 2 wins
    assert orc.furthest_point_sample(xyz, 2)[0, 1] == 2
    assert orc.fps_numpy_keyed(xyz, 2)[0, 1] == 2


@pytest.mark.parametrize("N,m", [(37, 9), (256, 64), (1000, 100), (1500, 64)])
def test_fps_c_oracle_vs_numpy_keyed(N, m):
    xyz = clouds(5, 2, N, "body", dup_frac=0.1)
    assert np.array_equal(orc.furthest_point_sample(xyz, m), orc.fps_numpy_keyed(xyz, m))


def test_ball_query_semantics():
    xyz = np.array([[[0, 0, 0], [0.05, 0, 0], [5, 5, 5], [0.01, 0, 0], [0.02, 0, 0]]], np.float32)
    q = np.array([[[0, 0, 0], [9, 9, 9]]], np.float32)
    idx = orc.ball_query(0.1, 3, xyz, q)
    assert idx[0, 0].tolist() == [0, 1, 3]            # first 3 hits in index order (4 is a hit too, dropped)
    assert idx[0, 1].tolist() == [0, 0, 0]            # no hit: row untouched (zero-filled by the caller)
    assert orc.ball_query(0.1, 6, xyz, q)[0, 0].tolist() == [0, 1, 3, 4, 0, 0]   # padded with the first hit
    # strict '<' against rn(r*r) in float32
    r = np.float32(0.05)
    xyz2 = np.array([[[r, 0, 0]]], np.float32)
    assert orc.ball_query(float(r), 1, xyz2, np.zeros((1, 1, 3), np.float32))[0, 0, 0] == 0  # d2 == r2 -> not a hit -> stays 0 (ambiguous with idx 0)
    xyz3 = np.array([[[9, 9, 9], [r, 0, 0]]], np.float32)
    assert orc.ball_query(float(r), 1, xyz3, np.zeros((1, 1, 3), np.float32))[0, 0, 0] == 0  # point 1 at exactly r is excluded


def test_three_nn_fewer_than_three_known():
    d, i = orc.three_nn(np.zeros((1, 2, 3), np.float32), np.ones((1, 2, 3), np.float32))
    assert np.isinf(d[..., 2]).all() and (i[..., 2] == 0).all()
    assert np.allclose(d[..., 0] ** 2, 3.0) and (i[..., 0] == 0).all() and (i[..., 1] == 1).all()


def test_grads_are_adjoint_of_forward():
    rs = np.random.RandomState(0)
    B, C, N, P, S = 2, 3, 50, 7, 4
    idx = rs.randint(0, N, (B, P, S)).astype(np.int32)
    f = rs.randn(B, C, N).astype(np.float32)
    g = rs.randn(B, C, P, S).astype(np.float32)
    lhs = (orc.grouping_operation(f, idx).astype(np.float64) * g).sum()
    rhs = (f.astype(np.float64) * orc.grouping_operation_grad(g, idx, N)).sum()
    assert abs(lhs - rhs) < 1e-3
    w = rs.rand(B, N, 3).astype(np.float32)
    i3 = rs.randint(0, P, (B, N, 3)).astype(np.int32)
    fk = rs.randn(B, C, P).astype(np.float32)
    go = rs.randn(B, C, N).astype(np.float32)
    lhs = (orc.three_interpolate(fk, i3, w).astype(np.float64) * go).sum()
    rhs = (fk.astype(np.float64) * orc.three_interpolate_grad(go, i3, w, P)).sum()
    assert abs(lhs - rhs) < 1e-3


# ---- LBS oracle vs the reference's lbs.py (golden vectors generated by tests/golden/make_golden.py lbs) ----

def test_lbs_oracle_vs_reference_golden_small():
    g = np.load(os.path.join(GOLD, "lbs_ref_small.npz"))
    m = {k: g[k] for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights")}
    v, j = olbs.lbs(g["betas"], g["pose"], **m, pose2rot=True)
    assert np.abs(v - g["verts_pose2rot"]).max() < 2e-6 and np.abs(j - g["joints_pose2rot"]).max() < 2e-6
    v, j = olbs.lbs(g["betas"], g["rot_mats"], **m, pose2rot=False)
    assert np.abs(v - g["verts_rotmat"]).max() < 2e-6
    assert np.abs(olbs.batch_rodrigues(g["pose"].reshape(-1, 3)).reshape(g["rot_mats"].shape) - g["rot_mats"]).max() < 1e-6
    pj, A = olbs.batch_rigid_transform(g["rot_mats"], g["J"], g["parents"])
    assert np.abs(pj - g["posed_joints"]).max() < 1e-6 and np.abs(A - g["A"]).max() < 2e-6
    assert np.abs(olbs.vertices2joints(g["J_regressor"], g["v_shaped"]) - g["J"]).max() < 1e-6


@pytest.mark.parametrize("tag,sparse", [("sparse", True), ("dense", False)])
def test_lbs_oracle_vs_reference_golden_smpl_size(tag, sparse):
    G = np.load(os.path.join(GOLD, "lbs_ref_smpl.npz"))
    m = olbs.synthetic_smpl(seed=int(G["seed_model"]), sparse_weights=sparse)
    b, p = olbs.synthetic_frames(int(G["F"]), seed=int(G["seed_frames"]))
    v, j = olbs.lbs(b, p, **m)
    assert np.abs(v[:, ::int(G["stride"])] - G[tag + "_verts_strided"]).max() < 2e-6
    assert np.abs(j - G[tag + "_joints"]).max() < 2e-6
    assert np.allclose(np.abs(v.astype(np.float64)).sum(axis=(1, 2)), G[tag + "_verts_abs_sum"], rtol=1e-6)


def test_lbs_identity_pose_is_shape_blend_only():
    m = olbs.synthetic_smpl(V=200, seed=2)
    b, _ = olbs.synthetic_frames(3, seed=3)
    v, j = olbs.lbs(b, np.zeros((3, 72), np.float32), **m)
    v_shaped = m["v_template"] + olbs.blend_shapes(b, m["shapedirs"])
    assert np.abs(v - v_shaped).max() < 1e-5          # rodrigues(0 + 1e-8) is the identity to ~1e-8
    assert np.abs(j - olbs.vertices2joints(m["J_regressor"], v_shaped)).max() < 1e-5
