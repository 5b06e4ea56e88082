"""-m gpu: LBS kernels vs the numpy oracle and the golden vectors generated from the reference's lbs.py.
Tolerance: 1e-5 absolute on vertex positions (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import lbs as olbs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5

from garment4d_b200 import lbs as glbs   # noqa: E402


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _model(m, dev):
    return [_t(m[k], dev) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights")]


def test_lbs_small_golden(cuda):
    g = np.load(os.path.join(GOLD, "lbs_ref_small.npz"))
    m = _model(g, cuda)
    v, j = glbs.lbs(_t(g["betas"], cuda), _t(g["pose"], cuda), *m, pose2rot=True)
    assert np.abs(v.cpu().numpy() - g["verts_pose2rot"]).max() <= TOL
    assert np.abs(j.cpu().numpy() - g["joints_pose2rot"]).max() <= TOL
    v, j = glbs.lbs(_t(g["betas"], cuda), _t(g["rot_mats"], cuda), *m, pose2rot=False)
    assert np.abs(v.cpu().numpy() - g["verts_rotmat"]).max() <= TOL
    assert np.abs(j.cpu().numpy() - g["joints_rotmat"]).max() <= TOL
    R = glbs.batch_rodrigues(_t(g["pose"].reshape(-1, 3), cuda))
    assert np.abs(R.cpu().numpy().reshape(g["rot_mats"].shape) - g["rot_mats"]).max() <= 1e-6
    pj, A = glbs.batch_rigid_transform(_t(g["rot_mats"], cuda), _t(g["J"], cuda), _t(g["parents"], cuda))
    assert np.abs(pj.cpu().numpy() - g["posed_joints"]).max() <= 1e-6
    assert np.abs(A.cpu().numpy() - g["A"]).max() <= 1e-6
    J = glbs.vertices2joints(_t(g["J_regressor"], cuda), _t(g["v_shaped"], cuda))
    assert np.abs(J.cpu().numpy() - g["J"]).max() <= 1e-6
    JB = glbs.vertices2jointsB(_t(np.broadcast_to(g["J_regressor"], (5,) + g["J_regressor"].shape).copy(), cuda), _t(g["v_shaped"], cuda))
    assert np.abs(JB.cpu().numpy() - g["JB"]).max() <= 1e-6


@pytest.mark.parametrize("tag,sparse", [("sparse", True), ("dense", False)])
def test_lbs_smpl_size_golden(cuda, tag, sparse):
    G = np.load(os.path.join(GOLD, "lbs_ref_smpl.npz"))
    m = olbs.synthetic_smpl(seed=int(G["seed_model"]), sparse_weights=sparse)
    b, p = olbs.synthetic_frames(int(G["F"]), seed=int(G["seed_frames"]))
    v, j = glbs.lbs(_t(b, cuda), _t(p, cuda), *_model(m, cuda))
    assert np.abs(v.cpu().numpy()[:, ::int(G["stride"])] - G[tag + "_verts_strided"]).max() <= TOL
    assert np.abs(j.cpu().numpy() - G[tag + "_joints"]).max() <= TOL


@pytest.mark.parametrize("F", [1, 7, 64, 240])
def test_lbs_vs_oracle_batched(cuda, F):
    m = olbs.synthetic_smpl(seed=5)
    b, p = olbs.synthetic_frames(F, seed=6 + F)
    v, j = glbs.lbs(_t(b, cuda), _t(p, cuda), *_model(m, cuda))
    wv, wj = olbs.lbs(b, p, **m)
    assert np.abs(v.cpu().numpy() - wv).max() <= TOL
    assert np.abs(j.cpu().numpy() - wj).max() <= TOL
    # broadcast betas (1, NB) against F poses, as lbs.py:201 allows
    v1, _ = glbs.lbs(_t(b[:1], cuda), _t(p, cuda), *_model(m, cuda))
    wv1, _ = olbs.lbs(np.repeat(b[:1], F, 0), p, **m)
    assert np.abs(v1.cpu().numpy() - wv1).max() <= TOL


def test_skin_tail_and_per_frame_weights(cuda):
    m = olbs.synthetic_smpl(V=1000, seed=8)
    rs = np.random.RandomState(9)
    F = 6
    vp = rs.randn(F, 1000, 3).astype(np.float32)
    rot = olbs.batch_rodrigues((rs.randn(F * 24, 3) * 0.3).astype(np.float32)).reshape(F, 24, 3, 3)
    J = (rs.randn(F, 24, 3) * 0.3).astype(np.float32)
    _, A = olbs.batch_rigid_transform(rot, J, m["parents"])
    out = glbs.skin(_t(vp, cuda), _t(A, cuda), _t(m["lbs_weights"], cuda))
    assert np.abs(out.cpu().numpy() - olbs.skin(vp, A, m["lbs_weights"])).max() <= TOL
    Wf = rs.rand(F, 1000, 24).astype(np.float32)
    Wf /= Wf.sum(-1, keepdims=True)
    out = glbs.skin(_t(vp, cuda), _t(A, cuda), _t(Wf, cuda))
    want = np.stack([olbs.skin(vp[f:f + 1], A[f:f + 1], Wf[f])[0] for f in range(F)])
    assert np.abs(out.cpu().numpy() - want).max() <= TOL


def test_lbs_grad_path_matches_kernels(cuda):
    m = olbs.synthetic_smpl(V=500, seed=3)
    b, p = olbs.synthetic_frames(4, seed=4)
    mods = _model(m, cuda)
    v0, j0 = glbs.lbs(_t(b, cuda), _t(p, cuda), *mods)
    bt = _t(b, cuda).requires_grad_(True)
    v1, j1 = glbs.lbs(bt, _t(p, cuda), *mods)
    v1.sum().backward()
    assert bt.grad is not None and torch.isfinite(bt.grad).all()
    torch.testing.assert_close(v0, v1.detach(), rtol=0, atol=TOL)


def test_lbs_rejects_cpu_tensors():
    from garment4d_b200 import G4DError
    m = olbs.synthetic_smpl(V=50, seed=3)
    b, p = olbs.synthetic_frames(1)
    with pytest.raises(G4DError):
        glbs.lbs(torch.from_numpy(b), torch.from_numpy(p), *[torch.from_numpy(np.asarray(m[k])) for k in
                 ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights")])


def test_blend_shapes_kernel(cuda):
    """Stand-alone blend_shapes (lbs.py:288-309) on its own kernel vs the einsum it replaces (fp64 reference)."""
    rs = np.random.RandomState(5)
    for F, V, NB in ((1, 6890, 10), (7, 513, 10), (64, 100, 16)):
        betas = rs.randn(F, NB).astype(np.float32)
        sd = (rs.randn(V, 3, NB) * 0.01).astype(np.float32)
        got = glbs.blend_shapes(torch.from_numpy(betas).to(cuda), torch.from_numpy(sd).to(cuda))
        want = np.einsum("bl,mkl->bmk", betas.astype(np.float64), sd.astype(np.float64))
        assert got.shape == (F, V, 3) and np.abs(got.cpu().numpy() - want).max() <= 1e-6
    # the autograd route (einsum) gives the same values
    b = torch.from_numpy(betas).to(cuda).requires_grad_(True)
    assert torch.allclose(glbs.blend_shapes(b, torch.from_numpy(sd).to(cuda)), got, atol=1e-6)
