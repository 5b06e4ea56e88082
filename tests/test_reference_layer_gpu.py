"""-m gpu: the drop-in boundary, shown with the reference's OWN Python operator layer.

oracle/build_ref.sh packs the reference's unmodified ``pointnet2_utils.py`` / ``pointnet2_modules.py`` / ``pytorch_utils.py``
into oracle/_ref/pointnet2_reference_layer.zip (git-ignored test infrastructure, like the reference kernels next to it; the GPU
box has no /root/reference); the archive goes on sys.path (zipimport).  They do ``import pointnet2_cuda as pointnet2`` (pointnet2_utils.py:7): with the repository root on sys.path
that resolves to the repository's top-level ``pointnet2_cuda.py``, i.e. the B200 kernels behind the reference's nine-function
extension API.  Results are compared with the CPU oracle (bit-exact for indices) and with this repository's own operator layer.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import pointnet2 as orc
from tests.util import clouds

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, "oracle", "_ref", "pointnet2_reference_layer.zip")


@pytest.fixture(scope="module")
def ref_layer():
    if not os.path.exists(REF_PKG):
        pytest.skip("oracle/_ref/pointnet2_reference_layer.zip was not built (oracle/build_ref.sh needs /root/reference)")
    sys.path.insert(0, REF_PKG)
    try:
        import pointnet2_cuda                                     # the repository's shim, found through ROOT on sys.path
        assert os.path.dirname(os.path.abspath(pointnet2_cuda.__file__)) == ROOT
        import pointnet2.pointnet2_utils as ref_pu                # the reference's file, unmodified
        import pointnet2.pointnet2_modules as ref_pm
        assert os.path.abspath(ref_pu.__file__).startswith(REF_PKG)
        yield ref_pu, ref_pm
    finally:
        sys.path.remove(REF_PKG)


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("case", [(1, 2, 1024, 256, 0.2, 32), (7, 2, 8192, 1024, 0.1, 32), (8, 1, 6890, 512, 0.05, 16)], ids=lambda c: f"N{c[2]}")
def test_reference_operators_on_the_b200_kernels(cuda, ref_layer, case):
    ref_pu, _ = ref_layer
    seed, B, N, m, radius, K = case
    xyz = clouds(seed, B, N, "body")
    x = _t(xyz, cuda)
    with torch.cuda.device(cuda):                                 # the reference allocates with torch.cuda.IntTensor(...): current device
        idx = ref_pu.furthest_point_sample(x, m)
        want = orc.furthest_point_sample(xyz, m)
        assert idx.dtype == torch.int32 and np.array_equal(idx.cpu().numpy(), want)
        new_xyz = ref_pu.gather_operation(x.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        assert np.array_equal(new_xyz.cpu().numpy(), orc.gather_operation(xyz.transpose(0, 2, 1), want).transpose(0, 2, 1))
        bq = ref_pu.ball_query(radius, K, x, new_xyz)
        assert np.array_equal(bq.cpu().numpy(), orc.ball_query(radius, K, xyz, new_xyz.cpu().numpy()))
        rs = np.random.RandomState(seed)
        feats = rs.randn(B, 8, N).astype(np.float32)
        f = _t(feats, cuda)
        grouped = ref_pu.QueryAndGroup(radius, K)(x, new_xyz, f)   # ball_query -> grouping x2 -> subtract -> cat, the reference's code
        want_g = orc.query_and_group(radius, K, xyz, new_xyz.cpu().numpy(), feats, use_xyz=True)
        assert np.array_equal(grouped.cpu().numpy(), want_g)
        dist, i3 = ref_pu.three_nn(x, new_xyz)
        wd, wi = orc.three_nn(xyz, new_xyz.cpu().numpy())
        assert np.array_equal(i3.cpu().numpy(), wi) and np.array_equal(dist.cpu().numpy(), wd)
        w = torch.rand(B, N, 3, device=cuda)
        w = w / w.sum(2, keepdim=True)
        kf = _t(rs.randn(B, 8, m).astype(np.float32), cuda)
        out = ref_pu.three_interpolate(kf, i3, w)
        assert np.array_equal(out.cpu().numpy(), orc.three_interpolate(kf.cpu().numpy(), wi, w.cpu().numpy()))


def test_reference_sa_module_equals_ours(cuda, ref_layer):
    """The reference's PointnetSAModuleMSG (its forward, its SharedMLP) on our kernels == our module's operator route, same weights."""
    _, ref_pm = ref_layer
    from garment4d_b200.pointnet2 import pointnet2_modules as pm
    torch.manual_seed(5)
    kw = dict(npoint=128, radii=[0.1, 0.2], nsamples=[16, 32])
    ours = pm.PointnetSAModuleMSG(mlps=[[6, 16, 16, 32], [6, 32, 32, 64]], **kw).to(cuda).eval()
    theirs = ref_pm.PointnetSAModuleMSG(mlps=[[6, 16, 16, 32], [6, 32, 32, 64]], **kw).to(cuda).eval()
    theirs.load_state_dict(ours.state_dict())                     # same parameter names: the checkpoint contract
    xyz = _t(clouds(11, 2, 2048, "body"), cuda)
    feats = torch.randn(2, 6, 2048, device=cuda)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad(), torch.cuda.device(cuda):
            ours.fused = False
            a_xyz, a = ours(xyz, feats)
            b_xyz, b = theirs(xyz, feats)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert torch.equal(a_xyz, b_xyz)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-5)
