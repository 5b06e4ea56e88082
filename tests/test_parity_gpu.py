"""-m gpu parity tests: the CUDA path (through the C ABI / the pointnet2_cuda mirror) against
  (1) the CPU oracle (oracle/pointnet2_oracle.c), bit-exact for indices, and
  (2) the reference's own kernels compiled unmodified (oracle/_ref), when that library travelled with the repo.
"""
import numpy as np
import pytest
import torch

from oracle import pointnet2 as orc
from oracle import refgpu
from tests.util import clouds

pytestmark = pytest.mark.gpu

from garment4d_b200.pointnet2 import pointnet2_utils as pu   # noqa: E402

CASES = [  # (seed, B, N, m, radius, K)
    (1, 2, 1024, 256, 0.2, 32),      # BASELINE config 1
    (2, 3, 1000, 200, 0.15, 16),     # bs = 512 < N
    (3, 2, 37, 9, 0.5, 8),           # tiny
    (4, 1, 300, 300, 0.3, 64),       # m == N
    (5, 2, 2048, 512, 0.1, 16),
    (6, 2, 4096, 256, 0.07, 32),
    (7, 2, 8192, 1024, 0.1, 32),     # SA1b
    (8, 1, 6890, 1024, 0.05, 16),    # the reference's real N
    (9, 1, 16384, 1024, 0.05, 16),   # config 5 cloud size
    (10, 1, 20000, 128, 0.05, 16),   # generic path (N > 16384)
]


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("kind", ["cube", "body"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"N{c[2]}m{c[3]}")
def test_fps_ball_group_vs_oracle(cuda, case, kind):
    seed, B, N, m, radius, K = case
    xyz = clouds(seed, B, N, kind)
    x = _t(xyz, cuda)
    # FPS: drop-in op, fused op, oracle, reference kernels
    idx = pu.furthest_point_sample(x, m)
    idx_f, new_xyz_f = pu.furthest_point_sample_and_gather(x, m)
    want = orc.furthest_point_sample(xyz, m)
    assert np.array_equal(idx.cpu().numpy(), want), "FPS indices differ from the oracle"
    assert torch.equal(idx, idx_f)
    new_xyz = pu.gather_operation(x.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    assert torch.equal(new_xyz, new_xyz_f)
    assert np.array_equal(new_xyz.cpu().numpy(), orc.gather_operation(xyz.transpose(0, 2, 1), want).transpose(0, 2, 1))
    if refgpu.available():
        assert torch.equal(idx, refgpu.furthest_point_sample(x, m)), "FPS differs from the reference kernel"
    # ball query (+ the two-scale scan)
    bq = pu.ball_query(radius, K, x, new_xyz)
    want_bq = orc.ball_query(radius, K, xyz, new_xyz.cpu().numpy())
    assert np.array_equal(bq.cpu().numpy(), want_bq), "ball_query differs from the oracle"
    bq0, bq1 = pu.ball_query_pair(radius, K, radius * 0.5, max(K // 2, 1), x, new_xyz)
    assert torch.equal(bq0, bq)
    assert np.array_equal(bq1.cpu().numpy(), orc.ball_query(radius * 0.5, max(K // 2, 1), xyz, new_xyz.cpu().numpy()))
    if refgpu.available():
        assert torch.equal(bq, refgpu.ball_query(radius, K, x, new_xyz)), "ball_query differs from the reference kernel"
    # group + fused QueryAndGroup
    C = 5
    feats = np.random.RandomState(seed + 100).randn(B, C, N).astype(np.float32)
    f = _t(feats, cuda)
    g = pu.grouping_operation(f, bq)
    assert np.array_equal(g.cpu().numpy(), orc.grouping_operation(feats, want_bq))
    if K in (4, 8, 16, 32, 64, 128):
        qg = pu.QueryAndGroup(radius, K)(x, new_xyz, f)
        want_qg = orc.query_and_group(radius, K, xyz, new_xyz.cpu().numpy(), feats)
        assert np.array_equal(qg.cpu().numpy(), want_qg), "fused QueryAndGroup differs from the oracle"
        qg0 = pu.QueryAndGroup(radius, K)(x, new_xyz, None)
        assert np.array_equal(qg0.cpu().numpy(), want_qg[:, :3])


def test_ball_query_no_hit_rows_stay_zero(cuda):
    xyz = clouds(11, 2, 512, "cube")
    q = xyz[:, :16].copy() + np.float32(10.0)          # far away: no hits
    idx = pu.ball_query(0.05, 8, _t(xyz, cuda), _t(q, cuda))
    assert int(idx.abs().sum()) == 0
    assert np.array_equal(idx.cpu().numpy(), orc.ball_query(0.05, 8, xyz, q))


@pytest.mark.parametrize("case", [(21, 2, 1024, 256), (22, 2, 8192, 1024), (23, 1, 300, 2), (24, 2, 256, 64), (25, 1, 1000, 1500)],
                         ids=lambda c: f"n{c[2]}m{c[3]}")
def test_three_nn_interpolate_vs_oracle(cuda, case):
    seed, B, n, m = case
    unknown = clouds(seed, B, n, "body", dup_frac=0.02)
    known = clouds(seed + 1, B, m, "body", dup_frac=0.0) if m > n else unknown[:, :m].copy()
    u, k = _t(unknown, cuda), _t(known, cuda)
    dist, idx = pu.three_nn(u, k)
    wd, wi = orc.three_nn(unknown, known)
    assert np.array_equal(idx.cpu().numpy(), wi)
    assert np.array_equal(dist.cpu().numpy(), wd)
    if refgpu.available():
        rd, ri = refgpu.three_nn(u, k)
        assert torch.equal(idx, ri) and torch.equal(dist, rd)
    rs = np.random.RandomState(seed + 2)
    feats = rs.randn(B, 7, m).astype(np.float32)
    w = rs.rand(B, n, 3).astype(np.float32)
    out = pu.three_interpolate(_t(feats, cuda), idx, _t(w, cuda))
    assert np.array_equal(out.cpu().numpy(), orc.three_interpolate(feats, wi, w))
    if refgpu.available():
        assert torch.equal(out, refgpu.three_interpolate(_t(feats, cuda), idx, _t(w, cuda)))


def test_backward_ops_vs_oracle(cuda):
    rs = np.random.RandomState(31)
    B, C, N, P, S = 2, 6, 500, 64, 8
    idx = rs.randint(0, N, (B, P, S)).astype(np.int32)
    feats = torch.from_numpy(rs.randn(B, C, N).astype(np.float32)).to(cuda).requires_grad_(True)
    go = rs.randn(B, C, P, S).astype(np.float32)
    out = pu.grouping_operation(feats, _t(idx, cuda))
    out.backward(_t(go, cuda))
    np.testing.assert_allclose(feats.grad.cpu().numpy(), orc.grouping_operation_grad(go, idx, N), rtol=1e-5, atol=1e-5)
    # gather
    gidx = rs.randint(0, N, (B, P)).astype(np.int32)
    feats2 = torch.from_numpy(rs.randn(B, C, N).astype(np.float32)).to(cuda).requires_grad_(True)
    go2 = rs.randn(B, C, P).astype(np.float32)
    pu.gather_operation(feats2, _t(gidx, cuda)).backward(_t(go2, cuda))
    np.testing.assert_allclose(feats2.grad.cpu().numpy(), orc.gather_operation_grad(go2, gidx, N), rtol=1e-5, atol=1e-5)
    # three_interpolate
    m, n = 40, 300
    iidx = rs.randint(0, m, (B, n, 3)).astype(np.int32)
    w = rs.rand(B, n, 3).astype(np.float32)
    feats3 = torch.from_numpy(rs.randn(B, C, m).astype(np.float32)).to(cuda).requires_grad_(True)
    go3 = rs.randn(B, C, n).astype(np.float32)
    pu.three_interpolate(feats3, _t(iidx, cuda), _t(w, cuda)).backward(_t(go3, cuda))
    np.testing.assert_allclose(feats3.grad.cpu().numpy(), orc.three_interpolate_grad(go3, iidx, w, m), rtol=1e-5, atol=1e-5)


def test_deterministic_backward_is_bit_exact_and_reproducible(cuda):
    """The segmented-reduction backward (scatter_det.cu) adds every destination's contributions in ascending source order:
    bit-identical to the serial loop of the oracle (the reference's atomicAdd scatter has no defined order), run after run,
    including the heavy-hitter case (all-zero idx rows of empty balls pile thousands of sources onto point 0)."""
    assert pu.BACKWARD == "det"
    rs = np.random.RandomState(5)
    B, C, N, P, S = 3, 7, 2000, 700, 32
    idx = rs.randint(0, N, (B, P, S)).astype(np.int32)
    idx[:, 100:400] = 0                                          # 9600 sources on point 0 of every cloud
    go = rs.randn(B, C, P, S).astype(np.float32)
    want = orc.grouping_operation_grad(go, idx, N)
    outs = []
    for _ in range(3):
        feats = torch.zeros(B, C, N, device=cuda, requires_grad=True)
        pu.grouping_operation(feats, _t(idx, cuda)).backward(_t(go, cuda))
        outs.append(feats.grad.clone())
    assert np.array_equal(outs[0].cpu().numpy(), want)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    # gather and three_interpolate through the same primitive
    gidx = rs.randint(0, N, (B, P)).astype(np.int32)
    go2 = rs.randn(B, C, P).astype(np.float32)
    f2 = torch.zeros(B, C, N, device=cuda, requires_grad=True)
    pu.gather_operation(f2, _t(gidx, cuda)).backward(_t(go2, cuda))
    assert np.array_equal(f2.grad.cpu().numpy(), orc.gather_operation_grad(go2, gidx, N))
    m, n = 300, 5000
    iidx = rs.randint(0, m, (B, n, 3)).astype(np.int32)
    w = rs.rand(B, n, 3).astype(np.float32)
    go3 = rs.randn(B, C, n).astype(np.float32)
    f3 = torch.zeros(B, C, m, device=cuda, requires_grad=True)
    pu.three_interpolate(f3, _t(iidx, cuda), _t(w, cuda)).backward(_t(go3, cuda))
    assert np.array_equal(f3.grad.cpu().numpy(), orc.three_interpolate_grad(go3, iidx, w, m))
    # the 1:1 atomic entry points (the reference's kernels' shape) agree up to summation order
    prev, pu.BACKWARD = pu.BACKWARD, "atomic"
    try:
        fa = torch.zeros(B, C, N, device=cuda, requires_grad=True)
        pu.grouping_operation(fa, _t(idx, cuda)).backward(_t(go, cuda))
    finally:
        pu.BACKWARD = prev
    np.testing.assert_allclose(fa.grad.cpu().numpy(), want, rtol=1e-4, atol=1e-3)


def test_query_and_group_backward_matches_composition(cuda):
    xyz = clouds(41, 2, 600, "cube")
    x = _t(xyz, cuda)
    idx, new_xyz = pu.furthest_point_sample_and_gather(x, 50)
    feats = torch.randn(2, 4, 600, device=cuda)
    go = torch.randn(2, 7, 50, 16, device=cuda)
    grads = []
    for fused in (True, False):
        xa = x.clone().requires_grad_(True)
        na = new_xyz.clone().requires_grad_(True)
        fa = feats.clone().requires_grad_(True)
        if fused:
            out = pu.QueryAndGroup(0.2, 16)(xa, na, fa)
        else:
            bq = pu.ball_query(0.2, 16, xa, na)
            gx = pu.grouping_operation(xa.transpose(1, 2).contiguous(), bq) - na.transpose(1, 2).unsqueeze(-1)
            out = torch.cat([gx, pu.grouping_operation(fa, bq)], dim=1)
        out.backward(go)
        grads.append((out.detach(), xa.grad, na.grad, fa.grad))
    for a, b in zip(*grads):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)


def test_full_size_properties(cuda):
    """BASELINE full size (N=8192 -> 1024): size-independent properties instead of the (slow) oracle."""
    B, N, m = 16, 8192, 1024
    x = _t(clouds(51, B, N, "body"), cuda)
    idx, new_xyz = pu.furthest_point_sample_and_gather(x, m)
    i = idx.long()
    assert int(i[:, 0].abs().max()) == 0
    # duplicates (exact copies of earlier points) can only be picked once every distinct point is taken: not at m << N
    d = torch.cdist(new_xyz, new_xyz)
    d = d + torch.eye(m, device=cuda)[None] * 10
    assert float(d.min()) > 0, "FPS picked two coincident points"
    # FPS is greedy: the min-distance-to-chosen-set of each newly chosen point never increases
    steps = []
    for b in range(2):
        dist = torch.full((N,), 1e10, device=cuda)
        prev = None
        for j in range(m - 1):
            dist = torch.minimum(dist, ((x[b] - x[b, i[b, j]]) ** 2).sum(-1))
            cur = float(dist[i[b, j + 1]])
            assert abs(cur - float(dist.max())) <= 1e-6 * max(cur, 1e-12) + 1e-12
            if prev is not None:
                assert cur <= prev * (1 + 1e-5)
            prev = cur
    bq = pu.ball_query(0.1, 32, x, new_xyz)
    g = torch.gather(x, 1, bq.long().reshape(B, -1, 1).expand(-1, -1, 3)).reshape(B, m, 32, 3)
    d2 = ((g - new_xyz[:, :, None]) ** 2).sum(-1)
    assert float(d2.max()) < 0.1 ** 2 * (1 + 1e-5)
    # ascending index order up to the padding
    first = bq[:, :, :1]
    inc = (bq[:, :, 1:] > bq[:, :, :-1]) | (bq[:, :, 1:] == first)
    assert bool(inc.all())


def _brute_ball(radius, K, x, new_xyz):
    from garment4d_b200 import pointnet2_cuda
    B, N, _ = x.shape
    idx = torch.zeros(B, new_xyz.shape[1], K, dtype=torch.int32, device=x.device)
    pointnet2_cuda.ball_query_wrapper(B, N, new_xyz.shape[1], radius, K, new_xyz, x, idx)
    return idx


def _brute_nn(u, k):
    from garment4d_b200 import pointnet2_cuda
    B, n, _ = u.shape
    d2 = torch.empty(B, n, 3, device=u.device)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=u.device)
    pointnet2_cuda.three_nn_wrapper(B, n, k.shape[1], u, k, d2, idx)
    return d2, idx


@pytest.mark.parametrize("shape", ["body", "cube", "coincident", "outlier", "line", "planar_dups", "far_origin", "far_origin_big", "needle"])
def test_grid_searches_equal_brute_force(cuda, shape):
    """The uniform-grid ball query / three_nn must reproduce the brute-force kernels bit for bit on awkward clouds."""
    B, N, m = 2, 4096, 600
    rs = np.random.RandomState(77)
    if shape in ("body", "cube"):
        xyz = clouds(61, B, N, shape, dup_frac=0.1)
    elif shape == "coincident":
        xyz = np.full((B, N, 3), 0.25, np.float32)
    elif shape == "outlier":
        xyz = clouds(62, B, N, "body")
        xyz[:, 7] = 1e6
        xyz[:, 9] = -3e5
    elif shape == "line":
        xyz = np.zeros((B, N, 3), np.float32)
        xyz[..., 0] = rs.rand(B, N).astype(np.float32)
    elif shape == "far_origin":
        # far from the origin relative to the radii (|coord| ~ 1e3 = 20000 r for r = 0.05): coordinates are coarse (ulp 6e-5),
        # cell faces and binning round -- the grid searches must still see every hit the brute-force scan sees
        xyz = (clouds(63, B, N, "body", dup_frac=0.1) + np.array([1000.0, -2000.0, 512.0], np.float32)).astype(np.float32)
    elif shape == "far_origin_big":
        xyz = (clouds(64, B, N, "cube", dup_frac=0.1) + np.array([65536.0, 0.0, -30000.0], np.float32)).astype(np.float32)
    elif shape == "needle":
        # extremely elongated: thousands of cells along one axis would be needed for cell edge = radius
        xyz = (rs.rand(B, N, 3) * np.array([400.0, 0.05, 0.05])).astype(np.float32)
    else:
        xyz = np.zeros((B, N, 3), np.float32)
        xyz[..., :2] = (rs.randint(0, 40, (B, N, 2)) * 0.025).astype(np.float32)     # lattice: many exact distance ties
    x = _t(xyz, cuda)
    _, new_xyz = pu.furthest_point_sample_and_gather(x, m)
    for radius, K in ((0.05, 16), (0.1, 32), (0.3, 64), (5.0, 128)):
        got = pu.ball_query(radius, K, x, new_xyz)
        assert torch.equal(got, _brute_ball(radius, K, x, new_xyz)), f"grid ball query differs (r={radius})"
    a, b = pu.ball_query_pair(0.05, 16, 0.1, 32, x, new_xyz)
    assert torch.equal(a, _brute_ball(0.05, 16, x, new_xyz)) and torch.equal(b, _brute_ball(0.1, 32, x, new_xyz))
    # three_nn: unknown = whole cloud (grid order available from the ball query above), known = the 600 centroids
    d, i = pu.three_nn(x, new_xyz)
    d2b, ib = _brute_nn(x, new_xyz)
    assert torch.equal(i, ib), "grid three_nn indices differ"
    assert torch.equal(d, torch.sqrt(d2b))
    # queries outside the cloud's bounding box
    far = new_xyz + 0.07
    assert torch.equal(pu.ball_query(0.1, 32, x, far), _brute_ball(0.1, 32, x, far))
    d, i = pu.three_nn((x + 0.3).contiguous(), new_xyz)
    d2b, ib = _brute_nn((x + 0.3).contiguous(), new_xyz)
    assert torch.equal(i, ib) and torch.equal(d, torch.sqrt(d2b))


@pytest.mark.parametrize("case", [(41, 2, 1024, 256, 256, 96), (42, 3, 256, 64, 384, 192), (43, 2, 1000, 77, 20, 0), (44, 1, 300, 5, 3, 7)],
                         ids=lambda c: f"n{c[2]}m{c[3]}c{c[4]}+{c[5]}")
def test_fp_interp_concat_vs_operator_sequence(cuda, case):
    """g4d_fp_interp_concat == the reference's operator sequence of PointnetFPModule.forward (pointnet2_modules.py:138-152):
    three_nn -> 1/(dist+1e-8) -> normalise -> three_interpolate -> cat, run here through torch + the 1:1 operators."""
    import ctypes
    from garment4d_b200 import _lib
    seed, B, n, m, c2, c1 = case
    unknown = clouds(seed, B, n, "body", dup_frac=0.02)       # duplicates: zero distances exercise the 1e-8 term
    known = unknown[:, :m].copy()
    rs = np.random.RandomState(seed)
    u, k = _t(unknown, cuda), _t(known, cuda)
    kf = _t(rs.randn(B, c2, m).astype(np.float32), cuda)
    skip = _t(rs.randn(B, c1, n).astype(np.float32), cuda) if c1 else None
    dist, idx = pu.three_nn(u, k)
    dist_recip = 1.0 / (dist + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
    want = pu.three_interpolate(kf, idx, weight)
    if skip is not None:
        want = torch.cat([want, skip], dim=1)
    dist2 = torch.empty(B, n, 3, dtype=torch.float32, device=cuda)
    idx2 = torch.empty(B, n, 3, dtype=torch.int32, device=cuda)
    pu.three_nn_raw(u, k, dist2, idx2)
    assert torch.equal(idx2, idx)
    out = torch.full((B, c2 + c1, n), float("nan"), dtype=torch.float32, device=cuda)
    rc = _lib.lib().g4d_fp_interp_concat(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx2), _lib.ptr(kf), _lib.ptr(skip), _lib.ptr(out),
                                         _lib.stream_ptr())
    _lib.check(rc, "g4d_fp_interp_concat")
    assert torch.isfinite(out).all()
    if c1:
        assert torch.equal(out[:, c2:], skip), "skip channels must be copied verbatim"
    # weights: same correctly rounded operations; torch's 3-term sum may associate differently -> 2 ulp of slack on the output
    np.testing.assert_allclose(out[:, :c2].cpu().numpy(), want[:, :c2].cpu().numpy(), rtol=5e-7, atol=1e-7 * float(kf.abs().max()))


def test_bias_relu_pm(cuda):
    """g4d_bias_relu_pm: in-place bias+ReLU (== g4d_bias_relu_inplace) plus the fp16 point-major copy."""
    from garment4d_b200 import _lib
    rs = np.random.RandomState(51)
    for (B, C, n) in [(2, 128, 1024), (3, 50, 77), (1, 7, 5)]:
        y0 = _t(rs.randn(B, C, n).astype(np.float32) * 3, cuda)
        b = _t(rs.randn(C).astype(np.float32), cuda)
        y = y0.clone()
        pm = torch.zeros(B, n, C, dtype=torch.float16, device=cuda)
        rc = _lib.lib().g4d_bias_relu_pm(B, C, n, _lib.ptr(y), _lib.ptr(b), 1, _lib.ptr(pm), _lib.stream_ptr())
        _lib.check(rc, "g4d_bias_relu_pm")
        want = torch.relu(y0 + b[None, :, None])
        assert torch.equal(y, want)
        assert torch.equal(pm, want.transpose(1, 2).to(torch.float16).contiguous())


def test_fp_batched_gemm_route_entry_points(cuda):
    """g4d_fp_interp_concat_cbn_h / g4d_bias_relu_h / g4d_bias_relu_unpack against the fp32 entry points and torch."""
    from garment4d_b200 import _lib
    L = _lib.lib()
    rs = np.random.RandomState(61)
    B, n, m, c2, c1 = 3, 256, 64, 40, 24
    unknown = clouds(61, B, n, "body", dup_frac=0.02)
    u, k = _t(unknown, cuda), _t(unknown[:, :m].copy(), cuda)
    kf = _t(rs.randn(B, c2, m).astype(np.float32), cuda)
    skip = _t(rs.randn(B, c1, n).astype(np.float32), cuda)
    dist2 = torch.empty(B, n, 3, dtype=torch.float32, device=cuda)
    idx = torch.empty(B, n, 3, dtype=torch.int32, device=cuda)
    pu.three_nn_raw(u, k, dist2, idx)
    ref = torch.empty(B, c2 + c1, n, dtype=torch.float32, device=cuda)
    _lib.check(L.g4d_fp_interp_concat(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx), _lib.ptr(kf), _lib.ptr(skip), _lib.ptr(ref),
                                      _lib.stream_ptr()), "g4d_fp_interp_concat")
    xh = torch.zeros(c2 + c1, B * n, dtype=torch.float16, device=cuda)
    _lib.check(L.g4d_fp_interp_concat_cbn_h(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx), _lib.ptr(kf), _lib.ptr(skip), _lib.ptr(xh),
                                            _lib.stream_ptr()), "g4d_fp_interp_concat_cbn_h")
    assert torch.equal(xh.view(c2 + c1, B, n), ref.permute(1, 0, 2).to(torch.float16))
    # point-major fp16 source: identical to the channel-major kernel run on the fp16-rounded features
    kpm = kf.transpose(1, 2).to(torch.float16).contiguous()
    kf_r = kpm.float().transpose(1, 2).contiguous()
    x_cm = torch.zeros(c2 + c1, B * n, dtype=torch.float16, device=cuda)
    _lib.check(L.g4d_fp_interp_concat_cbn_h(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx), _lib.ptr(kf_r), _lib.ptr(skip), _lib.ptr(x_cm),
                                            _lib.stream_ptr()), "g4d_fp_interp_concat_cbn_h")
    x_pm = torch.zeros_like(x_cm)
    _lib.check(L.g4d_fp_interp_concat_pm_cbn_h(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx), _lib.ptr(kpm), _lib.ptr(skip), _lib.ptr(x_pm),
                                               _lib.stream_ptr()), "g4d_fp_interp_concat_pm_cbn_h")
    assert torch.equal(x_pm, x_cm)
    # point-major rows route: same values as the (C, B*n) operand, transposed; then its two epilogues
    spm = skip.transpose(1, 2).to(torch.float16).contiguous()
    x_rows = torch.zeros(B * n, c2 + c1, dtype=torch.float16, device=cuda)
    _lib.check(L.g4d_fp_interp_concat_rows_h(B, c2, c1, m, n, _lib.ptr(dist2), _lib.ptr(idx), _lib.ptr(kpm), _lib.ptr(spm), _lib.ptr(x_rows),
                                             _lib.stream_ptr()), "g4d_fp_interp_concat_rows_h")
    assert torch.equal(x_rows.t().contiguous(), x_cm)
    for cw in (24, 64, 256):          # 24: generic path (modulo per element); 64, 256: biases held in registers
        yr0 = _t((rs.randn(B * n, cw) * 3).astype(np.float16), cuda)
        bw = _t(rs.randn(cw).astype(np.float32), cuda)
        yr = yr0.clone()
        _lib.check(L.g4d_bias_relu_rows_h(B * n, cw, _lib.ptr(yr), _lib.ptr(bw), 1, _lib.stream_ptr()), "g4d_bias_relu_rows_h")
        assert torch.equal(yr, torch.relu(yr0.float() + bw[None, :]).to(torch.float16))
    for (Bq, Cq, nq) in [(2, 128, 1024), (3, 50, 77)]:
        yin = _t((rs.randn(Bq, nq, Cq) * 3).astype(np.float32), cuda)
        bq = _t(rs.randn(Cq).astype(np.float32), cuda)
        out = torch.zeros(Bq, Cq, nq, dtype=torch.float32, device=cuda)
        pmq = torch.zeros(Bq, nq, Cq, dtype=torch.float16, device=cuda)
        _lib.check(L.g4d_bias_relu_rows_unpack(Bq, Cq, nq, _lib.ptr(yin), _lib.ptr(bq), 1, _lib.ptr(out), _lib.ptr(pmq), _lib.stream_ptr()),
                   "g4d_bias_relu_rows_unpack")
        want = torch.relu(yin + bq[None, None, :])
        assert torch.equal(out, want.transpose(1, 2).contiguous())
        assert torch.equal(pmq, want.to(torch.float16))
    # bias + ReLU on (c, len) fp16
    C, ln = 37, 8 * 123
    y0 = _t((rs.randn(C, ln) * 3).astype(np.float16), cuda)
    b = _t(rs.randn(C).astype(np.float32), cuda)
    y = y0.clone()
    _lib.check(L.g4d_bias_relu_h(C, ln, _lib.ptr(y), _lib.ptr(b), 1, _lib.stream_ptr()), "g4d_bias_relu_h")
    assert torch.equal(y, torch.relu(y0.float() + b[:, None]).to(torch.float16))
    # last-layer epilogue: (c, b, n) -> (b, c, n) fp32 (+ fp16 point-major)
    for in_half in (0, 1):
        for (Bq, Cq, nq) in [(2, 128, 1024), (3, 50, 77)]:
            yin = _t((rs.randn(Cq, Bq, nq) * 3).astype(np.float16 if in_half else np.float32), cuda)
            bq = _t(rs.randn(Cq).astype(np.float32), cuda)
            out = torch.zeros(Bq, Cq, nq, dtype=torch.float32, device=cuda)
            pm = torch.zeros(Bq, nq, Cq, dtype=torch.float16, device=cuda)
            _lib.check(L.g4d_bias_relu_unpack(Bq, Cq, nq, _lib.ptr(yin), in_half, _lib.ptr(bq), 1, _lib.ptr(out), _lib.ptr(pm),
                                              _lib.stream_ptr()), "g4d_bias_relu_unpack")
            want = torch.relu(yin.float() + bq[:, None, None]).permute(1, 0, 2).contiguous()
            assert torch.equal(out, want)
            assert torch.equal(pm, want.transpose(1, 2).to(torch.float16).contiguous())
            out2 = torch.zeros_like(out)
            _lib.check(L.g4d_bias_relu_unpack(Bq, Cq, nq, _lib.ptr(yin), in_half, _lib.ptr(bq), 1, _lib.ptr(out2), None,
                                              _lib.stream_ptr()), "g4d_bias_relu_unpack")
            assert torch.equal(out2, want)


def test_group_all_equals_the_reference_views(cuda):
    """GroupAll.forward (pointnet2_utils.py:273-291): (B, 3+C, 1, N) = cat(xyz^T, features) -- no kernel, torch views."""
    rs = np.random.RandomState(3)
    xyz = rs.rand(2, 77, 3).astype(np.float32)
    feats = rs.randn(2, 5, 77).astype(np.float32)
    x, f = _t(xyz, cuda), _t(feats, cuda)
    out = pu.GroupAll(use_xyz=True)(x, None, f)
    want = np.concatenate([xyz.transpose(0, 2, 1)[:, :, None, :], feats[:, :, None, :]], axis=1)
    assert out.shape == (2, 8, 1, 77) and np.array_equal(out.cpu().numpy(), want)
    assert np.array_equal(pu.GroupAll(use_xyz=False)(x, None, f).cpu().numpy(), feats[:, :, None, :])
    assert np.array_equal(pu.GroupAll()(x, None, None).cpu().numpy(), xyz.transpose(0, 2, 1)[:, :, None, :])
    # and through a set-abstraction module with npoint=None (PointnetSAModule, pointnet2_modules.py:94-113): operator route
    from garment4d_b200.pointnet2 import pointnet2_modules as pm
    torch.manual_seed(0)
    sa = pm.PointnetSAModule(mlp=[5, 16, 32], use_xyz=True).to(cuda).eval()
    with torch.no_grad():
        new_xyz, y = sa(x, f)
    assert new_xyz is None and y.shape == (2, 32, 1)
    with torch.no_grad():
        ref = torch.nn.functional.max_pool2d(sa.mlps[0](torch.from_numpy(want).to(cuda)), kernel_size=[1, 77]).squeeze(-1)
    assert torch.allclose(y, ref)


@pytest.mark.parametrize("case", [(16, 4, 50), (17, 8, 33), (70, 16, 100), (96, 32, 256), (195, 64, 64), (130, 12, 7)], ids=lambda c: f"C{c[0]}K{c[1]}m{c[2]}")
def test_query_and_group_from_point_major_rows(cuda, case):
    """QueryAndGroup at feature levels (C >= 16) groups from a point-major copy of the features (g4d_group_fused_pm): channel counts
    that are not multiples of 4 or 64, planes (m * K) that are not multiples of the 64-position tile, with and without xyz; and the
    copy is refreshed when the features change in place."""
    C, K, m = case
    B, N, radius = 2, 700, 0.3
    xyz = clouds(C + K, B, N, "body")
    x = _t(xyz, cuda)
    new_xyz = x[:, :m].contiguous()
    feats = np.random.RandomState(C).randn(B, C, N).astype(np.float32)
    f = _t(feats, cuda)
    for use_xyz in (True, False):
        got = pu.QueryAndGroup(radius, K, use_xyz=use_xyz)(x, new_xyz, f)
        want = orc.query_and_group(radius, K, xyz, xyz[:, :m], feats, use_xyz=use_xyz)
        assert np.array_equal(got.cpu().numpy(), want)
    f.mul_(2.0)                                                   # in place: the remembered point-major copy is stale now
    got = pu.QueryAndGroup(radius, K)(x, new_xyz, f)
    assert np.array_equal(got.cpu().numpy(), orc.query_and_group(radius, K, xyz, xyz[:, :m], feats * np.float32(2.0), use_xyz=True))


@pytest.mark.parametrize("kind", ["cube", "body"])
def test_fps_both_kernel_shapes_give_the_reference_indices(cuda, kind):
    """The pruned FPS has two shapes (1024 threads x 8 points when every cloud of the step gets an SM, 512 x 16 otherwise); the
    caller's concurrency hint picks one, the indices never change."""
    from garment4d_b200 import _lib
    xyz = clouds(31, 3, 8192, kind)
    x = _t(xyz, cuda)
    want = orc.furthest_point_sample(xyz, 300)
    L = _lib.lib()
    try:
        for hint in (1, 100000):                                   # few clouds -> wide shape; many -> two clouds per SM
            L.g4d_fps_concurrency_hint(hint)
            idx, new_xyz = pu.furthest_point_sample_and_gather(x, 300)
            assert np.array_equal(idx.cpu().numpy(), want), f"hint {hint}"
            assert np.array_equal(new_xyz.cpu().numpy(), np.take_along_axis(xyz, want[:, :, None].astype(np.int64).repeat(3, 2), axis=1))
    finally:
        L.g4d_fps_concurrency_hint(0)
