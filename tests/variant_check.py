"""Run in a subprocess by tests/test_env_variants_gpu.py with one tuning variable set (the C library reads most of them once per
process): a short parity pass over the kernels the variables steer.  Exit code 0 = every check passed."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pointnet2 as orc                                  # noqa: E402
from tests.util import clouds                                        # noqa: E402
from garment4d_b200.pointnet2 import pointnet2_modules as pm         # noqa: E402
from garment4d_b200.pointnet2 import pointnet2_utils as pu           # noqa: E402

dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def close(a, b, tol=2e-3):
    scale = float(b.abs().max())
    err = (a - b).abs()
    assert bool((err <= tol * scale + tol * b.abs()).all()), float(err.max())


# FPS + two-radius ball query: bit-exact against the CPU oracle (grid route, pruned FPS)
xyz = clouds(5, 3, 8192, "body")
x = t(xyz)
idx, new_xyz = pu.furthest_point_sample_and_gather(x, 160)
want = orc.furthest_point_sample(xyz, 160)
assert np.array_equal(idx.cpu().numpy(), want)
a, b = pu.ball_query_pair(0.1, 16, 0.2, 32, x, new_xyz)
assert np.array_equal(a.cpu().numpy(), orc.ball_query(0.1, 16, xyz, new_xyz.cpu().numpy()))
assert np.array_equal(b.cpu().numpy(), orc.ball_query(0.2, 32, xyz, new_xyz.cpu().numpy()))

# fused SA levels (xyz-only, 4 slots; features; the wide one-slot branch) against the operator route in true fp32
torch.manual_seed(0)
old = torch.backends.cudnn.allow_tf32
torch.backends.cudnn.allow_tf32 = False
for N, npoint, cin, radii, nsamples, mlps in ((4096, 512, 0, [0.1, 0.2], [16, 32], [[0, 16, 16, 32], [0, 32, 32, 64]]),
                                              (1024, 128, 96, [0.2, 0.3], [16, 32], [[96, 32, 32, 64], [96, 64, 64, 128]]),
                                              (256, 64, 192, [0.3, 0.5], [32, 64], [[192, 64, 64, 128], [192, 128, 128, 256]])):
    mod = pm.PointnetSAModuleMSG(npoint=npoint, radii=radii, nsamples=nsamples, mlps=[list(m) for m in mlps], bn=True).to(dev).eval()
    pts = t(clouds(7, 3, N, "body"))
    feats = torch.randn(3, cin, N, device=dev) if cin else None
    with torch.no_grad():
        nx, out = mod(pts, feats)
        assert pu.point_major_of(out) is not None
        mod.fused = False
        rx, ref = mod(pts, feats)
        mod.fused = True
    assert torch.equal(nx, rx)
    close(out, ref)

# coarser feature-propagation levels (tcgen05 two-layer MLP) against the operator route
for n, m, c2, c1, mlp in ((1024, 256, 256, 96, [352, 256, 128]), (256, 64, 384, 192, [576, 512, 256])):
    mod = pm.PointnetFPModule(mlp=list(mlp), bn=True).to(dev).eval()
    u = t(clouds(12, 2, n, "body"))
    k = u[:, :m].contiguous()
    skip, kf = torch.randn(2, c1, n, device=dev), torch.randn(2, c2, m, device=dev)
    with torch.no_grad():
        out = mod(u, k, skip, kf)
        mod.fused = False
        ref = mod(u, k, skip, kf)
        mod.fused = True
    close(out, ref, tol=5e-3)
torch.backends.cudnn.allow_tf32 = old

# training-mode backward through the operator route (G4D_BACKWARD)
f = torch.randn(2, 6, 512, device=dev, requires_grad=True)
pts = t(clouds(9, 2, 512, "body"))
g = pu.QueryAndGroup(0.3, 16)(pts, pts[:, :64].contiguous(), f)
g.sum().backward()
idxq = pu.ball_query(0.3, 16, pts, pts[:, :64].contiguous()).cpu().numpy()
cnt = np.zeros((2, 512), np.float32)
for bb in range(2):
    np.add.at(cnt[bb], idxq[bb].ravel(), 1.0)
assert np.array_equal(f.grad[:, 0].cpu().numpy(), cnt)
torch.cuda.synchronize()
print("variant ok:", {k: v for k, v in os.environ.items() if k.startswith("G4D_")})
