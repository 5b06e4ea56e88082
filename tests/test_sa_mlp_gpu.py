"""-m gpu: the tcgen05 grouped-MLP + max kernel and the fused set-abstraction route against a plain PyTorch fp32
reference of the same op (conv1x1 -> eval BN -> ReLU x3 -> max over the neighbourhood).

Tolerance: operands are fp16 (11-bit significand, the same class as the TF32 cuDNN convolutions the reference
runs by default), accumulation fp32, relative xyz enters at ~22 bits via a hi/lo split.  We require
|err| <= 2e-3 * max|ref| + 2e-3 * |ref| per element.
"""
import numpy as np
import pytest
import torch
import torch.nn as nn

from tests.util import clouds

pytestmark = pytest.mark.gpu

from garment4d_b200.pointnet2 import pointnet2_modules as pm   # noqa: E402
from garment4d_b200.pointnet2 import pointnet2_utils as pu     # noqa: E402


def _randomise_bn(module, seed):
    g = torch.Generator().manual_seed(seed)
    for mod in module.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.weight.data = torch.rand(mod.num_features, generator=g) * 1.5 + 0.25
            mod.bias.data = torch.randn(mod.num_features, generator=g) * 0.2
            mod.running_mean = torch.randn(mod.num_features, generator=g) * 0.2
            mod.running_var = torch.rand(mod.num_features, generator=g) * 1.5 + 0.25


def _close(got, ref, tol=2e-3):
    err = (got - ref).abs()
    bound = tol * ref.abs().max() + tol * ref.abs()
    bad = err > bound
    assert not bool(bad.any()), f"max err {float(err.max()):.3e} (ref max {float(ref.abs().max()):.3e}), {int(bad.sum())} elements out of tolerance"


SPECS = [  # (N, npoint, C_in, radii, nsamples, mlps)
    (1024, 256, 0, [0.1, 0.2], [16, 32], [[0, 16, 16, 32], [0, 32, 32, 64]]),        # SA1-like on a small cloud
    (1024, 128, 96, [0.2, 0.3], [16, 32], [[96, 32, 32, 64], [96, 64, 64, 128]]),     # SA2
    (256, 64, 192, [0.3, 0.5], [32, 64], [[192, 64, 64, 128], [192, 128, 128, 256]]), # SA3
    (500, 50, 16, [0.25], [8], [[16, 16, 32, 48]]),                                   # ragged tile, nsample 8, single scale
    (300, 30, 8, [0.3], [128], [[8, 32, 16, 200]]),                                   # nsample 128, c3 = 200 (2 blocks, partial)
]


@pytest.mark.parametrize("spec", SPECS, ids=lambda s: f"N{s[0]}c{s[2]}")
def test_fused_sa_module_vs_torch(cuda, spec):
    N, npoint, cin, radii, nsamples, mlps = spec
    torch.manual_seed(0)
    mod = pm.PointnetSAModuleMSG(npoint=npoint, radii=radii, nsamples=nsamples, mlps=[list(m) for m in mlps], bn=True)
    _randomise_bn(mod, 1)
    mod = mod.to(cuda).eval()
    B = 3
    xyz = torch.from_numpy(clouds(7, B, N, "body")).to(cuda)
    feats = torch.randn(B, cin, N, device=cuda) if cin else None
    with torch.no_grad():
        new_xyz, out = mod(xyz, feats)                       # fused route
        assert pu.point_major_of(out) is not None, "fused route was not taken"
        mod.fused = False
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False              # true fp32 reference
        try:
            ref_xyz, ref = mod(xyz, feats)
        finally:
            torch.backends.cudnn.allow_tf32 = old
            mod.fused = True
    assert torch.equal(new_xyz, ref_xyz)
    assert out.shape == ref.shape
    _close(out, ref)
    _close(pu.point_major_of(out).float().transpose(1, 2), ref)


def test_fused_route_tracks_weight_updates(cuda):
    torch.manual_seed(1)
    mod = pm.PointnetSAModule(npoint=32, radius=0.3, nsample=16, mlp=[0, 16, 16, 32]).to(cuda).eval()
    xyz = torch.from_numpy(clouds(9, 2, 256, "cube")).to(cuda)
    with torch.no_grad():
        _, a = mod(xyz)
        for p in mod.parameters():
            p.mul_(0.5)
        _, b = mod(xyz)
        mod.fused = False
        _, ref = mod(xyz)
    assert not torch.allclose(a, b)
    _close(b, ref)


def test_training_mode_uses_operator_route_and_backprops(cuda):
    torch.manual_seed(2)
    mod = pm.PointnetSAModuleMSG(npoint=32, radii=[0.3], nsamples=[16], mlps=[[4, 16, 16, 32]]).to(cuda).train()
    xyz = torch.from_numpy(clouds(9, 2, 256, "cube")).to(cuda)
    feats = torch.randn(2, 4, 256, device=cuda, requires_grad=True)
    _, out = mod(xyz, feats)
    assert pu.point_major_of(out) is None
    out.sum().backward()
    assert feats.grad is not None and torch.isfinite(feats.grad).all()
    assert all(p.grad is not None for p in mod.parameters())


def test_training_mode_channels_last_layers_equal_the_nchw_ones(cuda):
    """Training mode runs torch's conv / BatchNorm / ReLU layers in channels_last memory (pytorch_utils.SharedMLP): same outputs,
    same gradients as the NCHW layers, batch statistics included."""
    from garment4d_b200.pointnet2 import pytorch_utils as ptu
    res = {}
    for cl in (True, False):
        torch.manual_seed(5)
        mod = pm.PointnetSAModuleMSG(npoint=64, radii=[0.2, 0.4], nsamples=[16, 32], mlps=[[6, 16, 16, 32], [6, 32, 32, 64]], bn=True).to(cuda).train()
        xyz = torch.from_numpy(clouds(9, 3, 512, "body")).to(cuda)
        feats = torch.randn(3, 6, 512, device=cuda, requires_grad=True)
        old, old_tf32 = ptu.SharedMLP.channels_last_training, torch.backends.cudnn.allow_tf32
        ptu.SharedMLP.channels_last_training = cl
        torch.backends.cudnn.allow_tf32 = False              # true fp32 in both layouts (TF32 kernels differ in rounding order)
        try:
            _, out = mod(xyz, feats)
            (out * torch.linspace(0.5, 1.5, out.shape[1], device=cuda)[None, :, None]).sum().backward()
        finally:
            ptu.SharedMLP.channels_last_training = old
            torch.backends.cudnn.allow_tf32 = old_tf32
        res[cl] = (out.detach(), feats.grad.detach(), [p.grad.detach().clone() for p in mod.parameters()],
                   [b.detach().clone() for n, b in mod.named_buffers() if "running" in n])
    a, b = res[True], res[False]
    _close(a[0], b[0], tol=1e-4)
    _close(a[1], b[1], tol=1e-3)
    for ga, gb in zip(a[2], b[2]):
        _close(ga, gb, tol=1e-3)
    for ra, rb in zip(a[3], b[3]):
        _close(ra, rb, tol=1e-5)


@pytest.mark.parametrize("N", [2048, 1000, 8192, 16384])       # 16384: BASELINE config c5 (row-wise pruned FPS, grid searches)
def test_fused_fp0_head_vs_modules(cuda, N):
    """Encoder forward with the fused FP0+head kernel vs the module-by-module route (cuDNN, true fp32)."""
    from garment4d_b200.encoder import Pointnet2MSGSEG
    torch.manual_seed(3)
    model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False)
    _randomise_bn(model, 4)
    for mod in model.modules():
        if isinstance(mod, nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.25, 1.75); mod.weight.data.uniform_(0.25, 1.75); mod.bias.data.normal_(0, 0.2)
    model = model.to(cuda).eval()
    pc = torch.from_numpy(clouds(11, 2, N, "body")).to(cuda)
    with torch.no_grad():
        _, sem, lf, lx = model(pc)
        model.fused = False
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            _, sem_ref, lf_ref, lx_ref = model(pc)
        finally:
            torch.backends.cudnn.allow_tf32 = old
            model.fused = True
    assert sem.shape == sem_ref.shape == (2, N, 7)
    # 7 chained fp16-operand layers (3 SA + 2 FP + 2 head) end to end, FP1/FP2 through TF32 cuDNN: 5e-3 of the range
    _close(lf[0], lf_ref[0], tol=5e-3)
    _close(sem, sem_ref, tol=5e-3)
    # the head kernel's last epilogue also wrote the arg-max labels (mesh_encoder.py:113); they are dropped once the logits change
    from garment4d_b200.pointnet2.pointnet2_cuda_bridge import segmentation_labels
    assert getattr(sem, "_g4d_labels", None) is not None
    lab = segmentation_labels(sem)
    assert lab.dtype == torch.uint8 and lab.data_ptr() == sem._g4d_labels[0].data_ptr()
    assert torch.equal(lab, sem.argmax(dim=2).to(torch.uint8))
    sem[:, :, 3] += 100.0
    lab2 = segmentation_labels(sem)
    assert lab2.data_ptr() != lab.data_ptr() and bool((lab2 == 3).all())


@pytest.mark.parametrize("spec", [(1024, 256, 256, 96, [352, 256, 128]), (256, 64, 384, 192, [576, 512, 256]), (300, 41, 16, 0, [16, 24])],
                         ids=lambda s: f"n{s[0]}m{s[1]}")
def test_fp_module_eval_routes_vs_operator_sequence(cuda, spec):
    """PointnetFPModule in eval mode (fused prologue + batched fp16 GEMMs, and the per-cloud conv route) against the
    reference's operator sequence in true fp32 (training-style path of the same module under no_grad is not available:
    run the module with autograd enabled so that it takes the operator route, BN in eval mode)."""
    from garment4d_b200.pointnet2 import pointnet2_modules as pm
    n, m, c2, c1, mlp = spec
    torch.manual_seed(5)
    mod = pm.PointnetFPModule(mlp=list(mlp), bn=True)
    _randomise_bn(mod, 6)
    mod = mod.to(cuda).eval()
    B = 3
    unknown = torch.from_numpy(clouds(12, B, n, "body")).to(cuda)
    known = unknown[:, :m].contiguous()
    kf = torch.randn(B, c2, m, device=cuda)
    skip = torch.randn(B, c1, n, device=cuda) if c1 else None
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = mod(unknown, known, skip, kf.clone().requires_grad_(True)).detach()      # operator route (autograd on)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    with torch.no_grad():
        for route in ("half", "half+pm", "tc+pm", "conv"):
            prev, pm._FP_GEMM = pm._FP_GEMM, route.split("+")[0]
            kin = kf.clone()
            sin = None if skip is None else skip.clone()
            if route.endswith("+pm") and c2 % 8 == 0:      # features with the fp16 point-major copies the fused levels attach
                pu.attach_point_major(kin, kin.transpose(1, 2).to(torch.float16).contiguous())
                if sin is not None:
                    pu.attach_point_major(sin, sin.transpose(1, 2).to(torch.float16).contiguous())       # -> the point-major rows route
            try:
                out = mod(unknown, known, sin, kin)
            finally:
                pm._FP_GEMM = prev
            assert out.shape == ref.shape and out.dtype == torch.float32
            _close(out, ref, tol=4e-3)


@pytest.mark.parametrize("graphed", [False, True], ids=["streams", "graph"])
def test_runner_host_path_equals_device_path(cuda, graphed):
    """EncoderLBSRunner: the end-to-end path (pinned host buffers, chunked streams, lbs on its own stream) returns exactly what
    the device-resident path computes, kernel by kernel and as a captured CUDA graph."""
    from garment4d_b200 import synthetic
    from garment4d_b200.encoder import Pointnet2MSGSEG
    from garment4d_b200.runner import EncoderLBSRunner, GraphedEncoderLBSRunner
    torch.manual_seed(7)
    model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(cuda).eval()
    smpl_np = synthetic.synthetic_smpl(seed=3)
    smpl = [torch.from_numpy(np.ascontiguousarray(smpl_np[k])).to(cuda) for k in
            ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights")]
    C, N = 6, 2048
    pc_pin = torch.from_numpy(clouds(21, C, N, "body")).pin_memory()
    b_np, p_np = synthetic.synthetic_frames(C, seed=4)
    betas_pin, pose_pin = torch.from_numpy(b_np).pin_memory(), torch.from_numpy(p_np).pin_memory()
    V = smpl[0].shape[0]
    lab = torch.zeros(C, N, dtype=torch.uint8).pin_memory()
    verts = torch.zeros(C, V, 3).pin_memory()
    joints = torch.zeros(C, 24, 3).pin_memory()
    pc, betas, pose = pc_pin.to(cuda), betas_pin.to(cuda), pose_pin.to(cuda)
    if graphed:
        r = GraphedEncoderLBSRunner(model, smpl, chunks=3, device=cuda)
        r.capture(pc, betas, pose)
        sem, v, j = r.replay_device()
        r.capture_host(pc_pin, betas_pin, pose_pin, lab, verts, joints)
        lab.zero_(); verts.zero_(); joints.zero_()
        r.replay_host()
    else:
        r = EncoderLBSRunner(model, smpl, chunks=3, device=cuda)
        sem, v, j = r.forward_device(pc, betas, pose)
        r.forward_host(pc_pin, betas_pin, pose_pin, lab, verts, joints)
    torch.cuda.synchronize()
    with torch.no_grad():
        _, sem_one, _, _ = model(pc)          # one big call: chunking must not change anything (every kernel works per cloud)
    assert torch.equal(sem, sem_one)
    assert torch.equal(lab, sem.argmax(dim=2).to(torch.uint8).cpu())
    assert torch.equal(verts, v.cpu()) and torch.equal(joints, j.cpu())


def test_graphed_runner_at_c3_size_equals_one_call(cuda):
    """The benchmarked configuration itself: 240 frames x 8192 points, 4 frame groups, one captured CUDA graph -- identical
    logits to ONE eager call over all frames, and the label map of the end-to-end (pinned host) graph equals their arg-max."""
    from garment4d_b200 import synthetic
    from garment4d_b200.encoder import Pointnet2MSGSEG
    from garment4d_b200.runner import GraphedEncoderLBSRunner
    import bench
    torch.manual_seed(7)
    model = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False).to(cuda).eval()
    smpl_np = synthetic.synthetic_smpl(seed=3)
    smpl = [torch.from_numpy(np.ascontiguousarray(smpl_np[k])).to(cuda) for k in
            ("v_template", "shapedirs", "posedirs", "J_regressor", "parents", "lbs_weights")]
    C, N = 240, 8192
    pc_pin = torch.from_numpy(bench.make_inputs("body", 99, C, N)).pin_memory()
    b_np, p_np = synthetic.synthetic_frames(C, seed=4)
    betas_pin, pose_pin = torch.from_numpy(b_np).pin_memory(), torch.from_numpy(p_np).pin_memory()
    V = smpl[0].shape[0]
    lab = torch.zeros(C, N, dtype=torch.uint8).pin_memory()
    verts = torch.zeros(C, V, 3).pin_memory()
    joints = torch.zeros(C, 24, 3).pin_memory()
    pc, betas, pose = pc_pin.to(cuda), betas_pin.to(cuda), pose_pin.to(cuda)
    r = GraphedEncoderLBSRunner(model, smpl, chunks=4, device=cuda)
    r.capture(pc, betas, pose)
    sem, v, j = r.replay_device()
    r.capture_host(pc_pin, betas_pin, pose_pin, lab, verts, joints)
    r.replay_host()
    torch.cuda.synchronize()
    with torch.no_grad():
        _, sem_one, _, _ = model(pc)
    assert torch.equal(sem, sem_one)
    assert torch.equal(lab, sem.argmax(dim=2).to(torch.uint8).cpu())
    assert torch.equal(verts, v.cpu()) and torch.equal(joints, j.cpu())
