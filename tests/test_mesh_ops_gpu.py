"""-m gpu: the model-side callers of the hot path against the numpy restatement of the reference."""
import numpy as np
import pytest
import torch

from oracle import mesh_ops as omesh
from tests.util import clouds

pytestmark = pytest.mark.gpu

from garment4d_b200 import mesh_ops   # noqa: E402


@pytest.mark.parametrize("case", [(3, 2, 6890, 64, 1722, 1), (1, 4, 8192, 64, 2048, 4), (2, 1, 300, 5, 75, 0), (2, 2, 1000, 8, 250, 6)],
                         ids=lambda c: f"N{c[2]}n{c[4]}")
def test_garment_point_selection_equals_the_reference_loop(cuda, case):
    nbatch, T, N, Cf, n, label = case
    C = nbatch * T
    rs = np.random.RandomState(N + label)
    x = clouds(5, C, N, "body")
    logits = rs.randn(C, N, 7).astype(np.float32)
    logits[..., label] += 0.9                                  # ~ a third of the points: more and fewer than n per frame
    logits[0, :, label] -= 5.0 * (np.arange(N) % 3 != 0)       # frame 0: fewer than n selected -> zero padding
    logits[-1, ::7, :] = 0.25                                  # exact ties: arg-max takes the first class
    feat = rs.randn(C, Cf, N).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(cuda)
    gv, gf, cnt = mesh_ops.calc_segmentation_results(t(x).reshape(nbatch, T, N, 3), t(logits), n, nbatch, T, t(feat), label, return_counts=True)
    wv, wf = omesh.calc_segmentation_results(x, logits, n, feat, label)
    assert np.array_equal(gv.cpu().numpy(), wv) and np.array_equal(gf.cpu().numpy(), wf)
    assert np.array_equal(cnt.cpu().numpy(), (np.argmax(logits, 2) == label).sum(1))
    assert (cnt.cpu().numpy() < n).any() and (cnt.cpu().numpy() > n).any() or N < 1000


def test_garment_encoder_stack_fused_vs_operator_route(cuda):
    """Second SA stack on the selected garment points (mesh_encoder.py:54-78,149-161): fused tcgen05 route vs the module-by-module
    fp32 route; also the zero-padded rows (duplicates of the origin) go through FPS / ball query like any other point."""
    torch.manual_seed(3)
    stack = mesh_ops.GarmentEncoderStack(64).to(cuda).eval()
    C, n = 4, 2048
    gv = torch.from_numpy(clouds(9, C, n, "body")).to(cuda)
    gv[1, 1500:] = 0.0                                         # a frame with fewer than n garment points
    gf = torch.randn(C, 64, n, device=cuda)
    gf[1, :, 1500:] = 0.0
    with torch.no_grad():
        lx, lf, summ = stack(gv, gf)
        for m in list(stack.GarmentEncoder) + [stack.GarmentSummarize]:
            m.fused = False
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            lx_r, lf_r, summ_r = stack(gv, gf)
        finally:
            torch.backends.cudnn.allow_tf32 = old
    assert summ.shape == (C, 512)
    for a, b in zip(lx[1:], lx_r[1:]):
        assert torch.equal(a, b)
    for a, b in zip(lf[1:], lf_r[1:]):
        err = (a - b).abs()
        assert float(err.max()) <= 3e-3 * float(b.abs().max()) + 1e-3
    assert float((summ - summ_r).abs().max()) <= 3e-3 * float(summ_r.abs().max()) + 1e-3


@pytest.mark.parametrize("case", [(6890, 1500, 3, 0.1, 8), (6890, 1500, 3, 0.4, 32), (2048, 1500, 64, 0.1, 32), (512, 1500, 96, 0.2, 16),
                                  (64, 700, 384, 0.4, 8), (64, 700, 384, 0.05, 4)], ids=lambda c: f"N{c[0]}C{c[2]}K{c[4]}")
def test_positional_encoding_unit_vs_reference_sequence(cuda, case):
    """mesh_encoder.py:450-466: fused forward == the reference's operator sequence in fp32 (torch Linear, TF32 off), including
    rows whose ball is empty (idx all zero -> nsample copies of point 0); gradients (checkpointed backward) == autograd of that
    sequence, w.r.t. the garment vertices and the four Linear tensors."""
    N, P, C, radius, K = case
    B = 2
    torch.manual_seed(N + K)
    pe = mesh_ops.PositionalEncoding(radius, K, 3 + C).to(cuda)
    xyz = torch.from_numpy(clouds(3, B, N, "body")).to(cuda)
    new_xyz = (torch.from_numpy(clouds(4, B, P, "body")).to(cuda) + 0.01 * torch.randn(B, P, 3, device=cuda))
    new_xyz[:, :7] += 5.0                                        # centroids with an empty ball
    new_xyz = new_xyz.contiguous().requires_grad_(True)
    feats = torch.randn(B, C, N, device=cuda)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = pe(xyz, new_xyz, feats)
        g = torch.randn_like(out)
        out.backward(g)
        got = [new_xyz.grad.clone()] + [p.grad.clone() for p in pe.mlp.parameters()]
        new_xyz.grad = None
        pe.zero_grad()
        pe.fused = False
        ref = pe(xyz, new_xyz, feats)
        ref.backward(g)
        want = [new_xyz.grad.clone()] + [p.grad.clone() for p in pe.mlp.parameters()]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert out.shape == ref.shape == (B, P, 32)
    assert float((out - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    for a, b in zip(got, want):
        assert float((a - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max()))
