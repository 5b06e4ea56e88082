"""CPU tests (-m "not gpu") of the host side: the C-ABI library loads and exports every symbol the header declares,
the Python mirror keeps the reference's names / signatures / state-dict keys, parameter folding and packing,
sharding logic under a world_size-2 gloo group.  No compute call is made (no GPU here)."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "garment4d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(g4d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from garment4d_b200 import _lib
    assert os.path.exists(_lib.SO_PATH), "run __graft_entry__.build() first"
    L = ctypes.CDLL(_lib.SO_PATH)
    declared = _header_symbols()
    assert len(declared) >= 25
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/garment4d_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "ctypes signature table and header disagree"
    L.g4d_abi_version.restype = ctypes.c_int
    assert L.g4d_abi_version() == 1


def test_library_is_sm100a_native_and_has_tcgen05():
    from garment4d_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3g4d17sa_mlp_max_kernelILi32ELb1ELb1EEEvNS_9SaMlpArgsE",
                           _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass, "grouped-MLP kernel is not on the tcgen05 path"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from garment4d_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.G4DError, match="no CPU fallback"):
        _lib.lib()


def test_pointnet2_cuda_mirror_has_the_reference_surface():
    import pointnet2_cuda                                 # top-level drop-in name (pointnet2_utils.py:7)
    want = {  # pointnet2_api.cpp:10-23 with the wrapper argument counts
        "ball_query_wrapper": 8, "group_points_wrapper": 8, "group_points_grad_wrapper": 8, "gather_points_wrapper": 7,
        "gather_points_grad_wrapper": 7, "furthest_point_sampling_wrapper": 6, "three_nn_wrapper": 7,
        "three_interpolate_wrapper": 8, "three_interpolate_grad_wrapper": 8}
    for name, nargs in want.items():
        fn = getattr(pointnet2_cuda, name)
        assert len(inspect.signature(fn).parameters) == nargs
    # argument order of ball_query: new_xyz before xyz (ball_query.cpp:14-15)
    assert list(inspect.signature(pointnet2_cuda.ball_query_wrapper).parameters)[5:7] == ["new_xyz_tensor", "xyz_tensor"]
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        pointnet2_cuda.ball_query_wrapper(1, 4, 1, 0.1, 2, torch.zeros(1, 1, 3), torch.zeros(1, 4, 3), torch.zeros(1, 1, 2, dtype=torch.int32))


def test_operator_and_module_names_match_reference():
    from garment4d_b200.pointnet2 import pointnet2_modules as pm, pointnet2_utils as pu, pytorch_utils as ptu
    from garment4d_b200 import lbs
    for n in ("furthest_point_sample", "gather_operation", "three_nn", "three_interpolate", "grouping_operation", "ball_query",
              "QueryAndGroup", "GroupAll", "FurthestPointSampling", "GatherOperation", "ThreeNN", "ThreeInterpolate",
              "GroupingOperation", "BallQuery"):
        assert hasattr(pu, n)
    for n in ("PointnetSAModuleMSG", "PointnetSAModule", "PointnetFPModule", "_PointnetSAModuleBase"):
        assert hasattr(pm, n)
    for n in ("SharedMLP", "Conv1d", "Conv2d", "FC", "BatchNorm1d", "BatchNorm2d"):
        assert hasattr(ptu, n)
    for n in ("lbs", "batch_rodrigues", "batch_rigid_transform", "vertices2joints", "vertices2jointsB", "blend_shapes", "transform_mat"):
        assert hasattr(lbs, n)
    assert list(inspect.signature(lbs.lbs).parameters) == ["betas", "pose", "v_template", "shapedirs", "posedirs", "J_regressor",
                                                           "parents", "lbs_weights", "pose2rot"]


def test_encoder_state_dict_keys_match_reference_naming():
    from garment4d_b200.encoder import Pointnet2MSGSEG
    m = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False)
    keys = set(m.state_dict().keys())
    # names from pytorch_utils.py:21-22,83,91,95,108 (checkpoint row)
    for k in ("SA_modules.0.mlps.1.layer2.conv.weight", "SA_modules.0.mlps.1.layer2.bn.bn.weight",
              "SA_modules.0.mlps.1.layer2.bn.bn.running_mean", "SA_modules.0.mlps.1.layer2.bn.bn.num_batches_tracked",
              "SA_modules.2.mlps.1.layer0.conv.weight", "FP_modules.0.mlp.layer0.conv.weight", "FP_modules.2.mlp.layer1.bn.bn.bias",
              "FC_layer.0.conv.weight", "FC_layer.0.bn.bn.running_var", "FC_layer.2.conv.weight", "FC_layer.2.conv.bias"):
        assert k in keys, k
    assert m.state_dict()["SA_modules.2.mlps.1.layer0.conv.weight"].shape == (128, 195, 1, 1)
    assert not any("conv.bias" in k for k in keys if "SA_modules" in k)     # bias dropped when bn=True (pytorch_utils.py:56)
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == sum(p.numel() for p in Pointnet2MSGSEG(input_channels=0, global_feat=False).parameters())
    assert Pointnet2MSGSEG(input_channels=0, global_feat=True).Middle_modules is not None


def test_fold_shared_mlp_equals_conv_bn_relu_on_cpu():
    from garment4d_b200.pointnet2 import pytorch_utils as ptu
    torch.manual_seed(0)
    mlp = ptu.SharedMLP([7, 16, 32, 24], bn=True).eval()
    for mod in mlp.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(); mod.running_var.uniform_(0.5, 2); mod.weight.data.uniform_(0.5, 2); mod.bias.data.normal_()
    x = torch.randn(2, 7, 5, 3)
    with torch.no_grad():
        ref = mlp(x.clone())
        y = x
        for w, b in ptu.fold_shared_mlp(mlp):
            y = torch.relu(torch.einsum("oc,bcpk->bopk", w, y) + b[None, :, None, None])
    assert torch.allclose(y, ref, atol=1e-5, rtol=1e-5)
    v0 = ptu.shared_mlp_version(mlp)
    with torch.no_grad():
        mlp[0][0].weight.mul_(2)          # optimizer steps / load_state_dict bump the version counter the same way
    assert ptu.shared_mlp_version(mlp) != v0
    assert ptu.fold_shared_mlp(ptu.SharedMLP([4, 8], bn=True, preact=True, first=False)) is None     # pre-activation order: not foldable


def test_sa_mlp_param_packing_layout():
    """g4d_sa_mlp_pack_params (host function, no GPU): UMMA canonical K-major images; hi/lo split of the xyz columns; the
    layer-1 bias in the two spare K positions behind them (they meet the constant 1.0 the producers write), the layer-2
    bias as one extra K-slice of W2 (it meets the constant ones operand), b3 in fp32; the fp16 range guard."""
    from garment4d_b200 import _lib
    L = _lib.lib()
    c_in, c1, c2, c3, K = 16, 32, 16, 40, 16
    d = _lib.SaMlpDesc(c_in, c1, c2, c3, K, L.g4d_sa_mlp_k0(c_in))
    assert d.k0 == 32
    nbytes = L.g4d_sa_mlp_param_bytes(ctypes.byref(d))
    c3p = 128
    assert nbytes == 2 * (d.k0 * c1 + (c1 + 16) * c2 + c2 * c3p) + 4 * c3p + 128 * 16 * 2        # (+ 16 * c1 for xyz-only levels)
    rs = np.random.RandomState(0)
    w1, w2, w3 = (rs.randn(c1, 3 + c_in).astype(np.float32), rs.randn(c2, c1).astype(np.float32), rs.randn(c3, c2).astype(np.float32))
    b1, b2, b3 = (rs.randn(c).astype(np.float32) for c in (c1, c2, c3))
    blob = np.zeros(nbytes, np.uint8)
    rc = L.g4d_sa_mlp_pack_params(ctypes.byref(d), *(a.ctypes.data for a in (w1, b1, w2, b2, w3, b3)), blob.ctypes.data)
    assert rc == 0

    def canon(off, rows, k):                                     # [k/8][row][k%8] fp16 -> (row, k) fp32
        img = blob[off:off + 2 * rows * k].view(np.float16).reshape(k // 8, rows, 8)
        return img.transpose(1, 0, 2).reshape(rows, k).astype(np.float32), off + 2 * rows * k

    h = lambda a: a.astype(np.float16).astype(np.float32)
    W1, o = canon(0, c1, d.k0)
    assert np.array_equal(W1[:, :c_in], h(w1[:, 3:]))                                            # features first
    wh = h(w1[:, :3])
    assert np.array_equal(W1[:, c_in:c_in + 3], wh) and np.array_equal(W1[:, c_in + 3:c_in + 6], wh)
    assert np.allclose(W1[:, c_in + 6:c_in + 9], w1[:, :3] - wh, atol=1e-6)
    assert np.abs(W1[:, c_in:c_in + 3] + W1[:, c_in + 6:c_in + 9] - w1[:, :3]).max() < 2e-6     # hi + lo recovers fp32 weights
    assert np.array_equal(W1[:, c_in + 9], h(b1)) and np.abs(W1[:, c_in + 9] + W1[:, c_in + 10] - b1).max() < 2e-6
    assert not W1[:, c_in + 11:].any()
    W2, o = canon(o, c2, c1 + 16)
    assert np.array_equal(W2[:, :c1], h(w2))
    assert np.array_equal(W2[:, c1], h(b2)) and np.abs(W2[:, c1] + W2[:, c1 + 1] - b2).max() < 2e-6 and not W2[:, c1 + 2:].any()
    W3, o = canon(o, c3p, c2)
    # c3 = 40 <= 64 and nsample = 16 <= 64: W3 is replicated twice down the 128 lanes (rows 0.. and 64..), zero padding between
    assert np.array_equal(W3[:c3], h(w3)) and np.array_equal(W3[64:64 + c3], h(w3)) and not W3[c3:64].any() and not W3[64 + c3:].any()
    assert np.array_equal(blob[o:o + 4 * c3].view(np.float32), b3)
    o += 4 * c3p
    ones, o = canon(o, 128, 16)
    assert (ones[:, :2] == 1).all() and not ones[:, 2:].any()
    assert o == nbytes
    # xyz-only level: the fp32 layer-1 weights (wx, wy, wz, b) ride behind the ones operand (layer 1 runs on CUDA cores)
    d0 = _lib.SaMlpDesc(0, c1, c2, c3, K, L.g4d_sa_mlp_k0(0))
    n0 = L.g4d_sa_mlp_param_bytes(ctypes.byref(d0))
    assert n0 == 2 * (d0.k0 * c1 + (c1 + 16) * c2 + c2 * c3p) + 4 * c3p + 128 * 16 * 2 + 16 * c1
    blob0 = np.zeros(n0, np.uint8)
    w1x = np.ascontiguousarray(w1[:, :3])
    assert L.g4d_sa_mlp_pack_params(ctypes.byref(d0), *(a.ctypes.data for a in (w1x, b1, w2, b2, w3, b3)), blob0.ctypes.data) == 0
    w1f = blob0[n0 - 16 * c1:].view(np.float32).reshape(c1, 4)
    assert np.array_equal(w1f[:, :3], w1x) and np.array_equal(w1f[:, 3], b1)
    bad = _lib.SaMlpDesc(c_in, 20, c2, c3, K, d.k0)
    assert L.g4d_sa_mlp_param_bytes(ctypes.byref(bad)) == 0 and b"multiples of 16" in L.g4d_last_error()
    # a folded weight or bias outside the fp16 range is refused (the caller then takes the operator route)
    w2_big = w2.copy(); w2_big[3, 5] = 7.0e4
    assert L.g4d_sa_mlp_pack_params(ctypes.byref(d), *(a.ctypes.data for a in (w1, b1, w2_big, b2, w3, b3)), blob.ctypes.data) != 0
    assert b"fp16 range" in L.g4d_last_error()
    b1_big = b1.copy(); b1_big[0] = -1.0e5
    assert L.g4d_sa_mlp_pack_params(ctypes.byref(d), *(a.ctypes.data for a in (w1, b1_big, w2, b2, w3, b3)), blob.ctypes.data) != 0


def test_sa_route_falls_back_when_folded_weights_leave_fp16_range():
    """BatchNorm statistics that push a folded weight past 65504: the fused branch is refused, the module takes the operator route."""
    import torch
    from garment4d_b200.pointnet2 import pointnet2_modules as pm
    sa = pm.PointnetSAModuleMSG(npoint=16, radii=[0.2], nsamples=[16], mlps=[[0, 16, 16, 32]]).eval()
    assert sa._branch(0, 0, "cpu") is not None
    with torch.no_grad():
        sa.mlps[0].layer1.bn.bn.running_var.fill_(1e-12)          # scale = gamma / sqrt(var + eps) ~ 316 ... not enough alone
        sa.mlps[0].layer1.bn.bn.weight.fill_(1e4)                 # ... times gamma = 1e4 -> folded weights ~ 3e6 * w
    assert sa._branch(0, 0, "cpu") is None


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-clouds", "1", "--config", "c2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0


# ---- multi-process sharding (world_size 2, gloo) -------------------------------------------------------------

_WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["G4D_ROOT"])
import numpy as np, torch, torch.distributed as dist
from garment4d_b200.sharding import shard_sequences, max_over_ranks, FlatGradientReducer, allreduce_flat_gradients
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
mine = shard_sequences(7, r, w)
t = max_over_ranks(float(10 + r))
own = torch.zeros(7, dtype=torch.int64)
for i in mine: own[i] += 1
dist.all_reduce(own)
# flat gradient all-reduce: rank 1 leaves one parameter without a gradient (it must still take part, as zeros)
torch.manual_seed(0)
lin = torch.nn.Linear(4, 3); extra = torch.nn.Parameter(torch.ones(5))
params = list(lin.parameters()) + [extra]
red = FlatGradientReducer(params)
x = torch.full((2, 4), float(r + 1))
loss = lin(x).sum() + (extra.sum() * 3.0 if r == 0 else 0.0)
loss.backward()
red.all_reduce()
g_w = lin.weight.grad.clone(); g_e = extra.grad.clone()
# one-shot helper, same result
lin2 = torch.nn.Linear(4, 3); lin2.load_state_dict(lin.state_dict()); extra2 = torch.nn.Parameter(torch.ones(5))
loss2 = lin2(x).sum() + (extra2.sum() * 3.0 if r == 0 else 0.0)
loss2.backward()
allreduce_flat_gradients(list(lin2.parameters()) + [extra2])
same = bool(torch.equal(lin2.weight.grad, g_w) and torch.equal(extra2.grad, g_e))
if r == 0:
    print(json.dumps({"cover": own.tolist(), "tmax": t, "mine": mine, "gw": g_w[0].tolist(), "ge": g_e.tolist(), "same": same,
                      "nbytes": red.nbytes}))
dist.destroy_process_group()
'''


def test_sharding_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, G4D_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29731", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    # the reference sampler's partition (utils/train_utils.py:12-31): strided, padded by wrap-around to equal shards
    assert res["mine"] == [0, 2, 4, 6]
    assert res["cover"] == [2, 1, 1, 1, 1, 1, 1]          # 7 sequences on 2 ranks: 4 + 4, sequence 0 owned twice (the padding)
    assert res["tmax"] == 11.0                # max over ranks, as bench.py reports
    # d(sum(lin(x)))/dW = column sums of x = 2*(r+1) per entry -> mean over ranks (2 + 4) / 2 = 3; extra: (3 + 0) / 2
    assert res["gw"] == [3.0] * 4 and res["ge"] == [1.5] * 5 and res["same"]
    assert res["nbytes"] == 4 * (12 + 3 + 5)


def test_shard_sequences_matches_the_reference_sampler():
    from garment4d_b200.sharding import shard_sequences
    for n, w in ((32, 8), (32, 4), (7, 2), (5, 8), (1, 4), (0, 2)):
        shards = [shard_sequences(n, r, w) for r in range(w)]
        per = -(-n // w) if n else 0
        assert all(len(s) == per for s in shards)
        total = per * w
        idx = list(range(n))
        while n and len(idx) < total:
            idx += idx[: total - len(idx)]
        assert shards == [idx[r:total:w] for r in range(w)]
        assert set(i for s in shards for i in s) == set(range(n))


def test_bench_reads_ncu_traffic_from_committed_profiles():
    """bench.py's roofline.traffic comes from profiles/*_ncu_summary.json: the lookups it makes must resolve."""
    import glob
    import json
    import bench
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summary.json")))
    assert cands, "no committed ncu summary"
    rows = json.load(open(cands[-1]))
    for key in [("fps_pruned_kernel<512", 0), ("ball_query_grid_kernel<2>", 0), ("sa_mlp_max_kernel<32, 1, ", 1), ("fp_interp_mlp_kernel", 0),
                ("three_nn_grid_kernel", 0), [("ball_query_grid_kernel<1>", 1), ("group_fused_kernel", 1)],
                [("ball_query_kernel<1>", 3), ("group_rows_kernel", 3)]]:
        v = bench.ncu_traffic(rows, key)
        assert v is not None and v > 0, key
    assert bench.ncu_traffic(rows, ("no_such_kernel", 0)) is None
    assert bench.ncu_traffic(rows, None) is None and bench.ncu_traffic([], ("fps", 0)) is None


def test_fp_param_packing_layout():
    """g4d_fp_pack_params (host function, no GPU): the four weight matrices as UMMA canonical K-major images, fp32 biases behind
    them, class rows padded to 16; and the descriptor checks."""
    from garment4d_b200 import _lib
    L = _lib.lib()
    c_in, c1, c2, h1, h2 = 128, 128, 64, 32, 7
    d = _lib.FpDesc(c_in, c1, c2, h1, h2)
    nbytes = L.g4d_fp_param_bytes(ctypes.byref(d))
    h2p = 16
    assert nbytes == 2 * (c_in * c1 + c1 * c2 + c2 * h1 + h1 * h2p) + 4 * (c1 + c2 + h1 + h2p)
    rs = np.random.RandomState(1)
    w1, w2, w3, w4 = (rs.randn(c1, c_in).astype(np.float32), rs.randn(c2, c1).astype(np.float32), rs.randn(h1, c2).astype(np.float32),
                      rs.randn(h2, h1).astype(np.float32))
    b1, b2, b3, b4 = (rs.randn(c).astype(np.float32) for c in (c1, c2, h1, h2))
    blob = np.zeros(nbytes, np.uint8)
    rc = L.g4d_fp_pack_params(ctypes.byref(d), *(a.ctypes.data for a in (w1, b1, w2, b2, w3, b3, w4, b4)), blob.ctypes.data)
    assert rc == 0

    def canon(off, rows, k):                                     # [k/8][row][k%8] fp16 -> (row, k) fp32
        img = blob[off:off + 2 * rows * k].view(np.float16).reshape(k // 8, rows, 8)
        return img.transpose(1, 0, 2).reshape(rows, k).astype(np.float32), off + 2 * rows * k

    h = lambda a: a.astype(np.float16).astype(np.float32)
    W1, o = canon(0, c1, c_in)
    W2, o = canon(o, c2, c1)
    W3, o = canon(o, h1, c2)
    W4, o = canon(o, h2p, h1)
    assert np.array_equal(W1, h(w1)) and np.array_equal(W2, h(w2)) and np.array_equal(W3, h(w3))
    assert np.array_equal(W4[:h2], h(w4)) and not W4[h2:].any()              # padded class rows are zero
    for b, n in ((b1, c1), (b2, c2), (b3, h1)):
        assert np.array_equal(blob[o:o + 4 * n].view(np.float32), b)
        o += 4 * n
    assert np.array_equal(blob[o:o + 4 * h2].view(np.float32), b4) and not blob[o + 4 * h2:o + 4 * h2p].any()
    # no head: h1 = 0 drops the head matrices; bad widths are refused with a message
    assert L.g4d_fp_param_bytes(ctypes.byref(_lib.FpDesc(c_in, c1, c2, 0, 0))) == 2 * (c_in * c1 + c1 * c2) + 4 * (c1 + c2)
    assert L.g4d_fp_param_bytes(ctypes.byref(_lib.FpDesc(c_in, 100, c2, h1, h2))) == 0 and b"multiples of 16" in L.g4d_last_error()
    assert L.g4d_fp_param_bytes(ctypes.byref(_lib.FpDesc(c_in, c1, c2, h1, 17))) == 0


def test_encoder_marks_the_fp_levels_that_feed_fused_gathers():
    """The two coarser FP levels emit the fp16 point-major copy their consumer gathers from; the finest one does not."""
    from garment4d_b200.encoder import Pointnet2MSGSEG
    m = Pointnet2MSGSEG(input_channels=0, bn=True, global_feat=False)
    assert [fp.emit_point_major for fp in m.FP_modules] == [False, True, True]


def _grid_adjacency(nu, nv):
    idx = np.arange(nu * nv).reshape(nu, nv)
    A = np.zeros((nu * nv, nu * nv), np.float32)
    for a, b in ((idx[:-1, :].ravel(), idx[1:, :].ravel()), (idx[:, :-1].ravel(), idx[:, 1:].ravel()), (idx[:-1, :-1].ravel(), idx[1:, 1:].ravel())):
        A[a, b] = 1
        A[b, a] = 1
    return A


def test_smoothing_operator_is_row_normalised_adjacency_minus_identity():
    """mesh_ops.smoothing_operator = pygcn/utils.py:56-63 normalize(adj_old) - eye (mesh_encoder.py:386), as CSR; dense, torch-sparse and
    scipy inputs agree."""
    import scipy.sparse as sp
    import torch
    from garment4d_b200 import mesh_ops
    adj = _grid_adjacency(7, 5)
    G = adj.shape[0]
    want = np.diag(np.power(adj.sum(1), -1.0)) @ adj - np.eye(G, dtype=np.float32)
    for inp in (sp.coo_matrix(adj), torch.from_numpy(adj), torch.from_numpy(adj).to_sparse()):
        rp, c, v = (x.numpy() for x in mesh_ops.smoothing_operator(inp, "cpu"))
        dense = np.zeros((G, G), np.float32)
        for g in range(G):
            dense[g, c[rp[g]:rp[g + 1]]] = v[rp[g]:rp[g + 1]]
        assert np.allclose(dense, want, atol=1e-7)


def test_oracle_knn_points_is_a_sorted_brute_force():
    from oracle import mesh_ops as omesh
    rs = np.random.RandomState(5)
    p1, p2 = rs.randn(2, 40, 3).astype(np.float32), rs.randn(2, 90, 3).astype(np.float32)
    p2[:, 50:60] = p2[:, :10]                                   # ties
    d, i = omesh.knn_points(p1, p2, K=17)
    for b in range(2):
        full = ((p1[b][:, None].astype(np.float64) - p2[b][None].astype(np.float64)) ** 2).sum(-1)
        assert np.allclose(d[b], np.sort(full, axis=1)[:, :17], rtol=1e-5, atol=1e-7)
        assert (np.diff(d[b], axis=1) >= 0).all()
        same = np.diff(d[b], axis=1) == 0
        assert (np.diff(i[b], axis=1)[same] > 0).all()          # equal distances: ascending index
        assert np.array_equal(np.take_along_axis(full, i[b], axis=1).astype(np.float32) >= 0, np.ones_like(d[b], bool))
