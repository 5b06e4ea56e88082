"""CPU checks of kernel LOGIC that is still experimental on the GPU side: tests/emul/*.c restate a kernel statement by statement
(one emulated warp = 32 entries) so that its control flow -- not its speed -- can be validated here against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pointnet2 as orc
from tests.util import clouds

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libemul.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "emul", "three_nn_coop_emul.c"), "-lm"])
    L = ctypes.CDLL(so)
    L.three_nn_coop_emul.restype = ctypes.c_longlong
    L.three_nn_coop_emul.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return L


def _case(shape, n, m, seed):
    rs = np.random.RandomState(seed)
    if shape in ("body", "cube"):
        xyz = clouds(seed, 1, n, shape, dup_frac=0.1)[0]
    elif shape == "coincident":
        xyz = np.full((n, 3), 0.25, np.float32)
    elif shape == "outlier":
        xyz = clouds(seed, 1, n, "body")[0]
        xyz[7] = 1e6
    elif shape == "line":
        xyz = np.zeros((n, 3), np.float32); xyz[:, 0] = rs.rand(n).astype(np.float32)
    else:  # planar_dups: few distinct positions on a plane
        base = rs.rand(50, 3).astype(np.float32); base[:, 2] = 0.5
        xyz = base[rs.randint(0, 50, n)]
    known = xyz[rs.permutation(n)[:m]].copy() if m <= n else clouds(seed + 1, 1, m, "body")[0]
    return np.ascontiguousarray(xyz, np.float32), np.ascontiguousarray(known, np.float32)


@pytest.mark.parametrize("shape", ["body", "cube", "coincident", "outlier", "line", "planar_dups"])
@pytest.mark.parametrize("nm", [(2048, 256), (4096, 1024), (1000, 2), (777, 1500)], ids=lambda v: f"n{v[0]}m{v[1]}")
def test_three_nn_coop_logic_equals_bruteforce(emul, shape, nm):
    """The warp-cooperative search (grow a block of cells, scan only new cells, stop on the face bound) returns exactly the
    oracle's three nearest neighbours, ties and degenerate clouds included; known-grid and unknown-grid cell sizes as used by
    three_nn_raw / the SA modules."""
    n, m = nm
    unknown, known = _case(shape, n, m, seed=100 + n + m)
    want_d, want_i = orc.three_nn(unknown[None], known[None])          # sqrt distances, indices
    for ucell, kcell in ((0.1, -max(4.0, round(0.7 * m ** 0.5))), (-24.0, -4.0), (0.05, -64.0)):
        d2 = np.full((n, 3), -1, np.float32)
        idx = np.full((n, 3), -1, np.int32)
        passes = ctypes.c_longlong(0)
        visits = emul.three_nn_coop_emul(n, m, unknown.ctypes.data, known.ctypes.data, ucell, kcell, d2.ctypes.data, idx.ctypes.data,
                                         ctypes.byref(passes))
        assert np.array_equal(idx, want_i[0]), (shape, ucell, kcell)
        assert np.array_equal(np.sqrt(d2), want_d[0])
        assert visits <= (n + 31) // 32 * m * 1, "never more work than one brute-force scan per warp"
