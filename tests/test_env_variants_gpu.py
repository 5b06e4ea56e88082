"""-m gpu: every tuning variable of the library (G4D_*) is exercised: a short parity pass (tests/variant_check.py) in a fresh process
per setting -- the C library reads most of them once per process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    {"G4D_SA_NSLOT": "2"}, {"G4D_SA_NSLOT": "1"}, {"G4D_SA_NI": "1"}, {"G4D_SA_SPIN": "1"},
    {"G4D_FPS_WS": "rows"}, {"G4D_FPS_WIDE": "0"}, {"G4D_FPS_WIDE": "1"}, {"G4D_FPS": "pruned"}, {"G4D_FPS": "plain"},
    {"G4D_MLP2_KS": "32"}, {"G4D_MLP2_KS": "16"}, {"G4D_BQ_QUERY_ORDER": "0"}, {"G4D_NN_CELLS": "0.7"},
    {"G4D_BACKWARD": "atomic"}, {"G4D_FP_GEMM": "half"}, {"G4D_FP_GEMM": "conv"},
]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_tuning_variable_keeps_parity(cuda, env):
    e = dict(os.environ)
    e.update(env)
    e["PYTHONPATH"] = ROOT + os.pathsep + e.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variant_check.py")], env=e, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert "variant ok" in r.stdout
