#!/usr/bin/env bash
# Runs ON THE GPU BOX: FP-kernel tests, then bench variants.
set -u
TAG="${1:-r01c}"
OUT=gpurun_out
mkdir -p $OUT
timeout -k 10 300 python -m pytest tests/test_sa_mlp_gpu.py tests/test_parity_gpu.py -x -q -m gpu -k "fp0 or fp_interp or bias_relu" > $OUT/${TAG}_fp_tests.log 2>&1
rc=$?; echo "fp tests exit $rc"; tail -5 $OUT/${TAG}_fp_tests.log
if [ $rc -ne 0 ]; then exit $rc; fi
timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
G4D_FP_FUSED_CONV=1 timeout -k 10 300 python -m pytest tests/test_sa_mlp_gpu.py -x -q -m gpu -k "fp0" > $OUT/${TAG}_fusedconv_tests.log 2>&1; echo "fused conv tests exit $?"; tail -3 $OUT/${TAG}_fusedconv_tests.log
G4D_FP_FUSED_CONV=1 timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_fusedconv.json 2>> $OUT/${TAG}_bench.err; echo "bench fusedconv exit $?"
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r01c_bench*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
        for k in d["kernels"]:
            if any(s in k["name"] for s in ("FP", "fp_", "three_nn")): print("    %-70s %.4f" % (k["name"][:70], k["ms"]))
    except Exception as e:
        print(f, "unreadable", e)
P
