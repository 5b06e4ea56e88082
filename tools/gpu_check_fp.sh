#!/usr/bin/env bash
# Runs ON THE GPU BOX: focused checks of the feature-propagation kernels first (short timeouts: a hang must not eat the
# budget), then the whole GPU suite, the bench line and a sweep over the number of frame groups per step.
set -u
TAG="${1:-r01b}"
OUT=gpurun_out
mkdir -p $OUT
timeout -k 10 300 python -m pytest tests/test_sa_mlp_gpu.py tests/test_parity_gpu.py -x -q -m gpu -k "fp0 or fp_interp or bias_relu" > $OUT/${TAG}_fp_tests.log 2>&1
rc=$?; echo "fp tests exit $rc"; tail -15 $OUT/${TAG}_fp_tests.log
if [ $rc -ne 0 ]; then exit $rc; fi
timeout -k 10 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_sa_mlp_gpu.py::test_fused_fp0_head_vs_modules[1000]" -x -q -m gpu > $OUT/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" $OUT/${TAG}_memcheck.log | tail -5
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
for ch in 1 2 3 6 8; do
  timeout 300 python bench.py --chunks $ch --steps 20 --no-cpu-baseline --no-kernel-breakdown > $OUT/${TAG}_bench_chunks$ch.json 2>> $OUT/${TAG}_bench.err
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/*_bench*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
    except Exception as e:
        print(f, "unreadable", e)
P
