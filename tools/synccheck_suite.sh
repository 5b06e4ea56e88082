#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
# (the reference's own FPS kernel misbehaves under the instrumentation: compare with the CPU oracle only)
G4D_NO_REFGPU=1 timeout -k 10 ${1:-1500} compute-sanitizer --tool synccheck --error-exitcode 3 --launch-timeout 0 \
    python -m pytest tests -x -q -m gpu --deselect tests/test_env_variants_gpu.py -p no:cacheprovider > $OUT/synccheck_suite.log 2>&1
echo "synccheck suite exit $?"
grep -E "ERROR SUMMARY|passed|failed|Barrier|error" $OUT/synccheck_suite.log | tail -8
