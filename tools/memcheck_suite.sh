#!/usr/bin/env bash
# Runs ON THE GPU BOX: the GPU test suite (without the per-variable subprocess tests) under compute-sanitizer memcheck.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 ${1:-1500} compute-sanitizer --tool memcheck --error-exitcode 3 --launch-timeout 0 \
    python -m pytest tests -x -q -m gpu --deselect tests/test_env_variants_gpu.py -p no:cacheprovider > $OUT/memcheck_suite.log 2>&1
echo "memcheck suite exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|error" $OUT/memcheck_suite.log | tail -8
