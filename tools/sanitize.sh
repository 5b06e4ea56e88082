#!/usr/bin/env bash
set -u
OUT=gpurun_out
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|Invalid|^(knn|garment|fps|bq|group|fp|sa) " $OUT/sanitize_memcheck.log | head -30
timeout -k 10 600 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_small.py knn fps bq group garment > $OUT/sanitize_racecheck.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|hazard|^(knn|garment|fps|bq|group) " $OUT/sanitize_racecheck.log | head -20
timeout -k 10 600 compute-sanitizer --tool synccheck --error-exitcode 3 python tools/sanitize_small.py fp sa fps > $OUT/sanitize_synccheck.log 2>&1; echo "synccheck exit $?"; grep -E "ERROR SUMMARY|Barrier|^(fp|sa|fps) " $OUT/sanitize_synccheck.log | head -20
