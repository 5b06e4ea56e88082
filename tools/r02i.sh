#!/usr/bin/env bash
set -u
TAG=${1:-r02i}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 600 python -m pytest tests/test_sa_mlp_gpu.py tests/test_parity_gpu.py -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_gpu.log
( timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing ) > $OUT/${TAG}_sa.txt 2>&1; cat $OUT/${TAG}_sa.txt
( G4D_SA_NSLOT=1 timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing ) 
( G4D_SA_NSLOT=2 timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing ) 
timeout -k 10 120 python tools/sa_timeline.py 1 2>&1 | tail -14
