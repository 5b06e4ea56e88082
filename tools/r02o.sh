#!/usr/bin/env bash
set -u
TAG=${1:-r02o}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 300 python -m pytest tests/test_mesh_ops_gpu.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -12 $OUT/${TAG}_pytest.log
timeout -k 10 300 compute-sanitizer --tool memcheck python tools/fp2_once.py 60 > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck exit $?"; grep -v "^$" $OUT/${TAG}_memcheck.log | tail -15
timeout -k 10 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_fp2_launches.csv python tools/fp2_once.py 60 > $OUT/${TAG}_ncu_fp2.log 2>&1; echo "ncu fp2 exit $?"; tail -3 $OUT/${TAG}_ncu_fp2.log
grep -i "error\|mlp2" $OUT/${TAG}_fp2_launches.csv | head -6 | cut -c1-260
