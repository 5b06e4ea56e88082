#!/usr/bin/env bash
set -u
TAG=${1:-r02n}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 300 python -m pytest tests/test_mesh_ops_gpu.py tests/test_sa_mlp_gpu.py -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/${TAG}_pytest.log
timeout -k 10 200 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-train --no-extras > $OUT/${TAG}_c3.json 2> $OUT/${TAG}_c3.err; echo "bench c3 exit $?"; tail -3 $OUT/${TAG}_c3.err; head -c 300 $OUT/${TAG}_c3.json; echo
timeout -k 10 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_fp_launches.csv \
    python -m pytest tests/test_sa_mlp_gpu.py -x -q -m gpu -k "fp_module" > $OUT/${TAG}_ncu_fp.log 2>&1; echo "ncu fp pytest exit $?"; tail -5 $OUT/${TAG}_ncu_fp.log
grep -i "error\|mlp2" $OUT/${TAG}_fp_launches.csv | head -8 | cut -c1-250
