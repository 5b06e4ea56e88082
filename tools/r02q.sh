#!/usr/bin/env bash
set -u
TAG=${1:-r02q}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 600 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest_gpu.log
timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --config c3 --chunks 4 --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-train --no-extras > $OUT/${TAG}_launches_run.log 2>&1
echo "ncu launches (no graph) exit $?"
python tools/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.md 2>&1
head -40 $OUT/${TAG}_launches_summary.md
gzip -f $OUT/${TAG}_launches.csv
