#!/usr/bin/env bash
set -u
TAG=${1:-r02m}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --config c3 --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-train --no-extras > $OUT/${TAG}_launches_run.log 2>&1
echo "ncu launches exit $?"
python tools/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.md 2>&1
head -60 $OUT/${TAG}_launches_summary.md
gzip -f $OUT/${TAG}_launches.csv
