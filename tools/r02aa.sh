#!/usr/bin/env bash
set -u
TAG=${1:-r02aa}; OUT=gpurun_out
timeout -k 10 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout -k 10 300 python bench.py --config c3 --no-cpu-baseline --no-train --no-extras --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<P
import json
d = json.load(open("$OUT/${TAG}_bench.json")); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
for k in d["kernels"]:
    if "fp_" in k["name"] or "FP" in k["name"] or "ball" in k["name"]: print("    %-70s %.4f" % (k["name"][:70], k["ms"]))
P
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --config c3 --chunks 1 --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-train --no-extras > $OUT/${TAG}_launches_run.log 2>&1
echo "ncu launches exit $?"
python tools/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.md 2>&1
gzip -f $OUT/${TAG}_launches.csv
head -12 $OUT/${TAG}_launches_summary.md | cut -c1-160
