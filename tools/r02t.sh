#!/usr/bin/env bash
set -u
TAG=${1:-r02t}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 700 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest_gpu.log
timeout -k 10 300 python bench.py --config c3 --no-cpu-baseline --no-train --no-extras --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<P
import json
d = json.load(open("$OUT/${TAG}_bench.json")); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
for k in d["kernels"]: print("    %-70s %.4f  frac %.3f ref %s" % (k["name"][:70], k["ms"], k["roofline"]["frac"], k.get("ref_ms")))
P
timeout -k 10 400 python tools/sweep_ball_query.py 120 > $OUT/${TAG}_c5_sweep.md 2> $OUT/${TAG}_c5_sweep.err; echo "sweep exit $?"; tail -3 $OUT/${TAG}_c5_sweep.err
timeout -k 10 200 python bench.py --config c5 --no-cpu-baseline --no-train --no-extras --no-kernel-breakdown --steps 3 --warmup 3 > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err; echo "bench c5 exit $?"; tail -2 $OUT/${TAG}_bench_c5.err
timeout -k 10 120 python tools/sa_timeline.py 1 3 5 > $OUT/${TAG}_sa_timeline.txt 2>&1; echo "timeline exit $?"
