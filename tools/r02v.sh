#!/usr/bin/env bash
set -u
TAG=${1:-r02v}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 400 python -m pytest tests/test_sa_mlp_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout -k 10 300 python bench.py --config c3 --no-cpu-baseline --no-train --no-extras --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}_bench.err; echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
python - <<P
import json
d = json.load(open("$OUT/${TAG}_bench.json")); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
for k in d["kernels"]: print("    %-70s %.4f" % (k["name"][:70], k["ms"]))
P
