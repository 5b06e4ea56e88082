#!/usr/bin/env bash
set -u
TAG=${1:-r02y}; OUT=gpurun_out
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py tests/test_sa_mlp_gpu.py -x -q -m gpu 2>&1 | tail -4
for qo in 0 1; do
G4D_BQ_QUERY_ORDER=$qo timeout -k 10 300 python bench.py --config c3 --no-cpu-baseline --no-train --no-extras --steps 5 --warmup 3 > $OUT/${TAG}_bench$qo.json 2>> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<P
import json
d = json.load(open("$OUT/${TAG}_bench$qo.json")); print("query order $qo:", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
for k in d["kernels"]:
    if "ball" in k["name"] or "query_and_group L0" in k["name"] or "fps_gather L0" in k["name"]: print("    %-70s %.4f" % (k["name"][:70], k["ms"]))
P
done
