#!/usr/bin/env bash
set -u
TAG=${1:-r02ac}; OUT=gpurun_out
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py tests/test_reference_layer_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout -k 10 300 python bench.py --config c3 --no-cpu-baseline --no-train --no-extras --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2>> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<P
import json
d = json.load(open("$OUT/${TAG}_bench.json")); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
tot = 0; by = 0
for k in d["kernels"]:
    if "query_and_group" in k["name"]:
        print("    %-70s %.4f frac %.3f" % (k["name"][:70], k["ms"], k["roofline"]["frac"])); tot += k["ms"]; by += k["roofline"]["achieved"] * k["ms"]
print("aggregate GB/s", by / tot, "frac", by / tot / 6550.4)
P
