#!/usr/bin/env bash
set -u
OUT=gpurun_out
timeout -k 10 700 python -m pytest tests/test_sa_mlp_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout -k 10 300 python bench.py --config c3 --no-cpu-baseline --no-train --no-extras --steps 5 --warmup 3 > $OUT/r02ai.json 2>> $OUT/r02ai.err
python - <<P
import json
d = json.load(open("$OUT/r02ai.json")); print("c3:", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
for k in d["kernels"]:
    if "fp_interp" in k["name"]: print(k["name"], k["ms"])
P
