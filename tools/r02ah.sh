#!/usr/bin/env bash
set -u
OUT=gpurun_out
timeout -k 10 700 python -m pytest tests/test_sa_mlp_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout -k 10 300 python bench.py --config c4 --no-cpu-baseline --no-train --no-extras --no-kernel-breakdown --steps 10 --warmup 5 > $OUT/r02ah_c4.json 2>> $OUT/r02ah.err
python - <<P
import json
d = json.load(open("$OUT/r02ah_c4.json")); print("c4:", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["ms_per_step"], 3), d["gpu_launches"])
P
