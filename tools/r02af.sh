#!/usr/bin/env bash
set -u
OUT=gpurun_out
timeout -k 10 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for cfg in c4_8 c3 c4; do
timeout -k 10 300 python bench.py --config $cfg --no-cpu-baseline --no-train --no-extras --no-kernel-breakdown --steps 10 --warmup 5 > $OUT/r02af_$cfg.json 2>> $OUT/r02af.err
python - <<P
import json
d = json.load(open("$OUT/r02af_$cfg.json")); print("$cfg:", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["ms_per_step"], 3))
P
done
