#!/usr/bin/env bash
set -u
OUT=gpurun_out
for ch in 1 2 4; do
timeout -k 10 300 python bench.py --config c4_8 --chunks $ch --no-cpu-baseline --no-train --no-extras --no-kernel-breakdown --steps 20 --warmup 5 > $OUT/r02ad_$ch.json 2>> $OUT/r02ad.err
python - <<P
import json
d = json.load(open("$OUT/r02ad_$ch.json")); print("120 frames, chunks $ch:", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["ms_per_step"], 3))
P
done
for ch in 2 8; do
timeout -k 10 300 python bench.py --config c4 --chunks $ch --no-cpu-baseline --no-train --no-extras --no-kernel-breakdown --steps 10 --warmup 5 > $OUT/r02ad_c4_$ch.json 2>> $OUT/r02ad.err
python - <<P
import json
d = json.load(open("$OUT/r02ad_c4_$ch.json")); print("960 frames, chunks $ch:", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["ms_per_step"], 3))
P
done
