#!/usr/bin/env bash
set -u
OUT=gpurun_out
timeout -k 10 300 python bench.py --config c4_8 --no-cpu-baseline --no-train --no-extras --steps 5 --warmup 3 > $OUT/r02ag_120.json 2>> $OUT/r02ag.err
timeout -k 10 300 python bench.py --config c3 --no-cpu-baseline --no-train --no-extras --steps 5 --warmup 3 > $OUT/r02ag_240.json 2>> $OUT/r02ag.err
python - <<P
import json
a = json.load(open("$OUT/r02ag_120.json")); b = json.load(open("$OUT/r02ag_240.json"))
print("step ms", a["ms_per_step"], b["ms_per_step"])
kb = {k["name"]: k["ms"] for k in b["kernels"]}
for k in a["kernels"]:
    if "query_and_group" in k["name"]: continue
    print("  %-66s %.4f  %.4f  ratio %.2f" % (k["name"][:66], k["ms"], kb.get(k["name"], 0), k["ms"] / max(kb.get(k["name"], 1e-9), 1e-9)))
P
