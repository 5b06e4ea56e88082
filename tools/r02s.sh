#!/usr/bin/env bash
set -u
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py tests/test_reference_layer_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout -k 10 300 python bench.py --config c3 --no-cpu-baseline --no-train --no-extras --steps 5 --warmup 3 > gpurun_out/r02s_bench.json 2>> gpurun_out/r02s_bench.err; echo "bench exit $?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02s_bench.json")); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
for k in d["kernels"]:
    if "query_and_group" in k["name"] or "ball" in k["name"]: print("    %-70s %.4f  frac %.3f ref %s" % (k["name"][:70], k["ms"], k["roofline"]["frac"], k.get("ref_ms")))
P
