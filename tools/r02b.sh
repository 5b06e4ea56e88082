#!/usr/bin/env bash
# Runs ON THE GPU BOX (1 GPU): GPU test suite + smoke + new bench flow (c4 strong) ; r02b
set -u
TAG=r02b; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/${TAG}_smoke.log
timeout -k 10 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
tail -3 $OUT/${TAG}_bench.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; echo "ref exit $?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02b_bench.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["gpu_launches"])
print("cube", d.get("cube")); print("labels", d.get("label_agreement")); print("train", d.get("train")); print("cpu", d.get("cpu_baseline"))
print("roofline", d.get("roofline"))
for k in d["kernels"]:
    print("  %-72s %8.4f  ref %s" % (k["name"][:72], k["ms"], k.get("ref_ms")))
r = json.load(open("gpurun_out/r02b_bench_reference.json")); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
P
