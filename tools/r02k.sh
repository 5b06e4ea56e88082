#!/usr/bin/env bash
set -u
TAG=${1:-r02k}; OUT=gpurun_out; mkdir -p $OUT
for c in 64 240; do echo "=== C=$c"; timeout -k 5 60 python tools/sa_each.py $c 2>&1 | grep -v Warning | grep -A1 "branch" | grep done | tr '\n' ' '; echo; done
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing
G4D_SA_NSLOT=2 timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing
G4D_SA_NSLOT=1 timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing
timeout -k 10 300 python bench.py --no-cpu-baseline --no-train --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02k_bench.json")); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), d.get("label_agreement"), d.get("cube"))
for k in d["kernels"]:
    print("    %-70s %.4f" % (k["name"][:70], k["ms"]))
P
