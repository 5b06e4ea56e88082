#!/usr/bin/env bash
set -u
TAG=${1:-r02g}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/${TAG}_pytest_gpu.log
( timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing
  timeout -k 10 120 python tools/sa_timeline.py 1 3 5 2>&1 | tail -45 ) > $OUT/${TAG}_sa.txt 2>&1
head -1 $OUT/${TAG}_sa.txt
( G4D_SA_NSLOT=1 timeout -k 10 120 python tools/sa_timeline.py 1 5 2>&1 | tail -30 ) > $OUT/${TAG}_sa_ns1.txt 2>&1
for v in "rows:G4D_FPS=rows" "prunedmorton:G4D_FPS=rows,G4D_FPS_WS=pruned" "pruned:G4D_FPS=pruned"; do
  name="${v%%:*}"; envs="${v#*:}"
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    timeout -k 10 300 python bench.py --no-cpu-baseline --no-train --no-extras --steps 10 --warmup 3 > $OUT/${TAG}_bench_${name}.json 2>> $OUT/${TAG}_bench.err )
  echo "bench $name exit $?"
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02g_bench_*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
        for k in d["kernels"]:
            if any(s in k["name"] for s in ("fps", "sa_mlp", "SA stack")): print("    %-70s %.4f" % (k["name"][:70], k["ms"]))
    except Exception as e:
        print(f, "unreadable", e)
P
