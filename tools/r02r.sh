#!/usr/bin/env bash
set -u
for v in "auto:" "narrow:G4D_FPS_WIDE=0" "wide:G4D_FPS_WIDE=1" "rows:G4D_FPS_WS=rows"; do
  name="${v%%:*}"; envs="${v#*:}"
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    echo "== $name"; timeout -k 10 120 python tools/fps_bench.py 2>&1 | grep "C=" | tr '\n' ';'; echo )
done
timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing
python tools/fp2_once.py 240 | tail -2
timeout -k 10 300 python bench.py --no-cpu-baseline --no-train --no-extras --steps 10 --warmup 3 > gpurun_out/r02r_bench.json 2>> gpurun_out/r02r_bench.err; echo "bench exit $?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r02r_bench.json")); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
for k in d["kernels"]:
    print("    %-70s %.4f" % (k["name"][:70], k["ms"]))
P
