#!/usr/bin/env bash
# Runs ON THE GPU BOX: whole GPU suite (hard timeout), then bench variants given as "NAME:ENV=VAL,ENV=VAL" arguments.
#   bash tools/gpu_check.sh r01e base: conv:G4D_FP_GEMM=conv
set -u
TAG="${1:-r01x}"; shift
OUT=gpurun_out
mkdir -p $OUT
timeout -k 10 600 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1
rc=$?; echo "pytest exit $rc"; tail -6 $OUT/${TAG}_pytest_gpu.log
if [ $rc -ne 0 ]; then exit $rc; fi
for v in "$@"; do
  name="${v%%:*}"; envs="${v#*:}"
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    timeout -k 10 400 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench_${name}.json 2>> $OUT/${TAG}_bench.err )
  echo "bench $name exit $?"
done
python - "$TAG" <<'P'
import json, glob, sys
for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_bench_*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
        for k in d["kernels"]:
            if any(s in k["name"] for s in ("FP", "fp_", "three_nn", "SA stack")): print("    %-70s %.4f" % (k["name"][:70], k["ms"]))
    except Exception as e:
        print(f, "unreadable", e)
P
