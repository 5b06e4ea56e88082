#!/usr/bin/env bash
# Runs ON THE GPU BOX (one GPU): round-2 first call.
set -u
TAG=r02a; OUT=gpurun_out; mkdir -p $OUT
bash tools/repro_seeds.sh $TAG
# experimental paths that were committed in round 1 without a GPU run
G4D_NN_COOP=1 timeout -k 5 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "three_nn or grid" > $OUT/${TAG}_nncoop_pytest.log 2>&1
echo "nn_coop pytest exit $?"; tail -3 $OUT/${TAG}_nncoop_pytest.log
for v in "base:" "occ:G4D_FPS_OCC=1" "prio:G4D_CHUNK_PRIORITY=1" "occprio:G4D_FPS_OCC=1,G4D_CHUNK_PRIORITY=1" "nncoop:G4D_NN_COOP=1" "lt:G4D_FP_LT_EPILOGUE=1"; do
  name="${v%%:*}"; envs="${v#*:}"
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    timeout -k 10 200 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > $OUT/${TAG}_bench_${name}.json 2>> $OUT/${TAG}_bench.err )
  echo "bench $name exit $?"
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02a_bench_*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3))
    except Exception as e:
        print(f, "unreadable", e)
P
