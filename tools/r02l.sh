#!/usr/bin/env bash
set -u
TAG=${1:-r02l}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 300 python -m pytest tests/test_sa_mlp_gpu.py -x -q -m gpu -k "fp_module or fused_fp0 or runner" > $OUT/${TAG}_pytest_fp.log 2>&1; echo "pytest fp exit $?"; tail -8 $OUT/${TAG}_pytest_fp.log
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/${TAG}_pytest_gpu.log
timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing
for v in "tc:" "half:G4D_FP_GEMM=half"; do
  name="${v%%:*}"; envs="${v#*:}"
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    timeout -k 10 300 python bench.py --no-cpu-baseline --no-train --no-extras --steps 10 --warmup 3 > $OUT/${TAG}_bench_${name}.json 2>> $OUT/${TAG}_bench.err )
  echo "bench $name exit $?"
done
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02l_bench_*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"], 3), d["gpu_launches"])
        for k in d["kernels"]:
            if any(s in k["name"] for s in ("FP", "fp_", "three_nn", "sa_mlp_max,", "stack")): print("    %-70s %.4f" % (k["name"][:70], k["ms"]))
    except Exception as e:
        print(f, "unreadable", e)
P
