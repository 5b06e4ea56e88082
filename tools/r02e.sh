#!/usr/bin/env bash
set -u
TAG=${1:-r02e}; OUT=gpurun_out; mkdir -p $OUT
for v in "base:" "spin:G4D_SA_SPIN=1" "ns1:G4D_SA_NSLOT=1" "ns1spin:G4D_SA_NSLOT=1,G4D_SA_SPIN=1"; do
  name="${v%%:*}"; envs="${v#*:}"
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    echo "== $name"; timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing
    timeout -k 10 120 python tools/sa_timeline.py 1 3 5 2>&1 | tail -45 ) > $OUT/${TAG}_${name}.txt 2>&1
  head -2 $OUT/${TAG}_${name}.txt
done
