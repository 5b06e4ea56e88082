#!/usr/bin/env bash
# GPU box: sa_mlp tests, knob sweep, ncu of the sa_mlp kernels
set -u
TAG=${1:-r02d}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 300 python -m pytest tests/test_sa_mlp_gpu.py -x -q -m gpu > $OUT/${TAG}_pytest_sa.log 2>&1; echo "pytest sa exit $?"; tail -3 $OUT/${TAG}_pytest_sa.log
for v in "base:" "ni1:G4D_SA_NI=1" "ns2:G4D_SA_NSLOT=2" "ns1:G4D_SA_NSLOT=1" "ns2ni1:G4D_SA_NSLOT=2,G4D_SA_NI=1"; do
  name="${v%%:*}"; envs="${v#*:}"
  ( IFS=','; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    echo "== $name"; timeout -k 10 120 python tools/ncu_sa.py 2>&1 | grep timing )
done
timeout -k 10 600 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k regex:'sa_mlp_max' -f -o $OUT/${TAG}_sa python tools/ncu_sa.py > $OUT/${TAG}_sa_run.log 2>&1
echo "ncu exit $?"
ncu -i $OUT/${TAG}_sa.ncu-rep --page raw --csv > $OUT/${TAG}_sa_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_sa.ncu-rep --page source --csv > $OUT/${TAG}_sa_source.csv 2>/dev/null
python tools/ncu_hotspots.py $OUT/${TAG}_sa_source.csv 30 > $OUT/${TAG}_sa_hotspots.txt 2>&1
rm -f $OUT/${TAG}_sa.ncu-rep $OUT/${TAG}_sa_source.csv
ls -la $OUT | grep ${TAG}
