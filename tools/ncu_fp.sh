#!/usr/bin/env bash
# ON THE GPU BOX: ncu --set full of the kernels matching $1 (regex on the demangled name) during one pass of tools/ncu_once.py
set -u
PAT="${1:-fp_interp_mlp}"; TAG="${2:-r01d}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k regex:"$PAT" -f -o gpurun_out/${TAG}_ncu python tools/ncu_once.py c3 > gpurun_out/${TAG}_ncu_run.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/${TAG}_ncu_run.log
