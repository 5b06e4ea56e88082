#!/usr/bin/env bash
set -u
timeout -k 10 400 python -m pytest tests/test_sa_mlp_gpu.py -x -q -m gpu 2>&1 | tail -3
for ks in 32 64; do
echo "ks=$ks"; G4D_MLP2_KS=$ks G4D_MLP2_PROF=1 timeout -k 10 200 python tools/fp2_once.py 240 2>&1 | grep -A2 "^ok" | cut -c1-230
G4D_MLP2_KS=$ks timeout -k 10 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fp_mlp2' python tools/fp2_once.py 240 2>&1 | grep -E "gpu__time" | head -4
done
