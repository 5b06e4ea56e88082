#!/usr/bin/env bash
set -u
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fps or furthest or sample" 2>&1 | tail -4
timeout -k 10 200 python tools/fps_bench.py 2>&1 | tail -12
G4D_FPS_PROF=1 timeout -k 10 120 python tools/fps_phases.py 120 2>&1 | grep -A12 "== body"
timeout -k 10 300 python tools/garment_lbs_bench.py 2>&1 | tail -3
