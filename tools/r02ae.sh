#!/usr/bin/env bash
set -u
echo "default"; timeout -k 10 200 python tools/fps_bench.py 2>&1 | grep "C=120\|C=240\|C=148"
echo "wide"; G4D_FPS_WIDE=1 timeout -k 10 200 python tools/fps_bench.py 2>&1 | grep "C=120\|C=240\|C=148"
