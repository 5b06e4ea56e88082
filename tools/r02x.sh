#!/usr/bin/env bash
set -u
timeout -k 10 600 python -m pytest tests/test_garment_lbs_gpu.py -x -q -m gpu 2>&1 | tail -30
