#!/usr/bin/env bash
set -u
timeout -k 10 200 python tools/fp_counters.py 240 2>&1 | tail -4
