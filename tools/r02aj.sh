#!/usr/bin/env bash
timeout -k 10 300 python tools/nn_sweep.py 2>&1 | tail -20
