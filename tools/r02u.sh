#!/usr/bin/env bash
set -u
G4D_FPS_PROF=1 timeout -k 10 120 python tools/fps_phases.py 120 2>&1 | tail -50
