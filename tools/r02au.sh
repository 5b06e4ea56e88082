#!/usr/bin/env bash
timeout -k 10 900 python bench.py > gpurun_out/r02au_bench.json 2> gpurun_out/r02au_bench.err; echo "bench exit $?"
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02au_bench_reference.json 2>> gpurun_out/r02au_bench.err
