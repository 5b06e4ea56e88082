#!/usr/bin/env bash
# ON THE GPU BOX: final bench line + reference arm, ncu --set full of the kernels matching $2 only, launch list of the bench command.
set -u
TAG="${1:-r01z}"; PAT="${2:-rows}"
OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 400 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
timeout -k 10 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
timeout -k 10 150 ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled -k regex:"$PAT" \
    -f -o $OUT/${TAG}_part python tools/ncu_once.py c3 > $OUT/${TAG}_part_run.log 2>&1
echo "ncu part exit $?"
ncu -i $OUT/${TAG}_part.ncu-rep --page raw --csv > $OUT/${TAG}_part_raw.csv 2>/dev/null
timeout -k 10 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown > $OUT/${TAG}_launches_run.log 2>&1
echo "ncu launches exit $?"
ls -la $OUT | grep $TAG
