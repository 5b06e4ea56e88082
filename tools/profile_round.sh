#!/usr/bin/env bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/profile_round.sh <tag>'): GPU parity tests, smoke, the bench line (+ reference
# arm), the ncu launch list of the bench command and `ncu --set full` captures of every hot-path kernel.  Outputs: gpurun_out/.
# Every step has a hard timeout: a hung profiler pass must not eat the GPU budget.
set -u
TAG="${1:-r01}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout -k 10 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  tail -4 $OUT/${TAG}_pytest_gpu.log
  timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
  tail -2 $OUT/${TAG}_smoke.log
fi
if [ "${SKIP_BENCH:-0}" != 1 ]; then
timeout -k 10 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
fi
if [ "${SKIP_REF:-0}" != 1 ]; then
  timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
fi
if [ "${SKIP_NCU:-0}" != 1 ]; then
  # full sections, one launch of every kernel of libgarment4d_b200.so at c3 sizes -- all but fp_interp_mlp first ...
  timeout -k 10 420 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k regex:'g4d::(?!fp_interp_mlp)' -f -o $OUT/${TAG}_full python tools/ncu_once.py c3 > $OUT/${TAG}_full_run.log 2>&1
  echo "ncu full exit $?"
  # gpurun brings back at most 64 MiB: keep the CSV pages and per-kernel hot spots of the big report, not the report itself
  ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_full.ncu-rep --page source --csv > $OUT/${TAG}_full_source.csv 2>/dev/null
  python tools/ncu_hotspots.py $OUT/${TAG}_full_source.csv 25 > $OUT/${TAG}_full_hotspots.txt 2>&1
  rm -f $OUT/${TAG}_full.ncu-rep $OUT/${TAG}_full_source.csv
  # ... then fp_interp_mlp on its own (a capture of its first warp-specialised version never returned)
  timeout -k 10 ${FP_NCU_TIMEOUT:-120} ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k regex:'g4d::fp_interp_mlp' -f -o $OUT/${TAG}_full_fp python tools/ncu_once.py c3 > $OUT/${TAG}_full_fp_run.log 2>&1
  echo "ncu fp_interp_mlp exit $?"
  # launch list of the bench command (device time per launch; cold-cache, serialised: shares, not absolutes)
  # (kernel by kernel, --no-graph: Nsight Compute cannot launch fp_mlp2_kernel from inside the captured graph)
  timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --config c3 --chunks ${LAUNCH_CHUNKS:-4} --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-train --no-extras > $OUT/${TAG}_launches_run.log 2>&1
  echo "ncu launches exit $?"
  python tools/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.md 2>&1
  gzip -f $OUT/${TAG}_launches.csv
fi
ls -la $OUT | tail -20
