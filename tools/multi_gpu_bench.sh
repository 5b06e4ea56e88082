#!/usr/bin/env bash
# Runs ON THE GPU BOX with N GPUs: the multi-GPU bench exactly as the driver launches it, with NCCL logging. usage: multi_gpu_bench.sh TAG N [extra bench args]
set -u
TAG="${1:-mgpu}"; N="${2:-2}"; shift; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
export NCCL_DEBUG=INFO NCCL_DEBUG_FILE=$PWD/$OUT/${TAG}_nccl_%h_%p.log TORCH_NCCL_TRACE_BUFFER_SIZE=2000 TORCH_NCCL_DUMP_ON_TIMEOUT=1
timeout -k 10 ${BENCH_TIMEOUT:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 "$@" > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err
echo "bench N=$N exit $?"
grep -h "reached barrier\|process groups\|train arm" $OUT/${TAG}_bench_n${N}.err | tail -12
tail -5 $OUT/${TAG}_bench_n${N}.err
grep -h -i "nvls\|nranks\|channels\|Connected all" $OUT/${TAG}_nccl_*.log 2>/dev/null | sort | uniq -c | sort -rn | head -12
python - "$OUT/${TAG}_bench_n${N}.json" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "n_gpus", d["n_gpus"], d["scaling"])
    print("cube", d.get("cube")); print("train", d.get("train"))
except Exception as e:
    print("no bench line:", e)
P
# keep the NCCL logs small
for f in $OUT/${TAG}_nccl_*.log; do head -c 60000 "$f" > "$f.head"; rm -f "$f"; done
