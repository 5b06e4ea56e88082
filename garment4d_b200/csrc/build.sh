#!/usr/bin/env bash
# Builds garment4d_b200/libgarment4d_b200.so (all kernels + the C-ABI) for sm_100a, in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libgarment4d_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden
       -Xptxas -v -I"$HERE/../../include")
mkdir -p "$HERE/obj"
pids=()
for f in "$HERE"/*.cu; do
    o="$HERE/obj/$(basename "${f%.cu}").o"
    stale=0
    for h in "$HERE"/*.cuh "$HERE"/../../include/*.h "$HERE/build.sh"; do [ "$h" -nt "$o" ] && stale=1; done
    if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ $stale = 1 ] || [ "${1:-}" = "--force" ]; then
        ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; exit 1; } ) &
        pids+=($!)
    fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "$HERE"/obj/*.o
echo "built $OUT"
