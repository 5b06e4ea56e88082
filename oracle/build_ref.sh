#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own pointnet2 kernels -- the four .cu files where
# they lie under /root/reference, unmodified, with the reference's own nvcc
# optimisation level (-O2, modules/pointnet2/pointnet2/setup.py:19-20) -- for
# sm_100a, behind the extern "C" shim in ref_shim.cu.  Output goes ONLY into
# oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
# The reference's setup.py / .cpp wrappers are not used: they need THC/THC.h,
# which torch 2.11 no longer ships.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${G4D_REFERENCE_ROOT:-/root/reference}/modules/pointnet2/pointnet2/src"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
    echo "build_ref.sh: $REF not present (GPU box?) -- using prebuilt $OUT if any" >&2
    exit 0
fi
mkdir -p "$OUT"
# The reference's own Python operator layer (pointnet2_utils.py, pointnet2_modules.py, pytorch_utils.py), unmodified, as ONE
# archive next to its kernels: tests/test_reference_layer_gpu.py puts the archive on sys.path (zipimport) ON TOP OF the repository's
# pointnet2_cuda.py to show the drop-in boundary holds.  The GPU box has no /root/reference; like the .so, the archive is a build
# product of the reference made by this recipe: test infrastructure, git-ignored, never part of the product, no loose copies.
rm -rf "$OUT/pointnet2"
"${PYTHON:-python}" - "$REF/.." "$OUT/pointnet2_reference_layer.zip" <<'EOF'
import os, sys, zipfile
src, dst = sys.argv[1], sys.argv[2]
with zipfile.ZipFile(dst, "w", zipfile.ZIP_DEFLATED) as z:
    z.writestr("pointnet2/__init__.py", "")
    for f in ("pointnet2_utils.py", "pointnet2_modules.py", "pytorch_utils.py"):
        z.write(os.path.join(src, f), "pointnet2/" + f)
EOF
if [ "$OUT/libpointnet2_ref.so" -nt "$HERE/ref_shim.cu" ] && [ "${1:-}" != "--force" ]; then
    echo "build_ref.sh: $OUT/libpointnet2_ref.so up to date"; exit 0
fi
PY="${PYTHON:-python}"
TORCH_INC="$($PY - <<'EOF'
import os, torch, sysconfig
base = os.path.join(os.path.dirname(torch.__file__), "include")
print(" ".join("-I" + p for p in (base, os.path.join(base, "torch/csrc/api/include"),
                                  sysconfig.get_paths()["include"])))
EOF
)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -w"
pids=()
for f in sampling_gpu ball_query_gpu group_points_gpu interpolate_gpu; do
    $NVCC $FLAGS -I"$REF" $TORCH_INC -c "$REF/$f.cu" -o "$OUT/$f.o" &
    pids+=($!)
done
$NVCC $FLAGS -c "$HERE/ref_shim.cu" -o "$OUT/ref_shim.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -shared -o "$OUT/libpointnet2_ref.so" "$OUT"/sampling_gpu.o "$OUT"/ball_query_gpu.o \
    "$OUT"/group_points_gpu.o "$OUT"/interpolate_gpu.o "$OUT"/ref_shim.o
rm -f "$OUT"/*.o
echo "build_ref.sh: built $OUT/libpointnet2_ref.so"
