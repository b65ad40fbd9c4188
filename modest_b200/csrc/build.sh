#!/bin/bash
# Build libmodest_b200.so in-tree for sm_100a (B200).  Usage: modest_b200/csrc/build.sh [-v]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libmodest_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr --expt-extended-lambda"
[ "$1" = "-v" ] && FLAGS="$FLAGS -Xptxas -v"
mkdir -p "$HERE/obj"
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/obj/$(basename "${f%.cu}").o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ "$HERE/common.cuh" -nt "$o" ] || [ "$HERE/../../include/modest_b200.h" -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -shared -o "$OUT" "$HERE"/obj/*.o -gencode arch=compute_100a,code=sm_100a
echo "built $OUT"
