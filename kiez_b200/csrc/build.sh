#!/bin/bash
# Builds kiez_b200/lib/libkiez_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${KB2_OUT_DIR:-$HERE/../lib}"
OBJ="${KB2_OBJ_DIR:-$HERE/../../build/obj}"
SRC="${KB2_SRC_DIR:-$HERE}"
mkdir -p "$OUT" "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr ${KB2_NVCC_EXTRA:-})
pids=()
for f in api prep knn_simt knn_tc knn_tc2 knn_fused knn_screen refine rescale analysis; do
  if [ ! -f "$OBJ/$f.o" ] || [ "$SRC/$f.cu" -nt "$OBJ/$f.o" ] || \
     [ -n "$(find "$SRC" "$HERE/../../include" -name '*.cuh' -newer "$OBJ/$f.o" -o -name '*.h' -newer "$OBJ/$f.o" | head -1)" ]; then
    "$NVCC" "${FLAGS[@]}" -I"$HERE/../../include" -c "$SRC/$f.cu" -o "$OBJ/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT/libkiez_b200.so" "$OBJ"/{api,prep,knn_simt,knn_tc,knn_tc2,knn_fused,knn_screen,refine,rescale,analysis}.o
echo "built $OUT/libkiez_b200.so"
