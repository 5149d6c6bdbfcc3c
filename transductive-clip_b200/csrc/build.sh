#!/usr/bin/env bash
# Builds libtclip_b200.so in-tree for sm_100a (cross-compiles without a GPU).  Usage: csrc/build.sh [-j N]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../tclip_b200/libtclip_b200.so"
OBJ="$HERE/build"
mkdir -p "$OBJ"
NVCC="${NVCC:-nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v)
pids=()
for src in "$HERE"/*.cu; do
  obj="$OBJ/$(basename "${src%.cu}").o"
  if [[ ! -f "$obj" || "$src" -nt "$obj" || "$HERE/tclip_math.cuh" -nt "$obj" || "$HERE/tclip_kernels.cuh" -nt "$obj" || "$HERE/../../include/tclip_b200.h" -nt "$obj" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" > "$obj.log" 2>&1 || { cat "$obj.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "$OBJ"/*.o -lcudart
echo "built $OUT"
