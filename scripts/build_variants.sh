#!/usr/bin/env bash
# Diagnostic builds of libtclip_b200 (error attribution / tuning sweeps): scripts/build_variants.sh NAME "-Dflags" [NAME "-Dflags" ...]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/../transductive-clip_b200/csrc"
OUT="$HERE/../gpurun_variants"
mkdir -p "$OUT"
while [[ $# -ge 2 ]]; do
  name="$1"; defs="$2"; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared $defs \
      "$SRC"/capi.cu "$SRC"/dirichlet_mm.cu "$SRC"/dirichlet_estep.cu "$SRC"/probe.cu -o "$OUT/libtclip_$name.so" -lcudart 2>&1 | grep -v deprecated || true ) &
done
wait
ls -la "$OUT"
