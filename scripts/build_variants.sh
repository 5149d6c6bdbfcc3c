#!/usr/bin/env bash
# Diagnostic builds of libtclip_b200 with one MUFU approximation replaced by the correctly rounded op (error attribution).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="$HERE/../transductive-clip_b200/csrc"
OUT="$HERE/../gpurun_variants"
mkdir -p "$OUT"
for v in LG2 RCP SQRT ALL; do
  defs=(-DTCLIP_EXACT_$v)
  [[ $v == ALL ]] && defs=(-DTCLIP_EXACT_LG2 -DTCLIP_EXACT_RCP -DTCLIP_EXACT_SQRT)
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared "${defs[@]}" \
      "$SRC"/capi.cu "$SRC"/dirichlet_mm.cu "$SRC"/dirichlet_estep.cu "$SRC"/probe.cu -o "$OUT/libtclip_$v.so" -lcudart 2>&1 | grep -v deprecated || true ) &
done
wait
ls -la "$OUT"
