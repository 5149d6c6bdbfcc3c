#!/usr/bin/env bash
# Evidence that the built library is tcgen05 / TMEM / TMA / packed-FP32 code: opcode counts per kernel and short SASS
# excerpts around the tensor-core, TMEM-load, TMA and FFMA2 instructions.  Usage: scripts/sass_excerpts.sh > profiles/r2_sass_excerpts.txt
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SO="$HERE/transductive-clip_b200/tclip_b200/libtclip_b200.so"
TMP="$(mktemp)"
cuobjdump -sass "$SO" > "$TMP"
echo "# cuobjdump -sass transductive-clip_b200/tclip_b200/libtclip_b200.so (built by csrc/build.sh: nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)"
echo "# $(grep -c 'Function :' "$TMP") kernels; opcode totals over the whole library:"
for op in UTCHMMA LDTM UTMALDG UTCBAR SYNCS LDGSTS FFMA2 FADD2 FMUL2 MUFU DFMA DADD REDUX; do
  printf "#   %-8s %6d\n" "$op" "$(grep -c "[ .]$op" "$TMP" || true)"
done
echo
echo "## per kernel (one instance per family): instructions, tensor-core MMA (UTCHMMA), TMEM load (LDTM), TMA load (UTMALDG),"
echo "## MMA-completion barrier (UTCBAR), packed fp32 (FFMA2/FADD2/FMUL2), MUFU, fp64 (DFMA/DADD)"
awk '/Function :/{name=$3} /UTCHMMA/{a[name]++} /LDTM/{b[name]++} /UTMALDG/{c[name]++} /UTCBAR/{d[name]++} /FFMA2|FADD2|FMUL2/{e[name]++} /MUFU/{f[name]++} /DFMA|DADD/{g[name]++} /^ +\/\*[0-9a-f]+\*\/ /{n[name]++}
     END{for(k in n) printf "%6d %3d %3d %2d %2d %5d %4d %4d %s\n", n[k], a[k], b[k], c[k], d[k], e[k], f[k], g[k], k}' "$TMP" \
  | grep -E "logits_tc_kernel|mm_chunk_kernelILi16E|mm_chunk_kernelILi2ELb0|mm_spec_kernelILi4ELi4ELb1ELb0|kproj_iter_kernelILi5ELi5ELi128ELb1ELi256ELb1|gram_kernel|chol_kernel|match_clusters|moments_kernel|lognorm_kernel|softmax_reg|pair_kernelILi1E|centroids_kernel|probe_ffma2|expand_kernel" \
  | while read n a b c d e f g name; do printf "%6d instr  UTCHMMA %3d  LDTM %3d  UTMALDG %2d  UTCBAR %2d  packed-fp32 %5d  MUFU %4d  fp64 %4d  %s\n" "$n" "$a" "$b" "$c" "$d" "$e" "$f" "$g" "$(echo "$name" | cu++filt | sed 's/tclip::(anonymous namespace):://; s/(.*//')"; done | sort -k13
echo
echo "## logits_tc_kernel<true> (contraction_tc.cu): the TMA loads, the tcgen05 MMAs (kind::tf32 -> UTCHMMA), their completion barriers and the TMEM drains"
awk '/Function :.*logits_tc_kernelILb1E/{p=1; next} /Function :/{p=0} p' "$TMP" | grep -E "UTMALDG|UTCHMMA|UTCBAR|LDTM" | sed 's/^ *//' | head -48
echo
echo "## mm_chunk_kernel<16, false> (dirichlet_mm.cu): a stretch of the MM update loop (packed FFMA2/FMUL2/FADD2 + MUFU)"
awk '/Function :.*mm_chunk_kernelILi16ELb0E/{p=1; next} /Function :/{p=0} p' "$TMP" | grep -E "FFMA2|MUFU|FMUL2|FADD2" | sed -n '200,236p' | sed 's/^ *//'
echo
echo "## kproj_iter_kernel<5, 5, 128, CHAIN, 256, TRI> (kmeans_run.cu): the asynchronous tile copies (cp.async -> LDGSTS) and the end of the"
echo "## distance loop (FADD2 with a broadcast scalar operand + FFMA2 on class pairs)"
awk '/Function :.*kproj_iter_kernelILi5ELi5ELi128ELb1ELi256ELb1/{p=1; next} /Function :/{p=0} p' "$TMP" | grep -E "LDGSTS|LDGDEPBAR|DEPBAR" | sed 's/^ *//' | head -8
awk '/Function :.*kproj_iter_kernelILi5ELi5ELi128ELb1ELi256ELb1/{p=1; next} /Function :/{p=0} p' "$TMP" | grep -E "FFMA2|FADD2" | sed 's/^ *//' | tail -32
rm -f "$TMP"
