#!/usr/bin/env bash
# few-rows M-step in one launch (mm_spec_kernel): warps-per-row variants, then one default bench line
set -uo pipefail
mkdir -p gpurun_out
for v in 8 4; do
  echo "== TCLIP_SPEC_W=$v" | tee -a gpurun_out/phase_spec2.txt
  TCLIP_SPEC_W=$v timeout 600 python scripts/gpu_phase_times.py --skip-only 2>&1 | tee -a gpurun_out/phase_spec2.txt
done
timeout 900 python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err
tail -c 2500 gpurun_out/bench7.json
