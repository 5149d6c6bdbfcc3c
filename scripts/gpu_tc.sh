#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_tc.txt
cat gpurun_out/pytest_gpu_tc.txt
timeout 300 python scripts/gpu_contraction.py > gpurun_out/contraction_tc2.txt 2>&1
echo "rc=$?" >> gpurun_out/contraction_tc2.txt
tail -9 gpurun_out/contraction_tc2.txt
for v in "4 1" "4 0" "2 0"; do
  set -- $v
  echo "== TCLIP_SPEC_W=$1 TCLIP_SPEC_PIPE=$2" | tee -a gpurun_out/phase_spec3.txt
  TCLIP_SPEC_W=$1 TCLIP_SPEC_PIPE=$2 timeout 600 python scripts/gpu_phase_times.py --skip-only 2>&1 | grep -A3 "hard=False" | tee -a gpurun_out/phase_spec3.txt
done
