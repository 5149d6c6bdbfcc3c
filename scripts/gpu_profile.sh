#!/usr/bin/env bash
# Round-1 profiling pass (B200_PROFILING.md recipe): launch list of one bench step + one full capture of the dominant kernel.
set -uo pipefail
mkdir -p gpurun_out
echo "== launch list (skip_dead step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_skip.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_skip.log 2>&1
echo "== launch list (dense step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_dense.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --mm-mode dense > gpurun_out/ncu_launch_dense.log 2>&1
echo "== full capture of mm_chunk_kernel on a full batch of rows"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mm_chunk_kernel -s 2 -c 1 -o gpurun_out/prof_mm_final \
  python scripts/gpu_probe2.py > gpurun_out/ncu_full_final.log 2>&1
ls -la gpurun_out | tail -8
