#!/usr/bin/env bash
# Round-1 profiling pass 2: parity, one bench line, launch list of one skip-dead step, full captures of the tcgen05
# contraction, the few-rows M-step and the dense M-step kernel.
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_r1d.txt
cat gpurun_out/pytest_gpu_r1d.txt
timeout 900 python bench.py > gpurun_out/bench8.json 2> gpurun_out/bench8.err
tail -c 1800 gpurun_out/bench8.json
echo "== launch list (skip_dead step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_skip2.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_skip2.log 2>&1
echo "== full capture: logits_tc_kernel (imagenet-shape batch)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logits_tc_kernel -s 1 -c 1 -o gpurun_out/prof_logits_tc \
  python scripts/gpu_contraction_one.py > gpurun_out/ncu_full_tc.log 2>&1
echo "== full capture: mm_spec_kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mm_spec_kernel -s 12 -c 1 -o gpurun_out/prof_mm_spec \
  python scripts/gpu_phase_times.py --skip-only > gpurun_out/ncu_full_spec.log 2>&1
echo "== full capture: mm_chunk_kernel on a full batch of rows"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mm_chunk_kernel -s 2 -c 1 -o gpurun_out/prof_mm_r1d \
  python scripts/gpu_probe2.py > gpurun_out/ncu_full_mm_r1d.log 2>&1
ls -la gpurun_out | tail -8
