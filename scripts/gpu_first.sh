#!/usr/bin/env bash
# First GPU pass: parity tests, smoke, bench (both MM schedules), ncu launch list + one full capture of the M-step kernel.
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "Executing" | tail -5 | tee gpurun_out/smoke.txt
echo "== bench skip_dead" ; timeout 900 python bench.py 2> gpurun_out/bench_skip.err | tee gpurun_out/bench_skip.json
echo "== bench dense" ; timeout 900 python bench.py --mm-mode dense --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/bench_dense.err | tee gpurun_out/bench_dense.json
echo "== bench hard" ; timeout 900 python bench.py --method hard --no-cpu-baseline 2> gpurun_out/bench_hard.err | tee gpurun_out/bench_hard.json
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mm_chunk -s 25 -c 2 -o gpurun_out/prof_mm \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --mm-mode dense > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
