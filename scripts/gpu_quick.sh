#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_quick.txt
cat gpurun_out/pytest_gpu_quick.txt
timeout 600 python scripts/gpu_phase_times.py --skip-only 2>&1 | tee gpurun_out/phase_quick.txt
