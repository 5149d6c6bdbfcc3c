#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_quick.txt
cat gpurun_out/pytest_gpu_quick.txt
timeout 900 python bench.py > gpurun_out/bench9.json 2> gpurun_out/bench9.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench9.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'frac',d['roofline']['frac'],'launches',d['gpu_launches'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_skip3.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_skip3.log 2>&1
ls -la gpurun_out/launches_skip3.csv
