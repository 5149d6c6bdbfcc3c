#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_quick.txt
cat gpurun_out/pytest_gpu_quick.txt
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('value',round(d['value']),round(d['ms_per_step'],2),'serial',round(d['serial']['value']),'| e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2),'serial',round(d['e2e']['serial']['value']),'| frac',round(d['roofline']['frac'],4),d['roofline']['element_updates_per_s'],'launches',d['gpu_launches'])
PY
