#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
PT_K=100 PT_T=100 timeout 300 python scripts/gpu_phase_times.py --skip-only 2>&1 | grep -A2 "hard=False" | head -3
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('value',round(d['value']),round(d['ms_per_step'],2),'serial',round(d['serial']['value']),'| e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2),'serial',round(d['e2e']['serial']['value']),'| frac',round(d['roofline']['frac'],4))
PY
