#!/usr/bin/env bash
# End-of-round validation: GPU parity suite, smoke, default bench line, reference arm, launch list of one serial step.
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke_final.txt
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value',round(d['value']),round(d['ms_per_step'],2),'serial',round(d['serial']['value']),'| e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2),'serial',round(d['e2e']['serial']['value']),'| frac',round(d['roofline']['frac'],3),'launches',d['gpu_launches'],'cpu',d.get('cpu_baseline',{}).get('value'),d['clocks'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --streams 1 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
ls -la gpurun_out/launches_final.csv
