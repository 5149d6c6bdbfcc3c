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
echo "== full captures: mm_chunk_kernel (dense), mm_spec_kernel, moments_kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mm_chunk_kernel -s 2 -c 1 -o gpurun_out/prof_mm_final2 \
  python scripts/gpu_probe2.py > gpurun_out/ncu_full_mm_final2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mm_spec_kernel -s 25 -c 1 -o gpurun_out/prof_spec_final \
  python scripts/gpu_phase_times.py --skip-only > gpurun_out/ncu_full_spec_final.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:moments_kernel -s 0 -c 1 -o gpurun_out/prof_moments_final \
  python scripts/gpu_phase_times.py --skip-only > gpurun_out/ncu_full_moments_final.log 2>&1
ls -la gpurun_out/*final*.ncu-rep
