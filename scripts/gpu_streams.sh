#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
run() {  # resident streams steps
  TCLIP_MM_RESIDENT=$1 timeout 600 python bench.py --streams $2 --steps $3 --warmup 3 --no-cpu-baseline > gpurun_out/bs.json 2> gpurun_out/bs.err || tail -5 gpurun_out/bs.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bs.json').read().strip().splitlines()[-1])
print('resident $1 streams',d['config']['streams'],'steps $3 value',round(d['value']),round(d['ms_per_step'],2),'serial',round(d['serial']['value']),'| e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2),'serial',round(d['e2e']['serial']['value']),'frac',round(d['roofline']['frac'],3))
PY
}
run 6 4 8
run 4 4 8
run 3 4 8
run 4 6 12
run 4 8 16
run 3 6 12
run 4 3 5
run 4 4 5
