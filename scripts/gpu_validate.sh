#!/usr/bin/env bash
# Round-end validation on one B200 (run through gpurun): GPU tests, smoke, the default bench line and the k-means lines.
# Every step under its own timeout; results under gpurun_out/validate_*.
set -u
tag="${1:-v}"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/validate_${tag}_em.json 2> gpurun_out/validate_${tag}_em.err
tail -c 400 gpurun_out/validate_${tag}_em.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.loads(open(f"gpurun_out/validate_{tag}_em.json").read().strip().splitlines()[-1])
print("em", d["value"], d["e2e"]["value"], d["ms_per_step"], d["serial"], d["roofline"]["frac"], d["roofline"].get("kernel_frac"),
      d["gpu_launches"], d["clocks"], d["cpu_baseline"]["value"])
PY
for m in soft gauss hardkm; do
  timeout 200 python bench.py --method $m --steps 24 --warmup 8 > gpurun_out/validate_${tag}_$m.json 2> /dev/null
  python - "$tag" "$m" <<'PY'
import json, sys
tag, m = sys.argv[1:3]
d = json.loads(open(f"gpurun_out/validate_{tag}_{m}.json").read().strip().splitlines()[-1])
print(m, d["value"], d["e2e"]["value"], d["ms_per_step"], d["serial"]["ms_per_step"], d["roofline"]["frac"],
      d["roofline"]["loop_ms_serial"], d["cpu_baseline"]["value"])
PY
done
