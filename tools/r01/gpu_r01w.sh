#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/w_pytest.log
timeout 400 python bench.py > gpurun_out/w_default.json 2>gpurun_out/w_default.err; echo "bench rc=$?"; wc -l gpurun_out/w_default.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/w_default.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'])
PY
