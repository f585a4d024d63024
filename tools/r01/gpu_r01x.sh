#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python tools/bench_rollout.py > gpurun_out/x_rollout.jsonl 2> gpurun_out/x_err.log; echo "rc=$?"; cat gpurun_out/x_rollout.jsonl; tail -3 gpurun_out/x_err.log
timeout 300 python bench.py --env PointTSP-v3 --envs 65536 --no-cpu-baseline --e2e-steps 20 --steps 16000 --warmup 1600 > gpurun_out/x_tsp_v3.json 2>>gpurun_out/x_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/x_tsp_v3.json').read().strip().splitlines()[-1])
print('PointTSP-v3 (EXT kernel)', d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, d['roofline']['bytes_per_env_step'], d['e2e']['value'])
PY
