#!/bin/bash
# round 2, call bp: result records cross the link only when they change (host-direct step): the whole GPU suite, then the driver's bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02bp_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02bp_pytest.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02bp_bench.json 2> gpurun_out/r02bp_err.log; echo "bench rc=$?"; tail -n 2 gpurun_out/r02bp_err.log
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02bp_bench.json').read().strip().splitlines()[-1])
e = d['e2e']
print('value %.4g frac %.4f e2e %.4g best %.4g floor %.4g pageable %.4g d2h %d' % (d['value'], d['roofline']['frac'], e['value'], e['best_group_value'], e['copy_engine_floor_value'], e['pageable_actions_value'], e['d2h_bytes_per_step']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f e2e %s' % (v['frac'], v.get('e2e_value')))
print(d['clocks'])
PY
