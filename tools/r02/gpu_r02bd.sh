#!/bin/bash
# round 2, call bd: HEAD check -- the whole GPU suite, smoke(), the driver's bench line and the reference arm
set -u
mkdir -p gpurun_out
t0=$(date +%s); timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02bd_pytest.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - t0 )) s"; tail -n 3 gpurun_out/r02bd_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02bd_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/r02bd_smoke.log
t0=$(date +%s); timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02bd_bench.json 2> gpurun_out/r02bd_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"; tail -n 2 gpurun_out/r02bd_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02bd_bench.json').read().strip().splitlines()[-1])
print('value %.4g frac %.4f e2e %.4g floor %.4g pageable %.4g launches %d' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['copy_engine_floor_value'], d['e2e']['pageable_actions_value'], d['gpu_launches']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f value %.4g e2e %s' % (v['frac'], v['value'], v.get('e2e_value')))
print({k: d['encoder'][k] for k in d.get('encoder', {}) if not isinstance(d['encoder'][k], (dict, list))} if 'encoder' in d else 'no encoder leg')
print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks'])
PY
t0=$(date +%s); timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r02bd_ref.json 2> gpurun_out/r02bd_ref.err; echo "ref rc=$? $(( $(date +%s) - t0 )) s"; cut -c1-200 gpurun_out/r02bd_ref.json
