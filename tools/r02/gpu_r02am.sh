#!/bin/bash
# round 2, call am: HEAD check -- whole GPU suite, smoke, the driver's bench command and the reference arm
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02am_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02am_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02am_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/r02am_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02am_bench_driver.json 2> gpurun_out/r02am_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02am_bench.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02am_bench_reference.json 2>> gpurun_out/r02am_bench.err; echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02am_bench_driver.json'))
print('value %.4g frac %.4f best %.4f e2e %.4g pageable %.4g (full copy %.4g) launches %d' % (d['value'], d['roofline']['frac'], d['roofline']['frac_best_segment'], d['e2e']['value'], d['e2e']['pageable_actions_value'], d['e2e']['full_copy_value'], d['gpu_launches']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f value %.4g e2e %.4g' % (v['frac'], v['value'], v.get('e2e_value', 0)), v['episode_stats']['resets_inline'])
e = d.get('encoder', {})
print('encoder', {k: (round(v, 1) if isinstance(v, float) else v) for k, v in e.items() if k in ('healthy', 'forward_us', 'zone_kernel_us', 'forward_from_state_us')}, e.get('roofline', {}).get('frac'))
print('cpu_baseline', d['cpu_baseline'].get('value'), d['cpu_baseline'].get('cores'))
r = json.load(open('gpurun_out/r02am_bench_reference.json'))
print('reference arm value %.4g cores %s' % (r['value'], r['cpu_baseline'].get('cores')))
PY
