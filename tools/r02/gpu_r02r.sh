#!/bin/bash
# round 2, call r: whole GPU suite after the EXT-kernel test extensions; smoke; bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/r02r_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02r_bench_driver.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02r_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02r_bench_driver.json'))
print('value %.4g frac %.4f best %.4f e2e %.4g pageable %.4g (full copy %.4g) launches %d' % (d['value'], d['roofline']['frac'], d['roofline']['frac_best_segment'], d['e2e']['value'], d['e2e']['pageable_actions_value'], d['e2e']['full_copy_value'], d['gpu_launches']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f value %.4g e2e %.4g' % (v['frac'], v['value'], v.get('e2e_value', 0)), v['episode_stats']['resets_inline'])
PY
