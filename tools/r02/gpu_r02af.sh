#!/bin/bash
# round 2, call af: walls (CrlConfig.walled, ABI 7): GPU tests incl. the walls_* fixtures; whole suite; bench line with the encoder leg
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_walls.py -m gpu -q -s > gpurun_out/r02af_walls.log 2>&1; echo "walls rc=$?"; grep -v "^$" gpurun_out/r02af_walls.log | tail -n 12
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02af_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02af_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02af_bench_driver.json 2> gpurun_out/r02af_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02af_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02af_bench_driver.json'))
print('value %.4g frac %.4f best %.4f e2e %.4g pageable %.4g launches %d' % (d['value'], d['roofline']['frac'], d['roofline']['frac_best_segment'], d['e2e']['value'], d['e2e']['pageable_actions_value'], d['gpu_launches']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f value %.4g e2e %.4g' % (v['frac'], v['value'], v.get('e2e_value', 0)), v['episode_stats']['resets_inline'])
print('encoder', json.dumps(d.get('encoder'))[:600])
PY
