#!/bin/bash
# round 2, call j: one-kernel zero-copy host step; the whole bench line with the driver's arguments
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reset_and_scale.py tests/test_gpu_goals.py -m gpu -q -x > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r02j_pytest.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02j_bench_driver.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r02j_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02j_bench_driver.json'))
print('value %.4g frac %.4f best %.4f e2e %.4g (full copy %.4g) launches %d' % (d['value'], d['roofline']['frac'], d['roofline']['frac_best_segment'], d['e2e']['value'], d['e2e']['full_copy_value'], d['gpu_launches']))
for k, v in d['configs'].items():
    print(k, 'frac %.4f value %.4g' % (v['frac'], v['value']), v['episode_stats']['resets_inline'])
print(json.dumps(d['cpu_baseline'])[:600])
PY
timeout 300 python tools/probe_e2e.py > gpurun_out/r02j_probe_e2e.txt 2>&1; tail -n 12 gpurun_out/r02j_probe_e2e.txt
