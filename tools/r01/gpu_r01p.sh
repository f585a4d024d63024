#!/bin/bash
# register-resident background sampler: TimedTSP with prefetch every 8 / 32 / 64 steps; other configs
set -u
mkdir -p gpurun_out
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s=d['episode_stats']; e=d['e2e']
    print("%-44s %.3e frac %.3f %.2f us/step pf %d inl %d" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
timeout 600 python -m pytest tests/test_gpu_reset_and_scale.py -x -q -m gpu > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.log
for v in "pe8:--prefetch-every 8" "pe32:--prefetch-every 32" "pe64:--prefetch-every 64" "pe32w4:--prefetch-every 32 --prefetch-warps 4"; do
  tag=${v%%:*}; opt=${v#*:}
  timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 $opt > gpurun_out/p_ttsp_$tag.json 2>>gpurun_out/p_err.log; show gpurun_out/p_ttsp_$tag.json
done
timeout 300 python bench.py --env PointTTSP-v0 --envs 1048576 --no-cpu-baseline --e2e-steps 2 > gpurun_out/p_ttsp_1m.json 2>>gpurun_out/p_err.log; show gpurun_out/p_ttsp_1m.json
timeout 300 python bench.py --env ColourMatch-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 > gpurun_out/p_cm.json 2>>gpurun_out/p_err.log; show gpurun_out/p_cm.json
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/p_tsp.json 2>>gpurun_out/p_err.log; show gpurun_out/p_tsp.json
