#!/bin/bash
# Round-1 second GPU pass: tests, the three named configs, then ncu on TimedTSP / ColourMatch.
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/b_pytest.log
for spec in PointTSP-v0:65536 PointTTSP-v0:262144 ColourMatch-v0:262144 PointTSP-v0:262144; do
  env=${spec%%:*}; n=${spec##*:}
  python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 10 > gpurun_out/bench_v6_${env}_${n}.json 2>> gpurun_out/b_err.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_v6_${env}_${n}.json").read().strip().splitlines()[-1])
print("${spec}", "%.3e"%d['value'], round(d['roofline']['frac'],3), d['episode_stats']['resets_prefetched'], d['episode_stats']['resets_inline'])
PY
done
python tools/probe_copy.py > gpurun_out/probe_copy.log 2>&1; cat gpurun_out/probe_copy.log
# ncu: steady-state launches (skip far enough that auto-resets are happening)
CMD_T="python bench.py --env PointTTSP-v0 --envs 262144 --steps 1400 --warmup 100 --no-cpu-baseline --e2e-steps 2"
CMD_C="python bench.py --env ColourMatch-v0 --envs 262144 --steps 400 --warmup 50 --no-cpu-baseline --e2e-steps 2"
$CMD_T > gpurun_out/plain_t.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1300 -c 2 -f -o gpurun_out/r01_step_ttsp_262144 $CMD_T > gpurun_out/ncu_t.log 2>&1
echo "ncu ttsp rc=$?"
$CMD_C > gpurun_out/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 300 -c 2 -f -o gpurun_out/r01_step_cm_262144 $CMD_C > gpurun_out/ncu_c.log 2>&1
echo "ncu cm rc=$?"
