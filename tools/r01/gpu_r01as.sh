#!/bin/bash
# adaptive sampler cadence
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['episode_stats']
    print("%-44s %.3e frac %.3f %.2f us/step pf %d inl %d launches %s" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], d['gpu_launches_detail']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
for spec in PointTSP-v0:65536 PointTTSP-v0:262144 PointTTSP-v0:65536 ColourMatch-v0:262144 ColourMatch-v0:65536 PointTSP-v0:262144; do
  env=${spec%%:*}; n=${spec##*:}
  timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 > gpurun_out/as_${env}_${n}.json 2>>gpurun_out/as_err.log; show gpurun_out/as_${env}_${n}.json
done
tail -3 gpurun_out/as_err.log
