#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "worst|passed|failed|Error" | tail -5
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['episode_stats']
    print("%-44s %.3e frac %.3f %.2f us/step e2e %.3e" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, d['e2e']['value']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
for spec in PointTSP-v0:65536 PointTSP-v0:262144 PointTTSP-v0:262144 ColourMatch-v0:262144 ColourMatch-v0:65536 ColourMatch-v0:1048576; do
  env=${spec%%:*}; n=${spec##*:}
  timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 20 --steps 16000 --warmup 1600 > gpurun_out/ai_${env}_${n}.json 2>>gpurun_out/ai_err.log; show gpurun_out/ai_${env}_${n}.json
done
