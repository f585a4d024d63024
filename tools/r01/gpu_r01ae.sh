#!/bin/bash
# reset path without local memory
set -u
mkdir -p gpurun_out
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['episode_stats']
    print("%-44s %.3e frac %.3f %.2f us/step pf %d inl %d" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
timeout 900 python -m pytest tests/test_gpu_reset_and_scale.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for spec in PointTTSP-v0:262144 PointTTSP-v0:1048576 PointTTSP-v0:65536 PointTSP-v0:65536 PointTSP-v0:262144 ColourMatch-v0:262144; do
  env=${spec%%:*}; n=${spec##*:}
  timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --steps 16000 --warmup 1600 > gpurun_out/ae_${env}_${n}.json 2>>gpurun_out/ae_err.log; show gpurun_out/ae_${env}_${n}.json
done
