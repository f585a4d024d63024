#!/bin/bash
# chained steps (per-warp ordering across launches) at the larger batches: does it hide the reset tail?
set -u
mkdir -p gpurun_out
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['episode_stats']
    print("%-44s %.3e frac %.3f %.2f us/step pf %d inl %d to %d" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s['chain_wait_timeouts']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
for spec in PointTTSP-v0:262144 PointTTSP-v0:1048576 PointTSP-v0:262144 PointTSP-v0:1048576 ColourMatch-v0:262144 ColourMatch-v0:1048576; do
  env=${spec%%:*}; n=${spec##*:}
  for c in 0 1; do
    timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --steps 16000 --warmup 1600 --chained $c > gpurun_out/ad_${env}_${n}_c$c.json 2>>gpurun_out/ad_err.log; show gpurun_out/ad_${env}_${n}_c$c.json
  done
done
