#!/bin/bash
# TimedTSP: in-kernel reset cost vs concurrent sampler cost (short runs that live off the two parked layouts)
set -u
mkdir -p gpurun_out
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d['episode_stats']
    print("%-40s %.3e frac %.3f %.2f us/step pf %d inl %d episodes %d" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s['episodes']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
for v in "nosampler:--prefetch-every 100000" "pe32:--prefetch-every 32" "notimeouts:--prefetch-every 32 --cfg beta_a=400 --cfg beta_b=0.5"; do
  tag=${v%%:*}; opt=${v#*:}
  timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 --steps 600 --warmup 6 --repeats 1 $opt > gpurun_out/ac_$tag.json 2>>gpurun_out/ac_err.log; show gpurun_out/ac_$tag.json
done
