#!/bin/bash
# does the background sampler's shared memory cost the TimedTSP step its 4th CTA per SM?
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
for v in "pf0:--prefetch-every 0" "w1:--prefetch-warps 1" "w2:--prefetch-warps 2" "w4:--prefetch-warps 4" "pe64:--prefetch-every 64" "pe2w1:--prefetch-every 2 --prefetch-warps 1"; do
  tag=${v%%:*}; opt=${v#*:}
  timeout 300 python bench.py --env PointTTSP-v0 --envs 262144 --no-cpu-baseline --e2e-steps 2 --steps 8000 --warmup 800 $opt > gpurun_out/o_ttsp_$tag.json 2>>gpurun_out/o_err.log; show gpurun_out/o_ttsp_$tag.json
done
