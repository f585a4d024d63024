#!/bin/bash
# Diagnostic: how much of a step's time is the reset machinery?  (not a bench line)
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py --no-cpu-baseline --e2e-steps 2 "$@" > gpurun_out/var_$tag.json 2>> gpurun_out/var_err.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/var_$tag.json").read().strip().splitlines()[-1])
s=d['episode_stats']
print("%-28s %.3e  frac %.3f  %.2f us/step  prefetched %d inline %d"%("$tag", d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline']))
PY
}
for spec in PointTTSP-v0:262144 ColourMatch-v0:262144 PointTSP-v0:65536; do
  env=${spec%%:*}; n=${spec##*:}
  run ${env}_noreset --env $env --envs $n --no-auto-reset --prefetch-every 0
  run ${env}_inline --env $env --envs $n --prefetch-every 0
  run ${env}_pf8 --env $env --envs $n --prefetch-every 8
  run ${env}_pf2 --env $env --envs $n --prefetch-every 2
  run ${env}_pf32 --env $env --envs $n --prefetch-every 32
done
