#!/bin/bash
# tests with the one-lane-per-env prefetch sampler, prefetch probe, bench of the named configs
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/e_pytest.log
timeout 300 python tools/probe_prefetch.py PointTTSP-v0 262144 2>&1 | tee gpurun_out/probe_prefetch2.log
timeout 300 python tools/probe_prefetch.py PointTSP-v0 262144 2>&1 | tee -a gpurun_out/probe_prefetch2.log
one() { # tag env n extra...
  tag=$1; env=$2; n=$3; shift 3
  timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --repeats 3 "$@" > gpurun_out/var4_${tag}.json 2>> gpurun_out/var4_err.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var4_${tag}.json").read().strip().splitlines()[-1])
    s=d['episode_stats']
    print("%-44s %.3e  frac %.3f  %.2f us/step  pf %d inl %d  to %s"%("${tag}", d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s.get('chain_wait_timeouts')))
except Exception as e:
    print("${tag} FAILED", e)
PY
}
for spec in PointTSP-v0:65536 PointTSP-v0:262144 PointTTSP-v0:262144 ColourMatch-v0:262144 PointTSP-v0:1048576; do
  env=${spec%%:*}; n=${spec##*:}
  for pf in 0 4 8 32; do
    one ${env}_${n}_pf${pf} $env $n --prefetch-every $pf
  done
done
