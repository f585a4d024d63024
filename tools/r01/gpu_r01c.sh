#!/bin/bash
# Round-1 third GPU pass: tests (new Beta sampler, chained steps), then launch-shape variants.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/d_pytest.log
one() { # tag lib env n extra...
  tag=$1; lib=$2; env=$3; n=$4; shift 4
  CRL_B200_LIB=$lib timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --repeats 3 "$@" > gpurun_out/var2_${tag}.json 2>> gpurun_out/var2_err.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var2_${tag}.json").read().strip().splitlines()[-1])
    s=d['episode_stats']
    print("%-44s %.3e  frac %.3f  %.2f us/step  pf %d inl %d  to %s"%("${tag}", d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s.get('chain_wait_timeouts')))
except Exception as e:
    print("${tag} FAILED", e)
PY
}
P=combinatorial_rl_tasks_b200
for spec in PointTSP-v0:65536 PointTSP-v0:262144 PointTTSP-v0:262144 ColourMatch-v0:262144 PointTSP-v0:1048576; do
  env=${spec%%:*}; n=${spec##*:}
  for t in 128 64 32; do
    lib=$PWD/$P/libcrl_b200_t$t.so; [ $t = 128 ] && lib=$PWD/$P/libcrl_b200.so
    for ch in 0 1; do
      one ${env}_${n}_t${t}_ch${ch}_pf0 $lib $env $n --chained $ch --prefetch-every 0
    done
  done
  one ${env}_${n}_t128_ch1_pf8 $PWD/$P/libcrl_b200.so $env $n --chained 1 --prefetch-every 8
  one ${env}_${n}_t64_ch1_pf8 $PWD/$P/libcrl_b200_t64.so $env $n --chained 1 --prefetch-every 8
done
