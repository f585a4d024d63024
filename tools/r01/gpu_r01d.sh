#!/bin/bash
# tests, then: plain-mode regression bisect (out-of-line vs inline reset), chained on/off
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/d_pytest.log
one() { # tag lib env n extra...
  tag=$1; lib=$2; env=$3; n=$4; shift 4
  CRL_B200_LIB=$lib timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --repeats 3 "$@" > gpurun_out/var3_${tag}.json 2>> gpurun_out/var3_err.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var3_${tag}.json").read().strip().splitlines()[-1])
    s=d['episode_stats']
    print("%-44s %.3e  frac %.3f  %.2f us/step  pf %d inl %d  to %s"%("${tag}", d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s.get('chain_wait_timeouts')))
except Exception as e:
    print("${tag} FAILED", e)
PY
}
P=$PWD/combinatorial_rl_tasks_b200
for spec in PointTSP-v0:65536 PointTSP-v0:262144 PointTSP-v0:1048576 ColourMatch-v0:262144; do
  env=${spec%%:*}; n=${spec##*:}
  one ${env}_${n}_default_ch0 $P/libcrl_b200.so $env $n --chained 0 --prefetch-every 0
  one ${env}_${n}_inline_ch0 $P/libcrl_b200_crl_inline_reset.so $env $n --chained 0 --prefetch-every 0
  one ${env}_${n}_t64_ch0 $P/libcrl_b200_t64.so $env $n --chained 0 --prefetch-every 0
  one ${env}_${n}_default_ch1 $P/libcrl_b200.so $env $n --chained 1 --prefetch-every 0
done
one TTSP_noreset_ch0 $P/libcrl_b200.so PointTTSP-v0 262144 --chained 0 --prefetch-every 0 --no-auto-reset
one TTSP_noreset_inl_ch0 $P/libcrl_b200_crl_inline_reset.so PointTTSP-v0 262144 --chained 0 --prefetch-every 0 --no-auto-reset
