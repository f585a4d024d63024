#!/bin/bash
# prefetch v3 (two slots, work list, lane-per-layout): tests, probe, bench variants
set -u
mkdir -p gpurun_out
P=$PWD/combinatorial_rl_tasks_b200
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/g_pytest.log
for v in libcrl_b200.so libcrl_b200_crl_inline_reset.so; do
  echo "== $v"
  CRL_B200_LIB=$P/$v timeout 300 python tools/probe_prefetch.py PointTTSP-v0 262144 2>&1 | tee -a gpurun_out/probe_prefetch3.log
done
one() { # tag lib env n extra...
  tag=$1; lib=$2; env=$3; n=$4; shift 4
  CRL_B200_LIB=$lib timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --repeats 3 "$@" > gpurun_out/var5_${tag}.json 2>> gpurun_out/var5_err.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var5_${tag}.json").read().strip().splitlines()[-1])
    s=d['episode_stats']
    print("%-44s %.3e  frac %.3f  %.2f us/step  pf %d inl %d  to %s"%("${tag}", d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s.get('chain_wait_timeouts')))
except Exception as e:
    print("${tag} FAILED", e)
PY
}
for spec in PointTTSP-v0:262144 PointTSP-v0:65536 PointTSP-v0:262144 ColourMatch-v0:262144 PointTSP-v0:1048576; do
  env=${spec%%:*}; n=${spec##*:}
  for pf in 8 32 128; do
    one ${env}_${n}_pf${pf} $P/libcrl_b200.so $env $n --prefetch-every $pf
  done
  one ${env}_${n}_pf32_inl $P/libcrl_b200_crl_inline_reset.so $env $n --prefetch-every 32
done
