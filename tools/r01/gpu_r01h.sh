#!/bin/bash
set -u
mkdir -p gpurun_out
P=$PWD/combinatorial_rl_tasks_b200
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/h_pytest.log
one() { # tag lib env n extra...
  tag=$1; lib=$2; env=$3; n=$4; shift 4
  CRL_B200_LIB=$lib timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --repeats 3 "$@" > gpurun_out/var6_${tag}.json 2>> gpurun_out/var6_err.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var6_${tag}.json").read().strip().splitlines()[-1])
    s=d['episode_stats']
    print("%-44s %.3e  frac %.3f  %.2f us/step  pf %d inl %d  to %s"%("${tag}", d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline'], s.get('chain_wait_timeouts')))
except Exception as e:
    print("${tag} FAILED", e)
PY
}
for t in 128 64 32; do
  lib=$P/libcrl_b200_t$t.so; [ $t = 128 ] && lib=$P/libcrl_b200.so
  one TTSP_262144_t${t} $lib PointTTSP-v0 262144 --prefetch-every 128
  one TTSP_65536_t${t} $lib PointTTSP-v0 65536 --prefetch-every 128
  one TTSP_1M_t${t} $lib PointTTSP-v0 1048576 --prefetch-every 128
  one CM_65536_t${t} $lib ColourMatch-v0 65536 --prefetch-every 128
  one TSP_65536_t${t} $lib PointTSP-v0 65536 --prefetch-every 128
done
one TTSP_262144_t128_ch1 $P/libcrl_b200.so PointTTSP-v0 262144 --prefetch-every 128 --chained 1
one TTSP_262144_t32_ch1 $P/libcrl_b200_t32.so PointTTSP-v0 262144 --prefetch-every 128 --chained 1
CMD_T="python bench.py --env PointTTSP-v0 --envs 262144 --steps 1400 --warmup 100 --no-cpu-baseline --e2e-steps 2 --prefetch-every 128"
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1300 -c 2 -f -o gpurun_out/r01c_step_ttsp_262144 $CMD_T > gpurun_out/ncu_t3.log 2>&1
echo "ncu ttsp rc=$?"
