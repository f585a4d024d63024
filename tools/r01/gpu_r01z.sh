#!/bin/bash
# are the same-address counter atomics what slows reset-heavy TimedTSP?
set -u
mkdir -p gpurun_out
P=$PWD/combinatorial_rl_tasks_b200
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-40s %.3e frac %.3f %.2f us/step" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
for spec in PointTTSP-v0:262144 PointTTSP-v0:1048576 PointTTSP-v0:65536; do
  env=${spec%%:*}; n=${spec##*:}
  for v in base nocnt; do
    lib=$P/libcrl_b200.so; [ $v = nocnt ] && lib=$P/libcrl_b200_crl_no_counters.so
    CRL_B200_LIB=$lib timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --steps 16000 --warmup 1600 > gpurun_out/z_${env}_${n}_$v.json 2>>gpurun_out/z_err.log; show gpurun_out/z_${env}_${n}_$v.json
  done
done
