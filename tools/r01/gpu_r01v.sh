#!/bin/bash
# CTA size A/B (64 / 32 threads vs 128) + stdout check of the 2-rank bench
set -u
mkdir -p gpurun_out
P=$PWD/combinatorial_rl_tasks_b200
show() {
  python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    s=d['episode_stats']
    print("%-40s %.3e frac %.3f %.2f us/step pf %d inl %d" % (sys.argv[1].split('/')[-1], d['value'], d['roofline']['frac'], d['ms_per_step']*1e3, s['resets_prefetched'], s['resets_inline']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
for spec in PointTTSP-v0:262144 PointTTSP-v0:1048576 PointTSP-v0:65536 ColourMatch-v0:262144; do
  env=${spec%%:*}; n=${spec##*:}
  for v in t64 t32; do
    CRL_B200_LIB=$P/libcrl_b200_$v.so timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 > gpurun_out/v_${env}_${n}_$v.json 2>>gpurun_out/v_err.log; show gpurun_out/v_${env}_${n}_$v.json
  done
done
