#!/bin/bash
# A/B: straight-line physics interleaved with the row build (CRL_PHYS_FIRST) vs the default order
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
CRL_B200_LIB=$P/libcrl_b200_physfirst.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reset_and_scale.py -x -q -m gpu > gpurun_out/t_pytest.log 2>&1; echo "pytest(physfirst) rc=$?"; tail -3 gpurun_out/t_pytest.log
for spec in PointTSP-v0:65536 PointTSP-v0:262144 PointTTSP-v0:262144 PointTTSP-v0:1048576 ColourMatch-v0:262144 ColourMatch-v0:65536; do
  env=${spec%%:*}; n=${spec##*:}
  for v in base physfirst; do
    lib=$P/libcrl_b200.so; [ $v = physfirst ] && lib=$P/libcrl_b200_physfirst.so
    CRL_B200_LIB=$lib timeout 300 python bench.py --env $env --envs $n --no-cpu-baseline --e2e-steps 2 --repeats 2 > gpurun_out/t_${env}_${n}_$v.json 2>>gpurun_out/t_err.log; show gpurun_out/t_${env}_${n}_$v.json
  done
done
