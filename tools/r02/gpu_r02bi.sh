#!/bin/bash
# round 2, call bi: action parked in shared memory until the physics (draw stays early): tests, then A/B against the previous build
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reset_and_scale.py tests/test_gpu_hard.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02bi_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02bi_pytest.log
for lib in old new old new; do
  if [ $lib = old ]; then export CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200_old.so; else unset CRL_B200_LIB; fi
  timeout 300 python tools/sweep.py PointTTSP-v0:262144 PointTSP-v0:65536 ColourMatch-v0:262144 PointTTSP-v0:1048576 --seconds 0.6 2>> gpurun_out/r02bi_err.log | tee -a gpurun_out/r02bi_sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$lib', d['case'], d['frac'], d['best'], d['us_per_step'])"
done
tail -n 2 gpurun_out/r02bi_err.log
