#!/bin/bash
# round 2, call f: step kernel capped at 120 registers so that the background sampler's CTAs fit beside it
set -u
mkdir -p gpurun_out
L=$PWD/combinatorial_rl_tasks_b200
for v in default r112 r128; do
  echo "== $v"
  if [ $v = default ]; then unset CRL_B200_LIB; else export CRL_B200_LIB=$L/libcrl_b200_$v.so; fi
  timeout 600 python tools/sweep.py PointTTSP-v0:262144 PointTTSP-v0:262144:c1 PointTTSP-v0:262144:c0:s2 PointTTSP-v0:65536 PointTTSP-v0:65536:c0:s2 PointTTSP-v0:65536:c0:s6 PointTTSP-v0:1048576 \
     PointTSP-v0:65536 PointTSP-v0:65536:c0:s7 ColourMatch-v0:262144:c0:s3 --seconds 0.6 2>> gpurun_out/r02f_err.log | tee gpurun_out/r02f_sweep_$v.jsonl | cut -c1-48,80-200
done
unset CRL_B200_LIB
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -x > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/r02f_pytest.log
tail -n 3 gpurun_out/r02f_err.log
