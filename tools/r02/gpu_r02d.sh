#!/bin/bash
# round 2, call d: which of the three changes since call a cost the hot path?  (variants of the same source)
set -u
mkdir -p gpurun_out
CASES="PointTSP-v0:65536 PointTTSP-v0:262144 ColourMatch-v0:262144:c1 PointTSP-v0:262144"
L=$PWD/combinatorial_rl_tasks_b200
for v in old default early noinl nopf; do
  echo "== $v"
  if [ $v = default ]; then unset CRL_B200_LIB; else export CRL_B200_LIB=$L/libcrl_b200_$v.so; fi
  timeout 600 python tools/sweep.py $CASES --seconds 0.6 2>> gpurun_out/r02d_err.log | tee gpurun_out/r02d_sweep_$v.jsonl | cut -c1-40,100-215
done
tail -5 gpurun_out/r02d_err.log
