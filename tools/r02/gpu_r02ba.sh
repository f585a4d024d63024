#!/bin/bash
# round 2, call ba: TimedTSP 262,144 against the sampler cadence, replica / stream count and sampler warps
set -u
mkdir -p gpurun_out
timeout 500 python tools/sweep.py PointTTSP-v0:262144 PointTTSP-v0:262144:p16 PointTTSP-v0:262144:p64 PointTTSP-v0:262144:p128 \
  PointTTSP-v0:262144:r3 PointTTSP-v0:262144:r4 PointTTSP-v0:262144:r4:p64 PointTTSP-v0:262144:w1 PointTTSP-v0:262144:w1:p64 \
  PointTTSP-v0:262144:b100 PointTTSP-v0:1048576 PointTTSP-v0:1048576:p64 --seconds 0.6 > gpurun_out/r02ba_sweep.jsonl 2> gpurun_out/r02ba_err.log
echo "rc=$?"; tail -n 3 gpurun_out/r02ba_err.log; cut -c1-44,70-330 gpurun_out/r02ba_sweep.jsonl
