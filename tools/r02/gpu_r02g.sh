#!/bin/bash
# round 2, call g: what do TimedTSP's resets cost, and how does it scale with their rate? (diagnostic variants)
set -u
mkdir -p gpurun_out
L=$PWD/combinatorial_rl_tasks_b200
C="PointTTSP-v0:262144 PointTTSP-v0:262144:c0:s2"
echo "== default";   timeout 300 python tools/sweep.py $C --seconds 0.5 2>> gpurun_out/r02g_err.log | cut -c1-48,80-330
echo "== beta(6,1.5)";  timeout 300 python tools/sweep.py $C --seconds 0.5 --cfg beta_a=6 2>> gpurun_out/r02g_err.log | cut -c1-48,80-330
echo "== beta(12,1.5)"; timeout 300 python tools/sweep.py $C --seconds 0.5 --cfg beta_a=12 2>> gpurun_out/r02g_err.log | cut -c1-48,80-330
echo "== beta(1.5,1.5)"; timeout 300 python tools/sweep.py $C --seconds 0.5 --cfg beta_a=1.5 2>> gpurun_out/r02g_err.log | cut -c1-48,80-330
echo "== no counters"; CRL_B200_LIB=$L/libcrl_b200_nocnt.so timeout 300 python tools/sweep.py $C --seconds 0.5 2>> gpurun_out/r02g_err.log | cut -c1-48,130-330
echo "== no reset stores"; CRL_B200_LIB=$L/libcrl_b200_nost.so timeout 300 python tools/sweep.py $C --seconds 0.5 2>> gpurun_out/r02g_err.log | cut -c1-48,130-330
echo "== prefetch_every 128"; timeout 300 python tools/sweep.py PointTTSP-v0:262144:p128 PointTTSP-v0:262144:p8 --seconds 0.5 2>> gpurun_out/r02g_err.log | cut -c1-48,80-330
tail -n 3 gpurun_out/r02g_err.log
