#!/bin/bash
# round 2, call c: inline reset path + one-wave-ahead L2 prefetch + occupancy variants
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02c_pytest.log
CASES="PointTSP-v0:65536 PointTTSP-v0:262144 PointTTSP-v0:262144:c1 PointTTSP-v0:65536 ColourMatch-v0:262144 ColourMatch-v0:262144:c1 PointTSP-v0:262144 PointTTSP-v0:1048576"
L=combinatorial_rl_tasks_b200
echo "== default, prefetch ahead auto"; timeout 600 python tools/sweep.py $CASES --seconds 0.7 2> gpurun_out/r02c_err.log | tee gpurun_out/r02c_sweep_default.jsonl | cut -c1-200
echo "== default, no prefetch ahead"; CRL_PF_AHEAD=0 timeout 600 python tools/sweep.py $CASES --seconds 0.7 2>> gpurun_out/r02c_err.log | tee gpurun_out/r02c_sweep_pf0.jsonl | cut -c1-200
echo "== w24"; CRL_B200_LIB=$PWD/$L/libcrl_b200_w24.so timeout 600 python tools/sweep.py $CASES --seconds 0.7 2>> gpurun_out/r02c_err.log | tee gpurun_out/r02c_sweep_w24.jsonl | cut -c1-200
echo "== w20"; CRL_B200_LIB=$PWD/$L/libcrl_b200_w20.so timeout 600 python tools/sweep.py $CASES --seconds 0.7 2>> gpurun_out/r02c_err.log | tee gpurun_out/r02c_sweep_w20.jsonl | cut -c1-200
echo "== no resets (beta trick), default"; timeout 300 python tools/sweep.py PointTTSP-v0:262144 --seconds 0.7 --cfg beta_a=400 --cfg beta_b=0.5 2>> gpurun_out/r02c_err.log | cut -c1-200
echo "== bank 100, default"; timeout 300 python tools/sweep.py PointTTSP-v0:262144:b100 --seconds 0.7 2>> gpurun_out/r02c_err.log | cut -c1-200
tail -5 gpurun_out/r02c_err.log
