#!/bin/bash
# encoder with 8-slot packing for N <= 8: tests, then ColourMatch 262,144 and PointTSP 65,536
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_encode.py -x -q -s > gpurun_out/bg_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "vs bf16|passed|failed" gpurun_out/bg_pytest.log | grep -v print | cut -c1-120
timeout 200 python tools/bench_encode.py --envs 262144 --env ColourMatch-v0 > gpurun_out/bg_enc_cm.json 2>> gpurun_out/bg_err.log; echo "enc cm rc=$?"; cut -c1-330 gpurun_out/bg_enc_cm.json
timeout 200 python tools/bench_encode.py > gpurun_out/bg_enc_tsp.json 2>> gpurun_out/bg_err.log; echo "enc tsp rc=$?"; cut -c1-330 gpurun_out/bg_enc_tsp.json
