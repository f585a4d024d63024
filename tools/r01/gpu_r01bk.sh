#!/bin/bash
# last check of round 1: encoder tests, the rollout example test, smoke()
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_encode.py tests/test_gpu_rollout.py -x -q > gpurun_out/bk_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/bk_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bk_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/bk_smoke.log
