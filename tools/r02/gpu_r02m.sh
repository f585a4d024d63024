#!/bin/bash
# round 2, call m: encoder rows from the state planes + steps that do not write zone_obs
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encode.py -m gpu -q -x > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02m_pytest.log
timeout 300 python tools/bench_encode.py > gpurun_out/r02m_encode_65536.json 2> gpurun_out/r02m_err.log; echo "bench_encode rc=$?"; cut -c1-700 gpurun_out/r02m_encode_65536.json; tail -n 3 gpurun_out/r02m_err.log

