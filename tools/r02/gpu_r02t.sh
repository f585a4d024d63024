#!/bin/bash
# round 2, call t: encoder pipeline -- H1 overflow fix (parity), per-tile timeline of the pipeline on one SM
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -m gpu -q > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02t_pytest.log
CRL_B200_LIB=$PWD/combinatorial_rl_tasks_b200/libcrl_b200_tl.so timeout 300 python tools/enc_timeline.py > gpurun_out/r02t_timeline.txt 2>&1; echo "timeline rc=$?"
cat gpurun_out/r02t_timeline.txt
