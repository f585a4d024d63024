#!/bin/bash
# round 2, call l: the encoder's head kernel (forward = two kernels of this library, no library GEMM)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -m gpu -q -x -s > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/r02l_pytest.log | tail -n 30
timeout 300 python tools/bench_encode.py > gpurun_out/r02l_encode_65536.json 2> gpurun_out/r02l_err.log; echo "bench_encode rc=$?"; cat gpurun_out/r02l_encode_65536.json | cut -c1-1200; tail -n 3 gpurun_out/r02l_err.log
