#!/bin/bash
# first run of the tcgen05 zone encoder: tests (own timeout: a wrong barrier must not hang the box), then the bench
set -u
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_encode.py -x -q -s > gpurun_out/az_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/az_pytest.log
timeout 240 python tools/bench_encode.py > gpurun_out/az_bench_encode.json 2> gpurun_out/az_bench.err; echo "bench rc=$?"; cat gpurun_out/az_bench_encode.json; tail -5 gpurun_out/az_bench.err
