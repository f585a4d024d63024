#!/bin/bash
# zone encoder v4: pooling by a third GEMM (default) vs the shuffle butterfly (CRL_ENC_POOL=0); mode 2 = LBO/SBO swapped (diagnostic)
set -u
mkdir -p gpurun_out
for mode in 1 2 0; do
  CRL_ENC_POOL=$mode timeout 200 python -m pytest tests/test_gpu_encode.py -q -s -k "fixture or random" > gpurun_out/bb_pytest_$mode.log 2>&1; echo "mode $mode pytest rc=$?"; grep -E "vs bf16|passed|failed" gpurun_out/bb_pytest_$mode.log | grep -v print | cut -c1-150
done
for mode in 1 0; do
  CRL_ENC_POOL=$mode timeout 200 python tools/bench_encode.py > gpurun_out/bb_bench_$mode.json 2>> gpurun_out/bb_err.log; echo "mode $mode bench rc=$?"; cut -c1-420 gpurun_out/bb_bench_$mode.json
done
