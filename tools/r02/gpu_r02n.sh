#!/bin/bash
# round 2, call n: ncu full set of the two encoder kernels (one launch each, after the same command ran plainly)
set -u
mkdir -p gpurun_out
CMD="python tools/bench_encode.py --iters 6"
timeout 300 $CMD > gpurun_out/r02n_plain.json 2> gpurun_out/r02n_err.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"zone_encode_kernel|encoder_head_kernel" -s 6 -c 4 -f -o gpurun_out/r02n_encode $CMD > gpurun_out/r02n_ncu.log 2>&1; echo "ncu rc=$?"
tail -n 4 gpurun_out/r02n_ncu.log
