#!/bin/bash
# round 2, call ar: ncu full sets of the final encoder kernels (plain zone kernel, image-writing zone kernel, bulk-copy head)
set -u
mkdir -p gpurun_out
CMD="python tools/bench_encode.py --iters 6"
timeout 300 $CMD > gpurun_out/r02ar_plain.json 2> gpurun_out/r02ar_err.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"zone_encode_kernel|encoder_head_kernel" -s 16 -c 8 -f -o gpurun_out/r02ar_encode $CMD > gpurun_out/r02ar_ncu.log 2>&1; echo "ncu encode rc=$?"
tail -n 3 gpurun_out/r02ar_ncu.log
