#!/bin/bash
# round 2, call ax: ncu full sets of the two kernels of crl_encoder_forward (image-writing zone kernel, bulk-copy head)
set -u
mkdir -p gpurun_out
CMD="python tools/bench_encode.py --iters 6"
timeout 300 $CMD > gpurun_out/r02ax_plain.json 2> gpurun_out/r02ax_err.log &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"zone_encode_kernel|encoder_head_kernel" -s 14 -c 6 -f -o gpurun_out/r02ax_forward $CMD > gpurun_out/r02ax_ncu.log 2>&1; echo "ncu rc=$?"
tail -n 2 gpurun_out/r02ax_ncu.log
