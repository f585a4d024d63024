#!/bin/bash
# ncu capture of the zone encoder (after the same command has exited 0 without ncu)
set -u
mkdir -p gpurun_out
CMD="python tools/bench_encode.py --iters 10"
timeout 200 $CMD > gpurun_out/ba_plain.json 2> gpurun_out/ba_err.log; echo "plain rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:zone_encode -s 3 -c 1 -f -o gpurun_out/r01_encode_v3 $CMD > gpurun_out/ba_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ba_ncu.log
ls -la gpurun_out/*.ncu-rep
