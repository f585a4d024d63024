#!/bin/bash
# ncu captures at the HEAD of round 1 (each after the same command exited 0 without ncu): launch list of the default
# bench, full set of the PointTSP 65,536 step kernel, full set of the zone encoder
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 210 --warmup 21 --no-cpu-baseline --e2e-steps 3"
timeout 300 $CMD > gpurun_out/bd_plain.json 2>gpurun_out/bd_err.log; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"step_kernel|prefetch|reset_kernel|gather" -c 700 --csv --log-file gpurun_out/r01_head_launches_pointtsp_65536.csv $CMD > gpurun_out/bd_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 3 -f -o gpurun_out/r01_head_step_tsp_65536 $CMD > gpurun_out/bd_ncu2.log 2>&1; echo "ncu tsp rc=$?"
CMD_E="python tools/bench_encode.py --iters 10"
timeout 200 $CMD_E > gpurun_out/bd_plain_enc.json 2>> gpurun_out/bd_err.log; echo "plain enc rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:zone_encode -s 3 -c 1 -f -o gpurun_out/r01_head_encode $CMD_E > gpurun_out/bd_ncu3.log 2>&1; echo "ncu enc rc=$?"
ls -la gpurun_out/*.ncu-rep
