#!/bin/bash
# final ncu captures of round 1 at HEAD (each after the same command exited 0 without ncu)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 210 --warmup 21 --no-cpu-baseline --e2e-steps 3"
timeout 300 $CMD > gpurun_out/au_plain.json 2>gpurun_out/au_err.log; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"step_kernel|prefetch|reset_kernel|gather" -c 700 --csv --log-file gpurun_out/r01_final_launches_pointtsp_65536.csv $CMD > gpurun_out/au_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 60 -c 3 -f -o gpurun_out/r01_final_step_tsp_65536 $CMD > gpurun_out/au_ncu2.log 2>&1; echo "ncu tsp rc=$?"
CMD_C="python bench.py --env ColourMatch-v0 --envs 262144 --steps 300 --warmup 30 --no-cpu-baseline --e2e-steps 2"
timeout 300 $CMD_C > gpurun_out/au_plain_cm.json 2>>gpurun_out/au_err.log; echo "plain cm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 200 -c 2 -f -o gpurun_out/r01_final_step_cm_262144 $CMD_C > gpurun_out/au_ncu3.log 2>&1; echo "ncu cm rc=$?"
